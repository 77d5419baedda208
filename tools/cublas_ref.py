"""Context numbers: cuBLAS (torch.matmul, bf16) on the plain-GEMM shapes of the UNet's dominant contractions, next to
this library's kernels on the same shapes (the 3x3 convolutions as implicit GEMMs: no im2col matrix exists here, the
cuBLAS line is given the already-unfolded [M, 9C] operand for free)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device="cuda").manual_seed(0)
M = B * 64 * 64
for (K, N, what) in [(1152, 128, "conv3x3 128->128 @64x64"), (2304, 128, "conv3x3 256->128 @64x64"), (128, 128, "1x1 / linear C->C"),
                     (128, 1024, "linear C->8C"), (512, 128, "linear 4C->C"), (128, 384, "in_proj C->3C")]:
    a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.03).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t_cublas = timeit(lambda: torch.matmul(a, w.t(), out=out))
    fl = 2.0 * M * N * K
    line = f"M={M} K={K} N={N} ({what}): cuBLAS {t_cublas * 1e3:8.1f} us {fl / t_cublas / 1e9:6.0f} TF/s"
    if K % 1152 == 0:
        cin = K // 9
        x = torch.randn(M, cin, device=dev, generator=g).to(torch.bfloat16)
        t_ours = timeit(lambda: ops.conv3x3(x, B, 64, 64, w, N))
        line += f" | ours (implicit GEMM from the NHWC tensor) {t_ours * 1e3:8.1f} us {fl / t_ours / 1e9:6.0f} TF/s"
    else:
        t_ours = timeit(lambda: ops.gemm(a, w, N))
        line += f" | ours {t_ours * 1e3:8.1f} us {fl / t_ours / 1e9:6.0f} TF/s  ({(M * K + M * N) * 2 / t_ours / 1e6:5.0f} GB/s)"
    print(line)
    del a, w, out
