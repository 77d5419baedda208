"""The ci -> 128 head convolution and the data gradient of the 128 -> co tail convolution (fp32 NCHW image in, bf16 NHWC out):
error against torch in fp32 and CUDA-event timing with the achieved HBM rate of the output.
usage: head_bench.py [B]   (env TSD_IN_CONV_TF32=0: the CUDA-core kernels)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tag = f"TSD_IN_CONV_TF32={os.environ.get('TSD_IN_CONV_TF32', '1')}"
g = torch.Generator(device="cuda").manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, n=20):
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


for (ci, H, n) in ((3, 64, B), (4, 16, 16 * B), (3, 32, B)):
    W = H
    x = torch.randn(n, ci, H, W, device=dev, generator=g) * 3
    w = torch.randn(128, ci, 3, 3, device=dev, generator=g) * 0.2
    bias = torch.randn(128, device=dev, generator=g) * 0.1
    out = ops.head_conv_fwd(x, w, bias).float()
    nb = min(n, 4)
    for lo in (0, n - nb):
        ref = F.conv2d(x[lo:lo + nb], w, bias, padding=1).permute(0, 2, 3, 1).reshape(nb * H * W, 128)
        got = out[lo * H * W:(lo + nb) * H * W]
        e = ((got - ref).norm() / ref.norm()).item()
        e_bf = ((ref.to(torch.bfloat16).float() - ref).norm() / ref.norm()).item()
        print(f"{tag} head fwd ci={ci} {H}x{W} images {lo}..{lo + nb - 1}: rel-L2 {e:.3e} (bf16 rounding of the exact result alone: {e_bf:.3e})")
    tf = timeit(lambda: ops.head_conv_fwd(x, w, bias))
    print(f"{tag} head fwd ci={ci} {H}x{W} n={n}: {tf * 1e3:.1f} us  {n * H * W * 256 / tf / 1e6:.0f} GB/s")
    # tail data gradient: dy fp32 [n, co, H, W], w [co, 128, 3, 3] -> da bf16 [n*H*W, 128]
    co = ci
    dy = torch.randn(n, co, H, W, device=dev, generator=g)
    wt = torch.randn(co, 128, 3, 3, device=dev, generator=g) * 0.05
    a = torch.empty(n * H * W, 128, device=dev, dtype=torch.bfloat16)
    da = ops.tail_conv_dgrad(dy, a, wt, n, H, W).float()
    for lo in (0, n - nb):
        ref = F.conv_transpose2d(dy[lo:lo + nb], wt, padding=1).permute(0, 2, 3, 1).reshape(nb * H * W, 128)
        got = da[lo * H * W:(lo + nb) * H * W]
        e = ((got - ref).norm() / ref.norm()).item()
        print(f"{tag} tail dgrad co={co} {H}x{W} images {lo}..{lo + nb - 1}: rel-L2 {e:.3e}")
    td = timeit(lambda: ops.tail_conv_dgrad(dy, a, wt, n, H, W))
    print(f"{tag} tail dgrad co={co} {H}x{W} n={n}: {td * 1e3:.1f} us  {n * H * W * 256 / td / 1e6:.0f} GB/s")
