"""Timing of tsd_gemm_fwd / tsd_conv3x3_fwd at chosen shapes: python tools/gemm_bench.py [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576


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
for K, N, extra in [(128, 128, "plain"), (128, 128, "bias"), (128, 128, "bias+res"), (128, 384, "plain"), (128, 1024, "geglu"),
                    (256, 128, "plain"), (512, 128, "plain"), (1024, 128, "plain"), (1152, 128, "plain"), (2304, 128, "plain")]:
    a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev, generator=g) if extra != "plain" else None
    res = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16) if extra == "bias+res" else None
    fn = lambda: ops.gemm(a, w, N, bias=bias, residual=res, geglu=(extra == "geglu"))
    t = timeit(fn)
    nd = N // 2 if extra == "geglu" else N
    byts = (M * K + M * nd + (M * N if res is not None else 0)) * 2
    tiles = (M // 128) * (N // 128)
    print(f"K={K:5d} N={N:5d} {extra:9s}: {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:7.1f} TF/s  {byts/t/1e6:7.0f} GB/s  "
          f"{t*1e3/ (tiles/148):6.2f} us/tile/CTA")
