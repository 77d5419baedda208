"""The 128 -> 3 / 4 tail convolution (training forward) and the fused sampling tail: correctness against torch conv2d on
the same bf16-rounded operands and CUDA-event timing with the achieved HBM rate.
usage: tail_bench.py [B]   (env TSD_TAIL_Y=0: the tap-by-tap kernel)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tag = f"TSD_TAIL_Y={os.environ.get('TSD_TAIL_Y', '1')}"
g = torch.Generator(device="cuda").manual_seed(0)


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


for (co, H, n) in ((3, 64, B), (4, 16, 16 * B), (3, 32, B)):
    W = H
    a = torch.randn(2 * n * H * W, 128, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(co, 128, 3, 3, device=dev, generator=g) * 0.03
    bias = torch.randn(co, device=dev, generator=g) * 0.1
    # training forward on the first n images
    out = ops.tail_conv_fwd(a[:n * H * W], w, bias, n, H, W)
    nb = min(n, 4)
    ref = F.conv2d(a[:nb * H * W].float().view(nb, H, W, 128).permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1)
    err = (out[:nb] - ref).abs().max().item()
    last = F.conv2d(a[(n - 1) * H * W:n * H * W].float().view(1, H, W, 128).permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1)
    err_last = (out[n - 1:] - last).abs().max().item()
    tf = timeit(lambda: ops.tail_conv_fwd(a[:n * H * W], w, bias, n, H, W, out=out))
    by = n * H * W * 256 + n * co * H * W * 4
    print(f"{tag} tail fwd  co={co} {H}x{W} n={n}: max abs err {err:.2e} (last image {err_last:.2e}) | {tf * 1e3:.1f} us  {by / tf / 1e6:.0f} GB/s")
    # fused sampling tail: conditional rows [0, n), unconditional rows [n, 2n)
    T = 1000
    c1 = torch.rand(T, device=dev) + 0.5
    c2 = torch.rand(T, device=dev) * 0.1
    sig = torch.rand(T, device=dev) * 0.1
    step = torch.tensor([500], device=dev, dtype=torch.int32)
    nan_flag = torch.zeros(1, device=dev, dtype=torch.int32)
    x0 = torch.randn(2 * n, co, H, W, device=dev, generator=g)
    z = torch.randn(n, co, H, W, device=dev, generator=g)
    x = x0.clone()
    eps = torch.empty(2 * n, co, H, W, device=dev)
    ops.tail_conv_sample(a, w, bias, x, n, H, W, step, c1, c2, sig, 1.8, nan_flag, noise=z, eps_out=eps)
    ec = ops.tail_conv_fwd(a[:n * H * W], w, bias, n, H, W)
    eu = ops.tail_conv_fwd(a[n * H * W:], w, bias, n, H, W)
    ep = 2.8 * ec - 1.8 * eu
    want = c1[500] * x0[:n] - c2[500] * ep + sig[500] * z
    e_fused = (x[:n] - want).abs().max().item()
    same = torch.equal(x[:n], x[n:]) and torch.equal(eps[:n], ec) and torch.equal(eps[n:], eu)
    ts = timeit(lambda: ops.tail_conv_sample(a, w, bias, x, n, H, W, step, c1, c2, sig, 1.8, nan_flag))
    by = 2 * n * H * W * 256 + 3 * n * co * H * W * 4
    print(f"{tag} tail sample co={co} {H}x{W} pairs={n}: update max abs err {e_fused:.2e}, halves / eps consistent {same}, nan {int(nan_flag.item())} | {ts * 1e3:.1f} us  {by / ts / 1e6:.0f} GB/s")
