"""3x3 convolution in halo mode (TSD_CONV_HALO = 0 / 1 / 2 / 3, read once per process): error against torch fp32 on the
same bf16 inputs and time at the benchmark shapes.  Usage: TSD_CONV_HALO=m python tools/conv_halo_check.py [batch]"""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = os.environ.get("TSD_CONV_HALO", "3 (default)")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
# correctness on a small problem (2 images, 64x64 and 32x32, concat input, bias + residual)
for (H, W, c0, c1, cout) in [(64, 64, 64, 64, 128), (32, 32, 128, 0, 256), (16, 16, 128, 0, 256), (16, 32, 64, 64, 128)]:
    n = 2
    cin = c0 + c1
    x = torch.randn(n, H, W, cin, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05)
    wp = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().to(torch.bfloat16)
    bias = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(n * H * W, cout, device=dev, generator=g).to(torch.bfloat16)
    x0 = x[..., :c0].contiguous().view(n * H * W, c0)
    x1 = x[..., c0:].contiguous().view(n * H * W, c1) if c1 else None
    y = ops.conv3x3(x0, n, H, W, wp, cout, x1=x1, bias=bias, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1).permute(0, 2, 3, 1).reshape(n * H * W, cout) + res.float()
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    dy = torch.randn(n * H * W, cout, device=dev, generator=g).to(torch.bfloat16)
    dx = ops.conv3x3_dgrad(dy, n, H, W, wp, cin)
    dref = F.conv_transpose2d(dy.float().view(n, H, W, cout).permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1).reshape(n * H * W, cin)
    derr = (dx.float() - dref).abs().max().item() / dref.abs().max().item()
    print(f"mode {mode} H={H} W={W} {cin}->{cout}: fwd rel err {err:.2e}  dgrad rel err {derr:.2e}  {'OK' if err < 2e-2 and derr < 2e-2 else 'WRONG'}", flush=True)
for (H, cin, cout) in [(64, 128, 128), (64, 256, 128), (32, 256, 256), (32, 512, 256), (16, 256, 256), (16, 512, 256)]:
    M = B * H * H
    x = torch.randn(M, cin, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(M, cout, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=dev, generator=g) * 0.03).to(torch.bfloat16)
    fl = 2.0 * M * cout * 9 * cin
    tf = timeit(lambda: ops.conv3x3(x, B, H, H, w, cout))
    td = timeit(lambda: ops.conv3x3_dgrad(dy, B, H, H, w, cin))
    print(f"mode {mode} H={H} {cin}->{cout}: fwd {tf*1e3:7.1f} us {fl/tf/1e9:6.0f} TF/s | dgrad {td*1e3:7.1f} us {fl/td/1e9:6.0f} TF/s", flush=True)
