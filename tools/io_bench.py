"""HBM roofline of the image input / output kernels (csrc/imageio.cu) and the EMA sweep."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.training import means, stds  # noqa: E402

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


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


N = 16384  # 16384 x 3 x 64 x 64: 201 MB of uint8 in, 805 MB of fp32 out (larger than L2)
u8 = torch.randint(0, 256, (N, 64, 64, 3), device=dev, dtype=torch.uint8)
t = timeit(lambda: ops.u8_to_f32_norm(u8, means, stds))
by = u8.numel() * 5
print(f"u8_to_f32_norm  {N}x64x64x3: {t:.3f} ms  {by/t/1e6:.0f} GB/s  ({by/t/1e6/peak:.2f} of measured HBM copy peak)")
x = torch.randn(N, 3, 64, 64, device=dev)
t = timeit(lambda: ops.denorm_grid_u8(x, 128, 0, means, stds))
by = x.numel() * 5
print(f"denorm_grid_u8  {N}x3x64x64 nrow 128: {t:.3f} ms  {by/t/1e6:.0f} GB/s  ({by/t/1e6/peak:.2f})")
p = torch.randn(30945156 // 4 * 4, device=dev)
e = p.clone()
t = timeit(lambda: ops.ema_update(e, p, 0.999))
by = p.numel() * 12
print(f"ema_update 30.9 M params: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s  ({by/t/1e6/peak:.2f})")
