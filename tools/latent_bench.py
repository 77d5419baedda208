"""BASELINE configs[4]: latent-space DDPM sampling on 4x16x16 latents (03_train_with_vae.py path: Diffusion(channel_img=4,
num_class=10)), batch 4096 per GPU, CFG w=1.8: ms per reverse step and images/s at T=1000."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
model = Diffusion(4, [1, 2, 2, 2], 128, num_class=10, dropout=0.1).to(dev).eval()
sampler = SamplerDDPM(model, 0.0015, 0.0195, 1000, w=1.8).to(dev)
xT = torch.randn(B, 4, 16, 16, device=dev)
y = torch.randint(1, 11, (B,), device=dev)
sampler(xT, y, steps=range(999, 995, -1))  # capture + warm-up
torch.cuda.synchronize()
k = 16
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = sampler(xT, y, steps=range(999, 999 - k, -1))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / k
gf = 2 * 2.272  # necessary GF per image-step, before the shared prefix
print(f"latent 4x16x16 batch {B}: {ms:.2f} ms per reverse step -> {B / (ms * 1e-3 * 1000):.1f} images/s at T=1000, "
      f"{gf * B / ms:.0f} TF/s, finite={torch.isfinite(out).all().item()}")
