"""A small pass over every kernel family for compute-sanitizer (memcheck / synccheck / initcheck):
  compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_step.py
One eager training step and two reverse-diffusion steps of the 3x64x64 UNet at batch 2 (tcgen05 GEMM / halo and patch
convolutions, tcgen05 attention forward / backward for head_dim 16 and 32, mma.sync attention at L = 64, GroupNorm,
LayerNorm, GEGLU, conditioning, DDPM elementwise, clip + AdamW + EMA), one latent 4x16x16 reverse step, the VQ-VAE codec
and the image I/O kernels.  Sizes are small because the sanitizer slows kernels by one to two orders of magnitude."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM  # noqa: E402
from from_ddpm_to_stable_diffusion_b200 import training  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.vqvae import VQVAE  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2

model = Diffusion(3, [1, 2, 2, 2], 128, num_class=3, dropout=0.1).to(dev).train()
trainer = TrainerDDPM(model, 0.0015, 0.0195, 1000).to(dev)
opt = FusedClipAdamW(model, lr=2e-6, weight_decay=1e-5, max_norm=1.0, ema_decay=0.999)
x = torch.randn(B, 3, 64, 64, device=dev)
y = torch.randint(0, 4, (B,), device=dev)
for _ in range(2):
    loss = training.train_step(trainer, opt, x, y)
torch.cuda.synchronize()
print("train step ok, loss", float(loss))

model.eval()
sampler = SamplerDDPM(model, 0.0015, 0.0195, 1000, w=1.8).to(dev)
sampler.use_cuda_graph = False
out = sampler(torch.randn(B, 3, 64, 64, device=dev), torch.randint(1, 4, (B,), device=dev), steps=range(999, 997, -1))
torch.cuda.synchronize()
print("reverse steps ok", bool(torch.isfinite(out).all()))

lat = Diffusion(4, [1, 2, 2, 2], 128, num_class=10).to(dev).eval()
ls = SamplerDDPM(lat, 0.0015, 0.0195, 1000, w=1.8).to(dev)
ls.use_cuda_graph = False
z = ls(torch.randn(B, 4, 16, 16, device=dev), torch.randint(1, 11, (B,), device=dev), steps=range(999, 998, -1))
torch.cuda.synchronize()
print("latent reverse step ok", bool(torch.isfinite(z).all()))

vq = VQVAE(in_channels=3, embedding_dim=4, num_embeddings=64, hidden_dims=[32, 64], img_size=64).to(dev).eval()
img = torch.rand(B, 3, 64, 64, device=dev) * 2 - 1
latv = vq.encode(img)[0]
rec = vq(img)
torch.cuda.synchronize()
print("codec ok", tuple(latv.shape))

u8 = torch.randint(0, 256, (B, 64, 64, 3), device=dev, dtype=torch.uint8)
f = training.normalize_u8(u8)
g = training.image_grid_u8(f, nrow=2)
torch.cuda.synchronize()
print("image io ok", tuple(f.shape), tuple(g.shape))
print("SANITIZE_PASS_DONE")
