"""Run warm-up steps, then ONE profiled training step and ONE profiled reverse-diffusion step between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--mode", default="train", choices=["train", "sample", "both"])
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = Diffusion(3, [1, 2, 2, 2], 128, num_class=3, dropout=0.1).to(dev).train()
B = args.batch
x = torch.randn(B, 3, 64, 64, device=dev)
y = torch.randint(1, 4, (B,), device=dev)
rt = torch.cuda.cudart()
if args.mode in ("train", "both"):
    trainer = TrainerDDPM(model, 0.0015, 0.0195, 1000).to(dev)
    opt = FusedClipAdamW(model, lr=2e-6, weight_decay=1e-5, max_norm=1.0)

    def step():
        opt.zero_grad()
        loss = trainer(x, y).sum() / B ** 2
        loss.backward()
        opt.step()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
if args.mode in ("sample", "both"):
    model.eval()
    sampler = SamplerDDPM(model, 0.0015, 0.0195, 1000, w=1.8).to(dev)
    sampler.use_cuda_graph = False
    sampler(x, y, steps=range(999, 999 - args.warmup, -1))
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    sampler(x, y, steps=[990])
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
print("done")
