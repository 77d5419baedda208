"""Two steady-state training iterations (captured graph replays) inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/ncu_step.py
Also usable without ncu (prints the CUDA-event time of the two replays).  Optional argv[1]: batch (default 256);
argv[2] = "sample": profile two reverse-diffusion steps (256 images, CFG) instead."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = Diffusion(3, [1, 2, 2, 2], 128, num_class=3, dropout=0.1).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if mode == "train":
    model.train()
    trainer = TrainerDDPM(model, 0.0015, 0.0195, 1000).to(dev)
    opt = FusedClipAdamW(model, lr=2e-6, weight_decay=1e-5, max_norm=1.0)
    step = GraphedTrainStep(trainer, opt, train_rand=0.05)
    x = torch.randn(B, 3, 64, 64, device=dev)
    y = torch.randint(0, 3, (B,), device=dev)
    for _ in range(3):
        step(x, y)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(2):
        step(x, y)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"2 training iterations, batch {B}: {e0.elapsed_time(e1) / 2:.3f} ms per step, {step.launches_per_step()} library kernels per step")
else:
    model.eval()
    s = SamplerDDPM(model, 0.0015, 0.0195, 1000, w=1.8).to(dev)
    xT = torch.randn(B, 3, 64, 64, device=dev)
    ys = torch.randint(1, 4, (B,), device=dev)
    NS = int(os.environ.get("TSD_NCU_SAMPLE_STEPS", "3"))  # replays inside the profiled range (use ~40 for plain timing)
    s(xT, ys, steps=range(999, 993, -1))
    s(xT, ys, steps=range(999, 996, -1))  # hoisted conditioning and graph are warm now
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    s(xT, ys, steps=range(999, 999 - NS, -1))
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"{NS} reverse steps, {B} images: {e0.elapsed_time(e1) / NS:.3f} ms per step (includes the per-call conditioning set-up)")
