"""Per-call-shape device time of one training step and one reverse step (TSD_PROFILE=1, CUDA events)."""
import os
import sys

os.environ["TSD_PROFILE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM, _lib  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda:0")
torch.manual_seed(0)
latent = mode == "latent"  # BASELINE configs[4]: 4x16x16 latents, 10 classes, sampling
model = Diffusion(4 if latent else 3, [1, 2, 2, 2], 128, num_class=10 if latent else 3, dropout=0.1).to(dev).train()
x = torch.randn(B, 4, 16, 16, device=dev) if latent else torch.randn(B, 3, 64, 64, device=dev)
y = torch.randint(1, 11 if latent else 4, (B,), device=dev)


def flops(name, k):
    if name == "tsd_gemm_fwd":
        c0, c1, M, N = k[0], k[1], k[2], k[3]
        return 2.0 * M * N * (c0 + c1)
    if name == "tsd_conv3x3_fwd":
        c0, c1, n, H, W, s, cout = k[:7]
        return 2.0 * n * (H // s) * (W // s) * cout * 9 * (c0 + c1)
    if name == "tsd_gemm_dgrad":
        M, N, K = k[:3]
        return 2.0 * M * N * K
    if name == "tsd_conv3x3_dgrad":
        n, H, W, cout, cin = k[:5]
        return 2.0 * n * H * W * cout * 9 * cin
    if name == "tsd_gemm_wgrad":
        c0, c1, M, N = k[:4]
        return 2.0 * M * N * (c0 + c1)
    if name == "tsd_conv3x3_wgrad":
        c0, c1, n, H, W, s, cout = k[:7]
        return 2.0 * n * (H // s) * (W // s) * cout * 9 * (c0 + c1)
    if name in ("tsd_attn_fwd", "tsd_attn_fwd_ws", "tsd_attn_bwd", "tsd_attn_bwd_ws"):
        Bn, L, C, heads = k[:4]
        f = 4.0 * Bn * L * L * C
        return f if name.startswith("tsd_attn_fwd") else 3.5 * f
    return 0.0


if mode == "train":
    trainer = TrainerDDPM(model, 0.0015, 0.0195, 1000).to(dev)
    opt = FusedClipAdamW(model, lr=2e-6, weight_decay=1e-5, max_norm=1.0)

    def step():
        opt.zero_grad()
        loss = trainer(x, y).sum() / B ** 2
        loss.backward()
        opt.step()
else:
    model.eval()
    sampler = SamplerDDPM(model, 0.0015, 0.0195, 1000, w=1.8).to(dev)
    sampler.use_cuda_graph = False

    def step():
        sampler(x, y, steps=[500])

for _ in range(2):
    step()
_lib.profile_report()
step()
agg = _lib.profile_report()
tot = sum(v[1] for v in agg.values())
print(f"mode={mode} B={B} total {tot:.2f} ms in {sum(v[0] for v in agg.values())} calls")
byname = {}
for (name, key), (n, t) in agg.items():
    a = byname.setdefault(name, [0, 0.0, 0.0])
    a[0] += n; a[1] += t; a[2] += flops(name, key) * n
print("--- by entry point")
for name, (n, t, f) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} {n:5d} {t:9.3f} ms {100*t/tot:5.1f}%  {f/t/1e9 if f else 0:8.1f} TF/s")
print("--- top shapes")
for (name, key), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    f = flops(name, key) * n
    print(f"{name:22s} {str(key):52s} x{n:<3d} {t:8.3f} ms {100*t/tot:5.1f}% {f/t/1e9 if f else 0:8.1f} TF/s")
