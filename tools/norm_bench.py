"""HBM-bound kernels: achieved GB/s (algorithmic bytes / CUDA-event time) at the 64x64, C=128 stage."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
peak = 6548.8
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
n, hw, C = B, 4096, 128
M = n * hw
x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
dy = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
scratch = torch.zeros((8 * 160 + n) * 64 + 64, device=dev)
stats = ops.gn_stats(x, n, hw, 1e-5, scratch)
T = M * C * 2  # bytes of one bf16 tensor
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
h8 = torch.randn(M // 4, 8 * C, device=dev, generator=g).to(torch.bfloat16)
dgg = torch.randn(M // 4, 4 * C, device=dev, generator=g).to(torch.bfloat16)
rows = [
    ("gn_stats (1 read)", lambda: ops.gn_stats(x, n, hw, 1e-5, scratch), T),
    ("gn_apply+SiLU (1 read + 1 write)", lambda: ops.gn_apply(x, n, hw, stats, gamma, beta, True), 2 * T),
    ("gn_apply+SiLU+dropout", lambda: ops.gn_apply(x, n, hw, stats, gamma, beta, True, drop_p=0.1, seed=7), 2 * T),
    ("gn_bwd (2x(dy,x) reads + radd + write)", lambda: ops.gn_bwd(dy, x, n, hw, stats, gamma, beta, True, dg, db, radd=dy), 6 * T),
    ("ln_fwd", lambda: ops.ln_fwd(x, gamma, beta), 2 * T),
    ("ln_bwd (dy, x, radd -> dx)", lambda: ops.ln_bwd(dy, x, gamma, dg, db, radd=dy), 4 * T),
    ("geglu_fwd", lambda: ops.geglu_fwd(h8), h8.numel() * 2 + dgg.numel() * 2),
    ("geglu_bwd", lambda: ops.geglu_bwd(h8, dgg), 2 * h8.numel() * 2 + dgg.numel() * 2),
    ("add_bf16", lambda: ops.add(x, dy), 3 * T),
    ("colsum per sample", lambda: ops.colsum(x, n, hw), T),
]
npar = 30945156
p = torch.randn(npar, device=dev); gr = torch.randn(npar, device=dev); m = torch.zeros(npar, device=dev); v = torch.zeros(npar, device=dev)
ss = torch.ones(1, device=dev)
rows.append(("adamw_clip (30.9M params)", lambda: ops.adamw_clip(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-5, 1, 1.0, ss), npar * 7 * 4))
xs = torch.randn(2 * B, 3, 64, 64, device=dev); eps = torch.randn(2 * B, 3, 64, 64, device=dev)
step = torch.full((1,), 500, device=dev, dtype=torch.int32); flag = torch.zeros(1, device=dev, dtype=torch.int32)
tab = torch.rand(1000, device=dev)
rows.append(("sampler_update (x, 2 eps -> 2 x)", lambda: ops.sampler_update(xs, eps, step, tab, tab, tab, 1.8, xs, flag, seed=1, dup=True),
             B * 3 * 4096 * 4 * 5))
for name, fn, byts in rows:
    t = timeit(fn)
    print(f"{name:42s} {t*1e3:8.1f} us  {byts/t/1e6:7.0f} GB/s  {byts/t/1e6/peak*100:5.1f}% of measured HBM copy peak")
