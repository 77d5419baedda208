"""One launch each of the HBM-bound kernels at the benchmark shapes (64x64, C = 128, batch B) inside a
cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_hbm python tools/ncu_hbm_kernels.py
Printed next to each op: its ALGORITHMIC bytes (what `tools/norm_bench.py`, `tail_bench.py`, `head_bench.py` and bench.py
divide by), so that ncu's dram__bytes_read + dram__bytes_write can be put beside them (tools/ncu_traffic.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
n, hw, C, H = B, 4096, 128, 64
M = n * hw
x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
dy = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
scratch = torch.zeros(ops.gn_scratch_floats(n), device=dev)
stats = ops.gn_stats(x, n, hw, 1e-5, scratch)
T = M * C * 2  # bytes of one bf16 activation tensor
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
npar = 30945156
p, gr = torch.randn(npar, device=dev), torch.randn(npar, device=dev)
m, v = torch.zeros(npar, device=dev), torch.zeros(npar, device=dev)
ss = torch.ones(1, device=dev)
# fused sampling tail / tail forward / head forward / tail data gradient (tools/tail_bench.py, tools/head_bench.py)
a2 = torch.randn(2 * n * hw, C, device=dev, generator=g).to(torch.bfloat16)
wt = torch.randn(3, C, 3, 3, device=dev, generator=g) * 0.03
bt = torch.randn(3, device=dev, generator=g) * 0.1
c1 = torch.rand(1000, device=dev) + 0.5
c2 = torch.rand(1000, device=dev) * 0.1
sig = torch.rand(1000, device=dev) * 0.1
step = torch.tensor([500], device=dev, dtype=torch.int32)
nan_flag = torch.zeros(1, device=dev, dtype=torch.int32)
xs = torch.randn(2 * n, 3, H, H, device=dev, generator=g)
img = torch.randn(n, 3, H, H, device=dev, generator=g)
wh = torch.randn(C, 3, 3, 3, device=dev, generator=g) * 0.2
bh = torch.randn(C, device=dev, generator=g) * 0.1
dimg = torch.randn(n, 3, H, H, device=dev, generator=g)
da = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
tail_out = torch.empty(n, 3, H, H, device=dev)

ROWS = [
    ("gn_apply_kernel", "GroupNorm+SiLU+dropout: 1 read + 1 write", 2 * T,
     lambda: ops.gn_apply(x, n, hw, stats, gamma, beta, True, drop_p=0.1, seed=7)),
    ("gn_bwd_sums_kernel", "dy, x read", 2 * T,
     lambda: None),  # launched by gn_bwd below (two kernels of one entry point)
    ("gn_bwd_apply_kernel", "dy, x, residual gradient read, dx written", 4 * T,
     lambda: ops.gn_bwd(dy, x, n, hw, stats, gamma, beta, True, dg, db, radd=dy)),
    ("ln_fwd_kernel", "1 read + 1 write", 2 * T, lambda: ops.ln_fwd(x, gamma, beta)),
    ("ln_bwd_kernel", "dy, x, residual gradient read, dx written", 4 * T,
     lambda: ops.ln_bwd(dy, x, gamma, dg, db, radd=dy)),
    ("adamw_clip_kernel", "p, g, m, v read; p, m, v written (fp32)", npar * 7 * 4,
     lambda: ops.adamw_clip(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-5, 1, 1.0, ss, write_clipped_grad=False)),
    ("tail_conv_y_kernel<3, 1", "fused sampling tail: 2 x bf16 [hw,128] + x_t read, x_{t-1} written twice (CFG pair)",
     2 * n * hw * 256 + 3 * n * 3 * hw * 4,
     lambda: ops.tail_conv_sample(a2, wt, bt, xs, n, H, H, step, c1, c2, sig, 1.8, nan_flag)),
    ("tail_conv_y_kernel<3, 0", "128->3 conv, training forward: bf16 [hw,128] read, fp32 image written",
     n * hw * 256 + n * 3 * hw * 4, lambda: ops.tail_conv_fwd(a2[:M], wt, bt, n, H, H, out=tail_out)),
    ("in_conv_tf32_kernel<3, 64, 0", "3->128 head conv: fp32 image read, bf16 [hw,128] written",
     n * hw * 256 + n * 3 * hw * 4, lambda: ops.head_conv_fwd(img, wh, bh)),
    ("in_conv_tf32_kernel<3, 64, 1", "tail data gradient: fp32 image gradient read, bf16 [hw,128] written",
     n * hw * 256 + n * 3 * hw * 4, lambda: ops.tail_conv_dgrad(dimg, da, wt, n, H, H)),
]


def run():
    for _name, _what, _bytes, fn in ROWS:
        fn()


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(json.dumps({"B": B, "rows": [[nm, what, by] for nm, what, by, _ in ROWS]}))
