"""A/B: backward kernels with and without the column-sum by-product (vs the separate colsum pass they replace)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


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
    return e0.elapsed_time(e1) / n * 1e3


n, hw, C = 256, 4096, 128
M = n * hw
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, C, device=dev, generator=g).to(BF)
dy = torch.randn(M, C, device=dev, generator=g).to(BF)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
cs, ct = torch.zeros(n, C, device=dev), torch.zeros(C, device=dev)
print(f"colsum alone: {timeit(lambda: ops.colsum(dy, n, hw, total=ct, out=cs)):.1f} us")
print(f"ln_bwd: plain {timeit(lambda: ops.ln_bwd(dy, x, gamma, dg, db, radd=dy)):.1f} us, with colsum "
      f"{timeit(lambda: ops.ln_bwd(dy, x, gamma, dg, db, radd=dy, rows_per_sample=hw, colsum_out=cs, colsum_total=ct)):.1f} us")
scratch = torch.zeros(ops.gn_scratch_floats(n), device=dev)
stats = ops.gn_stats(x, n, hw, 1e-5, scratch)
print(f"gn_bwd: plain {timeit(lambda: ops.gn_bwd(dy, x, n, hw, stats, gamma, beta, True, dg, db)):.1f} us, with colsum "
      f"{timeit(lambda: ops.gn_bwd(dy, x, n, hw, stats, gamma, beta, True, dg, db, colsum_out=cs, colsum_total=ct)):.1f} us")
H = 512
h8 = torch.randn(M, 2 * H, device=dev, generator=g).to(BF)
dgg = torch.randn(M, H, device=dev, generator=g).to(BF)
dbias = torch.zeros(2 * H, device=dev)
dh8 = torch.empty_like(h8)
print(f"colsum 8C alone: {timeit(lambda: ops.colsum(dh8, n, hw, total=dbias)):.1f} us")
print(f"geglu_bwd: plain {timeit(lambda: ops.geglu_bwd(h8, dgg)):.1f} us, with dbias {timeit(lambda: ops.geglu_bwd(h8, dgg, dbias=dbias)):.1f} us")
