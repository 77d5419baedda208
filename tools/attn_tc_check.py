"""tcgen05 attention forward vs torch SDPA fp32 (same bf16 inputs): output / lse error and timing.
usage: attn_tc_check.py B,L,C   (env TSD_ATTN_TC=0/1, TSD_ATTN_TC_POLY=0/2/3/4)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, L, C = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64,4096,128").split(","))
H = 8
dh = C // H
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * L, 3 * C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, B, L, C, H, need_lse=True)
torch.cuda.synchronize()
nb = min(B, 2)
for bi in sorted({0, B - 1}):
    x = qkv[bi * L:(bi + 1) * L].float().view(1, L, 3, H, dh).permute(2, 0, 3, 1, 4).contiguous()
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2])
    ref_o = ref.permute(0, 2, 1, 3).reshape(L, C)
    s = (x[0] @ x[1].transpose(-1, -2)) / dh ** 0.5
    ref_lse2 = torch.logsumexp(s, -1) * 1.4426950408889634  # [1,H,L]
    eo = (out[bi * L:(bi + 1) * L].float() - ref_o).abs().max().item() / ref_o.abs().max().item()
    el = (lse.view(B, H, L)[bi] - ref_lse2[0]).abs().max().item()
    print(f"sample {bi}: rel max err out {eo:.3e}  abs err lse2 {el:.3e}  finite {torch.isfinite(out).all().item()}")


def timeit(fn, n=10):
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


tf = timeit(lambda: ops.attn_fwd(qkv, B, L, C, H, need_lse=True))
nexp = B * H * L * L
print(f"B={B} L={L} C={C} TC={os.environ.get('TSD_ATTN_TC','1')} POLY={os.environ.get("TSD_ATTN_TC_POLY","4")}: "
      f"fwd {tf:.3f} ms  {nexp/tf/1e9:.2f} Texp/s  {4*B*L*L*C/tf/1e9:.0f} TF/s")
