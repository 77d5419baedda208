"""Attention backward vs torch autograd fp32 (same bf16 inputs): dq / dk / dv error and timing.
usage: attn_bwd_check.py B,L,C [time_only]  (env TSD_ATTN_BWD_TC=0/1, TSD_ATTN_BWD_FUSED=0/1)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, L, C = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,1024,128").split(","))
time_only = len(sys.argv) > 2
H = 8
dh = C // H
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * L, 3 * C, device=dev, generator=g) * 1.3).to(torch.bfloat16)
dout = torch.randn(B * L, C, device=dev, generator=g).to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, B, L, C, H, need_lse=True)
d = ops.attn_bwd(qkv, out, dout, lse, B, L, C, H)
torch.cuda.synchronize()
tag = f"TC={os.environ.get('TSD_ATTN_BWD_TC', '1')} FUSED={os.environ.get('TSD_ATTN_BWD_FUSED', '1')}"
if not time_only:
    nb = min(B, 2)
    x = qkv[:nb * L].float().view(nb, L, 3, H, dh).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2])
    ref.backward(dout[:nb * L].float().view(nb, L, H, dh).permute(0, 2, 1, 3))
    ref_d = x.grad.permute(1, 3, 0, 2, 4).reshape(nb * L, 3 * C)
    got = d[:nb * L].float()
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        e = (got[:, sl] - ref_d[:, sl]).abs().max().item() / ref_d[:, sl].abs().max().item()
        print(f"{tag} {name}: rel max err {e:.3e} finite {torch.isfinite(got[:, sl]).all().item()}")
    d2 = ops.attn_bwd(qkv, out, dout, lse, B, L, C, H)
    print(f"{tag} rerun: dk/dv identical {torch.equal(d[:, C:], d2[:, C:])}, dq max diff "
          f"{(d[:, :C].float() - d2[:, :C].float()).abs().max().item():.3e}")


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


tb = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, B, L, C, H))
print(f"{tag} B={B} L={L} C={C}: bwd {tb:.3f} ms  {B * H * L * L / tb / 1e9:.2f} Tscore/s  {3.5 * 4 * B * L * L * C / tb / 1e9:.0f} TF/s")
