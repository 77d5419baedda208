"""Probe the P (TMEM) x V (MN-major smem) pairing of the tcgen05 attention kernel with one-hot softmax rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, L, C, H = 1, 256, 128, 8
# test 1: uniform P, V[key][d] = d + 1
qkv = torch.zeros(B * L, 3 * C, device=dev)
qkv[:, 2 * C:] = (torch.arange(C, device=dev) % 16 + 1).float()[None, :]
out, _ = ops.attn_fwd(qkv.to(torch.bfloat16), B, L, C, H, need_lse=True)
print("uniform P, V[key][d]=d+1 -> out[0][:16] =", out[0, :16].float().tolist())
print("                          out[200][16:32] =", out[200, 16:32].float().tolist())
# test 2: one-hot P at key k1, V[key][0] = key % 64, V[key][1] = key // 64
res = []
for k1 in list(range(0, 64)) + [64, 65, 130, 255]:
    qkv = torch.zeros(B * L, 3 * C, device=dev)
    qkv[:, 0] = 8.0                      # q[.,0] (head 0)
    qkv[:, C] = -8.0                     # k[key,0]
    qkv[k1, C] = 8.0
    qkv[:, 2 * C] = (torch.arange(L, device=dev) % 64).float()
    qkv[:, 2 * C + 1] = (torch.arange(L, device=dev) // 64).float()
    out, _ = ops.attn_fwd(qkv.to(torch.bfloat16), B, L, C, H, need_lse=True)
    res.append((k1, out[5, 0].item(), out[5, 1].item(), out[133, 0].item(), out[133, 1].item()))
for r in res:
    print("k1=%3d -> row5: V idx %.2f blk %.2f | row133: V idx %.2f blk %.2f" % r)
