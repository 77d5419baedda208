import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops
dev = torch.device("cuda:0")
B, L, C, H = 1, 256, 128, 8
key = torch.arange(L, device=dev)
d = torch.arange(16, device=dev)
for name, V in [("key%16==d", (key[:, None] % 16 == d[None, :]).float()), ("key//16==d", (key[:, None] // 16 == d[None, :]).float()),
                ("ones", torch.ones(L, 16, device=dev))]:
    qkv = torch.zeros(B * L, 3 * C, device=dev)
    qkv[:, 2 * C:2 * C + 16] = V
    out, lse = ops.attn_fwd(qkv.to(torch.bfloat16), B, L, C, H, need_lse=True)
    print(name, "row0 x256:", (out[0, :16].float() * 256).tolist(), "lse2", lse.view(B, H, L)[0, 0, 0].item())
    print(name, "row133 x256:", (out[133, :16].float() * 256).tolist())
