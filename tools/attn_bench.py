"""Correctness (vs torch SDPA fp32 on the same bf16 inputs) and timing of tsd_attn_fwd / tsd_attn_bwd."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
cases = [(64, 4096, 128), (64, 1024, 128), (64, 1024, 256), (64, 256, 256), (64, 64, 256), (32, 16, 128), (8, 4, 256)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in sys.argv[1].split(","))]
for B, L, C in cases:
    H = 8
    dh = C // H
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (torch.randn(B * L, 3 * C, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    dout = torch.randn(B * L, C, device=dev, generator=g).to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, L, C, H, need_lse=True)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, L, C, H)
    # reference on a slice of the batch (memory)
    nb = min(B, 2)
    x = qkv[: nb * L].float().view(nb, L, 3, H, dh).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2])
    ref_o = ref.permute(0, 2, 1, 3).reshape(nb * L, C)
    ref.backward(dout[: nb * L].float().view(nb, L, H, dh).permute(0, 2, 1, 3))
    ref_d = x.grad.permute(1, 3, 0, 2, 4).reshape(nb * L, 3 * C)
    eo = (out[: nb * L].float() - ref_o).abs().max().item() / ref_o.abs().max().item()
    ed = (dqkv[: nb * L].float() - ref_d).abs().max().item() / ref_d.abs().max().item()

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
    tb = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, B, L, C, H))
    nexp = B * H * L * L
    print(f"B={B} L={L} C={C} dh={dh}: fwd {tf:.3f} ms ({nexp/tf/1e9:.2f} Texp/s, {4*B*L*L*C/tf/1e9:.0f} TF/s) "
          f"bwd {tb:.3f} ms ({nexp/tb/1e9:.2f} Texp/s per pass-equivalent)  err_o {eo:.2e} err_dqkv {ed:.2e}")
