"""One launch each of the three dominant kernels at the benchmark shapes inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_kernels python tools/ncu_kernels.py
(gemm_tc_kernel: conv3x3 128->128 @64x64 forward (halo mode) and weight gradient (patch mode), batch 256; attn_fwd_tc_kernel / attn_bwd_tc_kernel: L=4096, head_dim 16, batch B_ATT; attn_bwd_tc32_kernel: L=1024, head_dim 32, batch B)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B_ATT = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator(device="cuda").manual_seed(0)
H, C = 64, 128
x = torch.randn(B * H * H, C, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(C, 9 * C, device=dev, generator=g) * 0.03).to(torch.bfloat16)
bias = torch.zeros(C, device=dev)
dyc = torch.randn(B * H * H, C, device=dev, generator=g).to(torch.bfloat16)
dw = torch.zeros(C, 9 * C, device=dev)
L = 4096
qkv = (torch.randn(B_ATT * L, 3 * C, device=dev, generator=g) * 1.3).to(torch.bfloat16)
dout = torch.randn(B_ATT * L, C, device=dev, generator=g).to(torch.bfloat16)
L32, C32 = 1024, 256  # head_dim 32 (the 32x32 stage): attn_bwd_tc32_kernel
qkv32 = (torch.randn(B * L32, 3 * C32, device=dev, generator=g) * 1.3).to(torch.bfloat16)
dout32 = torch.randn(B * L32, C32, device=dev, generator=g).to(torch.bfloat16)


def run():
    ops.conv3x3(x, B, H, H, w, C, bias=bias)       # gemm_tc_kernel<0,0,0>, halo mode
    ops.conv3x3_wgrad(dyc, x, B, H, H, dw)         # gemm_tc_kernel<1,1,1>, patch mode
    out, lse = ops.attn_fwd(qkv, B_ATT, L, C, 8, need_lse=True)
    ops.attn_bwd(qkv, out, dout, lse, B_ATT, L, C, 8)
    out32, lse32 = ops.attn_fwd(qkv32, B, L32, C32, 8, need_lse=True)
    ops.attn_bwd(qkv32, out32, dout32, lse32, B, L32, C32, 8)  # attn_bwd_tc32_kernel, batch B


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
