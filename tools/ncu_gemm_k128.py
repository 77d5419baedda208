"""One launch of the K = 128 GEGLU linear (C -> 8C with the fused activation epilogue, diffusion.py:151-152) inside a
cudaProfilerStart/Stop range:  ncu --profile-from-start off --set full --import-source on -o ... python tools/ncu_gemm_k128.py [M]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
g = torch.Generator(device="cuda").manual_seed(0)
K, N = 128, 1024
a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device=dev, generator=g)
w3 = (torch.randn(384, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
def run():
    ops.gemm(a, w, N, bias=bias, geglu=True)
    ops.gemm(a, w3, 384)
for _ in range(2): run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
