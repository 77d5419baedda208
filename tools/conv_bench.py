"""conv3x3 forward / dgrad / wgrad timings at the 64x64 C=128 stage (B images)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from from_ddpm_to_stable_diffusion_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(0)
for (H, cin, cout) in [(64, 128, 128), (64, 256, 128), (32, 256, 256), (16, 256, 256), (8, 256, 256)]:
    M = B * H * H
    x = torch.randn(M, cin, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(M, cout, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=dev, generator=g) * 0.03).to(torch.bfloat16)
    wT = (torch.randn(cin, 9 * cout, device=dev, generator=g) * 0.03).to(torch.bfloat16)
    dw = torch.zeros(cout, 9 * cin, device=dev)
    fl = 2.0 * M * cout * 9 * cin
    tf = timeit(lambda: ops.conv3x3(x, B, H, H, w, cout))
    td = timeit(lambda: ops.conv3x3_dgrad(dy, B, H, H, w, cin))
    td2 = timeit(lambda: ops.conv3x3(dy, B, H, H, wT, cin))
    tw = timeit(lambda: ops.conv3x3_wgrad(dy, x, B, H, H, dw))
    print(f"H={H} {cin}->{cout}: fwd {tf*1e3:7.1f} us {fl/tf/1e9:6.0f} TF/s | dgrad(MN-major W) {td*1e3:7.1f} us {fl/td/1e9:6.0f} | "
          f"dgrad as fwd(K-major W^T) {td2*1e3:7.1f} us {fl/td2/1e9:6.0f} | wgrad {tw*1e3:7.1f} us {fl/tw/1e9:6.0f}")
