import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_unet as R
from from_ddpm_to_stable_diffusion_b200 import Diffusion
dev = torch.device("cuda:0")
sd = R.init_state_dict(0, 3, [1, 2, 2, 2], 128, 3)
m = Diffusion(3, [1, 2, 2, 2], 128, num_class=3)
m.load_state_dict(sd)
m = m.to(dev).eval()
for S in (32, 64):
    x = torch.randn(2, 3, S, S, device=dev)
    t = torch.tensor([5, 600], device=dev)
    y = torch.tensor([1, 0], device=dev)
    with torch.no_grad():
        outs = []
        tapsl = []
        for i in range(3):
            taps = {}
            eps, _ = m._engine.forward(x, t, y, save=False, taps=taps)
            outs.append(eps.clone())
            tapsl.append({k: v[0].clone() for k, v in taps.items()})
    print(S, "eps diffs", [(outs[0] - o).abs().max().item() for o in outs[1:]])
    for k in tapsl[0]:
        d = (tapsl[0][k].float() - tapsl[1][k].float()).abs().max().item()
        if d > 0:
            print("  first differing block:", k, d, "scale", tapsl[0][k].float().abs().max().item())
            break
