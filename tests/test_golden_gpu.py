"""The CUDA path against the REFERENCE's own outputs (tests/golden, made by oracle/make_golden.py)."""
import os

import pytest
import torch

from oracle import ref_unet as R

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
MULTY = [1, 2, 2, 2]


def _load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def _model(cuda, channel_img, num_class, seed):
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    sd = R.init_state_dict(seed, channel_img, MULTY, 128, num_class)
    m = Diffusion(channel_img, MULTY, 128, num_class=num_class, dropout=0.1)
    m.load_state_dict(sd)
    return m.to(cuda).eval()


def _rel(a, b):
    return ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("name,ci,nc,seed", [("fwd_3x64.pt", 3, 3, 0), ("fwd_4x16.pt", 4, 10, 1)])
def test_forward_vs_reference(cuda, name, ci, nc, seed):
    f = _load(name)
    m = _model(cuda, ci, nc, seed)
    with torch.no_grad():
        eps = m(f["x"].to(cuda), f["t"].to(cuda), f["y"].to(cuda))
    assert eps.shape == f["eps"].shape and eps.dtype == torch.float32
    assert _rel(eps, f["eps"]) < 2e-2  # bf16 activations, fp32 accumulate


def test_trainer_vs_reference(cuda):
    from from_ddpm_to_stable_diffusion_b200 import TrainerDDPM
    f = _load("trainer_3x32.pt")
    m = _model(cuda, 3, 3, 0)  # eval mode like the golden run: dropout off
    tr = TrainerDDPM(m, 0.0015, 0.0195, 1000).to(cuda)
    loss = tr(f["x0"].to(cuda), f["labels"].to(cuda), t=f["t"].to(cuda), noise=f["noise"].to(cuda))
    assert loss.shape == f["loss"].shape
    assert abs(loss.sum().item() - f["loss"].sum().item()) / f["loss"].sum().item() < 1e-3  # 0.1 %
    (loss.sum() / 2 ** 2).backward()
    P = dict(m.named_parameters())
    for k, ref in f["grads"].items():
        got = P[k].grad.float().cpu()
        cos = (got.flatten() @ ref.flatten() / (got.norm() * ref.norm() + 1e-30)).item()
        assert cos > 0.995, (k, cos)
    n_ref = sum(v ** 2 for v in f["grad_norms"].values()) ** 0.5
    n_got = sum(float(p.grad.double().pow(2).sum()) for p in P.values()) ** 0.5
    assert abs(n_got - n_ref) / n_ref < 1e-3


def test_sampler_vs_reference(cuda):
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM
    f = _load("sampler_3x32.pt")
    m = _model(cuda, 3, 3, 0)
    w = f["w"]
    s1000 = SamplerDDPM(m, 0.0015, 0.0195, 1000, w=w).to(cuda)
    sched = R.make_schedule(0.0015, 0.0195, 1000)
    for s in f["singles"]:
        got = s1000(s["x_t"].to(cuda), f["labels"].to(cuda), steps=[s["t"]], noise_fn=lambda ts: s["z"].to(cuda))
        ref = s["x_prev"].clip(-1, 1) if s["t"] == 0 else s["x_prev"]
        tol = float(sched["coeff2"][s["t"]]) * (1 + 2 * w) * 5e-2 + 1e-5 * float(s["x_t"].abs().max())
        assert (got.cpu() - ref).abs().max().item() <= tol, s["t"]
    # the reference's whole loop (4-step schedule): noise replayed, t = 0 branch and final clip included
    s4 = SamplerDDPM(m, 0.0015, 0.0195, 4, w=w).to(cuda)
    zs = {3: f["T4"]["zs"][0], 2: f["T4"]["zs"][1], 1: f["T4"]["zs"][2], 0: f["T4"]["zs"][0]}
    got = s4(f["T4"]["x_T"].to(cuda), f["T4"]["labels"].to(cuda), noise_fn=lambda ts: zs[ts].to(cuda))
    assert got.abs().max().item() <= 1.0
    assert (got.cpu() - f["T4"]["x_0"]).abs().max().item() < 5e-2
