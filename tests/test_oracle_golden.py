"""Pins the oracle (oracle/ref_unet.py) to the reference's own outputs stored in tests/golden/ (made by
oracle/make_golden.py, which imports the unmodified reference).  CPU only."""
import json
import os

import pytest
import torch

from oracle import ref_unet as R

G = os.path.join(os.path.dirname(__file__), "golden")
MULTY = [1, 2, 2, 2]
B1, BT, T = 0.0015, 0.0195, 1000


def _load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


@pytest.fixture(scope="module")
def sd3():
    return R.init_state_dict(0, 3, MULTY, 128, 3)


def test_state_dict_contract():
    keys = json.load(open(os.path.join(G, "state_dict_keys.json")))
    shapes = R.param_shapes(3, MULTY, 128, 3)
    assert len(keys) == 425
    assert [k for k, _ in keys] == list(shapes.keys())
    assert all(tuple(s) == shapes[k] for k, s in keys)


def test_weights_regenerate_identically(sd3):
    f = _load("fwd_3x64.pt")
    assert R.state_dict_digest(sd3) == f["digest"]


def test_forward_3x64(sd3):
    f = _load("fwd_3x64.pt")
    with torch.no_grad():
        eps = R.unet_forward(sd3, f["x"], f["t"], f["y"], MULTY)
    assert (eps - f["eps"]).abs().max().item() < 2e-5
    with torch.no_grad():
        eps2 = R.unet_forward(sd3, f["x"], f["t"], f["y"], MULTY, use_sdpa=True)
    assert (eps2 - f["eps"]).abs().max().item() < 2e-5


def test_forward_latent_4x16():
    f = _load("fwd_4x16.pt")
    sd = R.init_state_dict(1, 4, MULTY, 128, 10)
    assert R.state_dict_digest(sd) == f["digest"]
    with torch.no_grad():
        eps = R.unet_forward(sd, f["x"], f["t"], f["y"], MULTY)
    assert (eps - f["eps"]).abs().max().item() < 2e-5


def test_schedule_tables_bit_exact():
    tab = _load("schedule.pt")
    s = R.make_schedule(B1, BT, T)
    for k in ("betas", "sqrt_alphas_bar", "sqrt_one_minus_alphas_bar", "coeff1", "coeff2", "posterior_var"):
        assert s[k].dtype == torch.float64
        assert torch.equal(s[k], tab[k]), k
    # known answers quoted in SURVEY 8(a12, a14)
    assert s["sqrt_alphas_bar"][0].item() == 0.9992497185323403
    assert s["sqrt_alphas_bar"][999].item() == 0.005068729615167281
    assert s["sqrt_one_minus_alphas_bar"][0].item() == 0.03872983363040069
    assert s["coeff1"][0].item() == 1.0007508448126077
    assert s["coeff1"][999].item() == 1.0098949513484736
    assert s["coeff2"][0].item() == 0.03875891372507523
    assert s["coeff2"][999].item() == 0.019693204938339242
    var = R.sampler_variance(s)
    assert var[0].item() == 0.0007550472963652511 and var[999].item() == 0.019500000402331352
    tt = torch.tensor([0, 1, 500, 999])
    assert torch.equal(R.extract(s["sqrt_alphas_bar"], tt, (4, 3, 8, 8)), tab["extract_sqrt_alphas_bar"])
    assert R.extract(s["sqrt_alphas_bar"], tt, (4, 3, 8, 8)).shape == (4, 1, 1, 1)


def test_trainer_loss_and_grads(sd3):
    f = _load("trainer_3x32.pt")
    sched = R.make_schedule(B1, BT, T)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd3.items()}
    loss = R.trainer_loss(sdg, sched, f["x0"], f["labels"], f["t"], f["noise"], MULTY)
    assert loss.shape == f["loss"].shape
    assert (loss - f["loss"]).abs().max().item() < 1e-4
    (loss.sum() / 2 ** 2).backward()
    for k, ref in f["grads"].items():
        got = sdg[k].grad
        assert (got - ref).abs().max().item() <= 1e-4 * (ref.abs().max().item() + 1e-6) + 1e-7, k
    worst = 0.0
    for k, n in f["grad_norms"].items():
        gn = float(sdg[k].grad.norm())
        if n > 1e-6:
            worst = max(worst, abs(gn - n) / n)
        else:  # the 40 mathematically dead cross-attention tensors
            assert gn <= 1e-6, k
    assert worst < 1e-3, worst
    # padding row of the label embedding gets no gradient (diffusion.py:197)
    assert sdg["label_embedding.0.weight"].grad[0].abs().max().item() == 0.0


def test_sampler_loop_and_single_steps(sd3):
    f = _load("sampler_3x32.pt")
    w = f["w"]
    # the reference's full loop on a 4-step schedule, replayed with its own noise draws
    s4 = R.make_schedule(B1, BT, 4)
    x = f["T4"]["x_T"]
    zs = f["T4"]["zs"]
    with torch.no_grad():
        for i, ts in enumerate(reversed(range(4))):
            x, _, _ = R.sampler_step(sd3, s4, x, f["T4"]["labels"], ts, zs[i] if ts > 0 else None, w, MULTY)
    assert (x.clip(-1, 1) - f["T4"]["x_0"]).abs().max().item() < 1e-4
    sched = R.make_schedule(B1, BT, T)
    for s in f["singles"]:
        with torch.no_grad():
            got, _, _ = R.sampler_step(sd3, sched, s["x_t"], f["labels"], s["t"], s["z"], w, MULTY)
        assert (got - s["x_prev"]).abs().max().item() < 1e-4, s["t"]
