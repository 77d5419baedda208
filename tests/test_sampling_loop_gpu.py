"""SamplerDDPM.forward (06_tiny_stable_diffusion/utils.py:157-171) as the replayed CUDA graph: the full T = 1000 loop,
fresh noise per call, sharding by global sample index, and robustness of the captured graph against buffer turnover."""
import pytest
import torch

from oracle import ref_unet as R

pytestmark = pytest.mark.gpu
MULTY = [1, 2, 2, 2]
BETAS = (0.0015, 0.0195, 1000)


def _model(cuda, seed=0):
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    sd = R.init_state_dict(seed, 3, MULTY, 128, 3)
    m = Diffusion(3, MULTY, 128, num_class=3, dropout=0.1)
    m.load_state_dict(sd)
    return m.to(cuda).eval()


def _sampler(m, cuda, w=1.8, graph=True):
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM
    s = SamplerDDPM(m, *BETAS, w=w).to(cuda)
    s.rng.load_state_dict({"seed": 777, "calls": 0})
    s.use_cuda_graph = graph
    return s


def _rewind(s):
    s.rng.load_state_dict({"seed": 777, "calls": 0})


def test_full_T1000_loop_graph_equals_eager(cuda):
    """One complete reverse process (T = 1000, CFG w = 1.8, batch 8, 3x64x64) through the graph: finite, NaN flag
    clear, clipped to [-1, 1]; and the graph equals eager stepping after 2, 500, 999 and all 1000 steps."""
    m = _model(cuda)
    g = torch.Generator().manual_seed(5)
    xT = torch.randn(8, 3, 64, 64, generator=g).to(cuda)
    y = torch.randint(1, 4, (8,), generator=g).to(cuda)
    sg, se = _sampler(m, cuda), _sampler(m, cuda, graph=False)
    for last in (998, 500, 1, 0):
        steps = range(999, last - 1, -1)
        _rewind(sg), _rewind(se)
        a = sg(xT, y, steps=steps)
        b = se(xT, y, steps=steps)
        assert torch.isfinite(a).all() and int(sg._plan.nan_flag.item()) == 0
        assert torch.equal(a, b), (last, (a - b).abs().max().item())
    assert float(a.abs().max()) <= 1.0  # utils.py:171
    assert float(a.std()) > 1e-3


def test_every_call_draws_a_new_noise_sequence(cuda):
    """utils.py:163 draws fresh randn_like per step and per call: two calls with the same x_T differ, while the same
    stream position reproduces the result; w and seed changes reach the captured graph."""
    m = _model(cuda)
    g = torch.Generator().manual_seed(6)
    xT = torch.randn(2, 3, 32, 32, generator=g).to(cuda)
    y = torch.tensor([1, 2], device=cuda)
    steps = range(999, 993, -1)
    s = _sampler(m, cuda)
    a1 = s(xT, y, steps=steps)
    a2 = s(xT, y, steps=steps)
    assert not torch.equal(a1, a2) and s.rng.calls == 2
    _rewind(s)
    assert torch.equal(s(xT, y, steps=steps), a1)
    # guidance weight baked into the captured launch: changing it must re-capture
    _rewind(s)
    s.w = 0.0
    b = s(xT, y, steps=steps)
    ref = _sampler(m, cuda, w=0.0, graph=False)
    assert torch.equal(b, ref(xT, y, steps=steps)) and not torch.equal(b, a1)
    _rewind(s)
    s.seed = 778
    assert not torch.equal(s(xT, y, steps=steps), b)


def test_shards_draw_by_global_sample_index(cuda):
    """a rank that samples images [4, 8) of a batch gets exactly what one process computes for them."""
    from from_ddpm_to_stable_diffusion_b200.parallel import set_shard
    m = _model(cuda)
    g = torch.Generator().manual_seed(7)
    xT = torch.randn(8, 3, 32, 32, generator=g).to(cuda)
    y = torch.randint(1, 4, (8,), generator=g).to(cuda)
    steps = range(999, 989, -1)
    full = _sampler(m, cuda)(xT, y, steps=steps)
    part = _sampler(m, cuda)
    set_shard(part, 4)
    got = part(xT[4:], y[4:], steps=steps)
    assert torch.equal(got, full[4:]), (got - full[4:]).abs().max().item()


def test_captured_graph_survives_buffer_turnover(cuda):
    """The graph bakes pointers to the engine's packed weights and GroupNorm scratch.  A grad-enabled forward (training
    packing) and a larger no_grad batch (scratch re-allocation) between two sampler calls must not leave the graph
    reading freed memory: the second call equals a fresh sampler at the same stream position."""
    m = _model(cuda)
    g = torch.Generator().manual_seed(8)
    xT = torch.randn(2, 3, 32, 32, generator=g).to(cuda)
    y = torch.tensor([3, 1], device=cuda)
    steps = range(999, 991, -1)
    s = _sampler(m, cuda)
    first = s(xT, y, steps=steps)
    # turnover 1: training-mode forward + backward (packs the plain linear_1 weights, allocates gradient buffers)
    m.train()
    xb = torch.randn(4, 3, 32, 32, device=cuda)
    out = m(xb, torch.randint(0, 1000, (4,), device=cuda), torch.randint(0, 4, (4,), device=cuda))
    out.square().mean().backward()
    m.eval()
    # turnover 2: a larger inference batch re-allocates the shared GroupNorm scratch
    with torch.no_grad():
        m(torch.randn(64, 3, 32, 32, device=cuda), torch.zeros(64, dtype=torch.long, device=cuda),
          torch.zeros(64, dtype=torch.long, device=cuda))
    junk = [torch.randn(1 << 20, device=cuda) for _ in range(8)]  # recycle whatever the caches released
    _rewind(s)
    again = s(xT, y, steps=steps)
    del junk
    assert torch.equal(again, first)
    fresh = _sampler(m, cuda, graph=False)(xT, y, steps=steps)
    assert torch.equal(again, fresh)
