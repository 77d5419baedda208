"""Host-side logic of the drop-in shells and the C-ABI surface (CPU only, no compute calls)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_state_dict_keys_match_reference():
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    keys = json.load(open(os.path.join(G, "state_dict_keys.json")))
    m = Diffusion(3, [1, 2, 2, 2], 128, num_class=3, dropout=0.1)
    sd = m.state_dict()
    assert [k for k, _ in keys] == list(sd.keys())
    assert all(list(sd[k].shape) == s for k, s in keys)
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert sum(v.numel() for v in sd.values()) == 30945155
    assert torch.count_nonzero(sd["label_embedding.0.weight"][0]) == 0  # padding_idx row


def test_constructor_guards_and_no_cpu_fallback():
    from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM
    with pytest.raises(AssertionError):
        Diffusion(3, [1, 2, 2], 128)
    m = Diffusion(3, [1, 2, 2, 2], 128, num_class=3)
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        TrainerDDPM(m, 0.0015, 0.0195, 1000)(x, torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        SamplerDDPM(m, 0.0015, 0.0195, 1000, w=1.8)(x, torch.zeros(1, dtype=torch.long))


def test_schedule_buffers_bit_exact_vs_reference():
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM, TrainerDDPM, extract
    tab = torch.load(os.path.join(G, "schedule.pt"), weights_only=False)
    tr = TrainerDDPM(torch.nn.Identity(), 0.0015, 0.0195, 1000)
    sa = SamplerDDPM(torch.nn.Identity(), 0.0015, 0.0195, 1000, w=1.8)
    assert [k for k, _ in tr.named_buffers()] == ["betas", "sqrt_alphas_bar", "sqrt_one_minus_alphas_bar"]
    assert [k for k, _ in sa.named_buffers()] == ["betas", "coeff1", "coeff2", "posterior_var"]
    for mod in (tr, sa):
        for k, v in mod.named_buffers():
            assert v.dtype == torch.float64 and torch.equal(v, tab[k]), k
    tt = torch.tensor([0, 1, 500, 999])
    assert torch.equal(extract(tr.sqrt_alphas_bar, tt, (4, 3, 8, 8)), tab["extract_sqrt_alphas_bar"])
    # the fp32 device tables are exactly what extract() would return element by element
    c1, c2, sigma = sa._f32_tables(torch.device("cpu"))
    allt = torch.arange(1000)
    assert torch.equal(c1, extract(sa.coeff1, allt, (1000,)).view(-1))
    assert torch.equal(c2, extract(sa.coeff2, allt, (1000,)).view(-1))
    var = torch.cat([sa.posterior_var[1:2], sa.betas[1:]])
    assert torch.equal(sigma, torch.sqrt(extract(var, allt, (1000,)).view(-1)))
    s0, s1 = tr._f32_tables(torch.device("cpu"))
    assert torch.equal(s0, extract(tr.sqrt_alphas_bar, allt, (1000,)).view(-1))
    assert torch.equal(s1, extract(tr.sqrt_one_minus_alphas_bar, allt, (1000,)).view(-1))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tinysd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from from_ddpm_to_stable_diffusion_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tinysd_b200.h but not exported"
    assert lib.tsd_abi_version() == 2


def test_every_exported_symbol_is_declared():
    import subprocess
    from from_ddpm_to_stable_diffusion_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (tsd_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == _declared_symbols()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "from_ddpm_to_stable_diffusion_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), f"{f} mentions the oracle"
