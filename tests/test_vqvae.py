"""VQ-VAE latent codec (03_variational_autoencoder/models.py:135-185, 268-378): oracle vs the reference's own
outputs on the CPU, the CUDA path vs both on the GPU, and the image -> latent -> denoise -> image pipeline."""
import os

import pytest
import torch

from oracle import ref_vqvae as V

G = os.path.join(os.path.dirname(__file__), "golden", "vqvae_3x64.pt")


def _golden():
    f = torch.load(G, weights_only=False)
    sd = V.init_state_dict(f["seed"], **f["cfg"])
    assert abs(V.state_dict_digest(sd) - f["digest"]) / f["digest"] < 1e-9  # same weights as when the fixture was made
    return f, sd


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


# ------------------------------------------------------------------ CPU: the oracle is pinned by the reference's outputs
def test_oracle_matches_reference_outputs():
    f, sd = _golden()
    nh = len(f["cfg"]["hidden_dims"])
    with torch.no_grad():
        lat = V.encode(sd, f["x"], nh)
        assert _rel(lat, f["latents"]) < 2e-6
        zq, loss, idx, _ = V.quantize(sd, f["latents"])
        assert torch.equal(idx, f["indices"]) and torch.equal(zq, f["zq"])
        assert abs(loss.item() - f["vq_loss"].item()) / f["vq_loss"].item() < 1e-6
        assert _rel(V.decode(sd, f["zq"], nh), f["recon"]) < 2e-6


def test_module_mirrors_reference_state_dict():
    from from_ddpm_to_stable_diffusion_b200.vqvae import VQVAE
    f, sd = _golden()
    m = VQVAE(in_channels=3, embedding_dim=4, num_embeddings=128, hidden_dims=[64, 128], img_size=64)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd)
    with pytest.raises(RuntimeError):
        m.encode(torch.zeros(1, 3, 64, 64))  # no CPU fallback


# ------------------------------------------------------------------ GPU
def _module(cuda, sd, cfg):
    from from_ddpm_to_stable_diffusion_b200.vqvae import VQVAE
    m = VQVAE(in_channels=cfg["in_channels"], embedding_dim=cfg["embedding_dim"], num_embeddings=cfg["num_embeddings"],
              hidden_dims=list(cfg["hidden_dims"]), img_size=64)
    m.load_state_dict(sd)
    return m.to(cuda).eval()


@pytest.mark.gpu
def test_encode_decode_vs_reference(cuda):
    """encode and decode against the reference's outputs (bf16 activations, fp32 accumulate: rel-L2 <= 2e-2, the
    tolerance of the denoiser path)."""
    f, sd = _golden()
    m = _module(cuda, sd, f["cfg"])
    lat = m.encode(f["x"].to(cuda))[0]
    assert lat.shape == f["latents"].shape and lat.dtype == torch.float32
    e_enc = _rel(lat, f["latents"])
    rec = m.decode(f["zq"].to(cuda))
    assert rec.shape == f["recon"].shape
    e_dec = _rel(rec, f["recon"])
    print(f"codec rel-L2: encode {e_enc:.3e} decode {e_dec:.3e}")
    assert e_enc < 2e-2 and e_dec < 2e-2
    assert float(rec.abs().max()) <= 1.0


@pytest.mark.gpu
def test_nearest_code_indices_bit_exact(cuda):
    """VectorQuantizer on the reference's fp32 latents: indices identical to the reference's argmin, the quantised
    latents are exactly the codebook rows, vq_loss matches.  A few thousand extra random vectors: any index that differs
    from the fp64 oracle must be a near-tie (distance gap below fp32 resolution)."""
    f, sd = _golden()
    m = _module(cuda, sd, f["cfg"])
    zq, loss, idx = m.quantize(f["latents"].to(cuda))
    assert torch.equal(idx.cpu(), f["indices"])
    assert torch.equal(zq.cpu(), f["zq"])
    assert abs(loss.item() - f["vq_loss"].item()) / f["vq_loss"].item() < 1e-5
    g = torch.Generator().manual_seed(3)
    z = torch.randn(16, 4, 16, 16, generator=g) * 0.4
    _, _, idx2 = m.quantize(z.to(cuda))
    _, _, ref_idx, dist = V.quantize({k: v.double() for k, v in sd.items()}, z.double())
    diff = (idx2.cpu() != ref_idx).nonzero().flatten()
    for i in diff.tolist():
        d = dist[i]
        assert abs(d[idx2[i].item()] - d[ref_idx[i]]) < 1e-6 * max(1.0, float(d.abs().max())), i
    assert diff.numel() <= 4


@pytest.mark.gpu
def test_image_to_latent_to_image_with_denoiser(cuda):
    """BASELINE configs[4] end to end on the GPU: images -> VQ-VAE latents 4x16x16 -> reverse diffusion steps of the
    4-channel denoiser -> decoded images (the data flow of 03_train_with_vae.py:24,69 with the repo's own codec)."""
    from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM
    f, sd = _golden()
    m = _module(cuda, sd, f["cfg"])
    x = f["x"].to(cuda)
    rec, x_in, vq_loss = m(x)
    ref_idx_match = (m.quantize(m.encode(x)[0])[2].cpu() == f["indices"]).float().mean().item()
    print(f"end-to-end: indices equal to the reference's {ref_idx_match:.3f}, recon rel-L2 {_rel(rec, f['recon']):.3e}")
    assert ref_idx_match > 0.9  # bf16 encoder error flips only latents that sit near a cell boundary
    assert _rel(rec, f["recon"]) < 0.15 and torch.isfinite(vq_loss)
    lat = m.quantize(m.encode(x)[0])[0]
    assert lat.shape == (2, 4, 16, 16)
    torch.manual_seed(5)
    den = Diffusion(4, [1, 2, 2, 2], 128, num_class=10).to(cuda).eval()
    sampler = SamplerDDPM(den, 0.0015, 0.0195, 1000, w=1.8).to(cuda)
    z0 = sampler(lat, torch.tensor([1, 2], device=cuda), steps=range(20, -1, -1))  # last 21 reverse steps from the latents
    img = m.decode(z0)
    assert img.shape == (2, 3, 64, 64) and torch.isfinite(img).all() and float(img.abs().max()) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("in_ch,D,K,hidden,img", [(3, 8, 64, [32, 64], 32), (1, 4, 96, [64, 128, 256], 64), (3, 16, 512, [128], 16)])
def test_codec_other_geometries_vs_oracle(cuda, in_ch, D, K, hidden, img):
    """Hidden widths that are not multiples of the GEMM tile (zero-padded channels), one / three down-sampling stages,
    a codebook larger than one shared-memory tile: encode, quantise and decode against the oracle on the same weights."""
    sd = V.init_state_dict(3, in_ch, D, K, hidden)
    cfg = dict(in_channels=in_ch, embedding_dim=D, num_embeddings=K, hidden_dims=hidden)
    m = _module(cuda, sd, cfg)
    g = torch.Generator().manual_seed(img)
    x = torch.randn(3, in_ch, img, img, generator=g).clamp(-2.5, 2.5)
    nh = len(hidden)
    with torch.no_grad():
        lat_ref = V.encode(sd, x, nh)
        zq_ref, loss_ref, idx_ref, dist = V.quantize({k: v.double() for k, v in sd.items()}, lat_ref.double())
        rec_ref = V.decode(sd, zq_ref.float(), nh)
    lat = m.encode(x.to(cuda))[0]
    assert lat.shape == lat_ref.shape and _rel(lat, lat_ref) < 2e-2
    zq, loss, idx = m.quantize(lat_ref.to(cuda))
    diff = (idx.cpu() != idx_ref).nonzero().flatten()
    for i in diff.tolist():  # only near-ties may differ from the fp64 oracle
        d = dist[i]
        assert abs(d[idx[i].item()] - d[idx_ref[i]]) < 1e-5 * max(1.0, float(d.abs().max())), i
    assert diff.numel() <= 2 and abs(loss.item() - loss_ref.item()) / loss_ref.item() < 1e-3
    rec = m.decode(zq_ref.float().to(cuda))
    assert rec.shape == (3, 3, img, img) and _rel(rec, rec_ref) < 2e-2
