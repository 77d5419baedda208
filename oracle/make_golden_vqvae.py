"""Generates tests/golden/vqvae_3x64.pt by running the UNMODIFIED reference VQVAE (03_variational_autoencoder/models.py)
on deterministic weights and inputs.  Run in the build container (needs /root/reference): python oracle/make_golden_vqvae.py
"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/03_variational_autoencoder")
from oracle import ref_vqvae as V  # noqa: E402

CFG = dict(in_channels=3, embedding_dim=4, num_embeddings=128, hidden_dims=[64, 128])  # 3x64x64 -> latent 4x16x16


def main():
    import models as ref_models  # the reference's own module
    sd = V.init_state_dict(7, **CFG)
    ref = ref_models.VQVAE(in_channels=3, embedding_dim=4, num_embeddings=128, hidden_dims=list(CFG["hidden_dims"]), img_size=64)
    assert set(ref.state_dict().keys()) == set(sd.keys()), set(ref.state_dict().keys()) ^ set(sd.keys())
    for k, v in ref.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    ref.load_state_dict(sd)
    ref.eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 64, 64, generator=g).clamp(-2.5, 2.5)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):  # VectorQuantizer.forward prints its tensors
        lat = ref.encode(x)[0]
        zq, vq_loss = ref.vq_layer(lat)
        recon = ref.decode(zq)
        full = ref(x)
    assert torch.equal(full[0], recon)
    E = sd["vq_layer.embedding.weight"]
    # the returned tensor is latents + (codebook[idx] - latents) (models.py:180): recover idx as the row closest to it
    idx = (zq.permute(0, 2, 3, 1).reshape(-1, 1, 4) - E[None]).abs().sum(-1).argmin(1)
    assert float((zq.permute(0, 2, 3, 1).reshape(-1, 4) - E[idx]).abs().max()) < 1e-6
    out = dict(cfg=CFG, seed=7, digest=V.state_dict_digest(sd), x=x, latents=lat, zq=zq, vq_loss=vq_loss, indices=idx,
               recon=recon, n_unique=int(idx.unique().numel()))
    path = os.path.join(ROOT, "tests", "golden", "vqvae_3x64.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; latent std", float(lat.std()), "codes used", out["n_unique"],
          "recon std", float(recon.std()))


if __name__ == "__main__":
    main()
