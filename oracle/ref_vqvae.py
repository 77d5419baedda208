"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not measured, never imported by the product.

CPU (torch fp32 / fp64) functional restatement of the reference's VQ-VAE latent codec, written against the reference's
state_dict key names:

  * VectorQuantizer.forward   03_variational_autoencoder/models.py:149-185
  * ResidualLayer.forward     03_variational_autoencoder/models.py:186-201
  * VQVAE.encode / decode     03_variational_autoencoder/models.py:345-363 (module lists built at :281-343)

Parity status: PINNED.  oracle/make_golden_vqvae.py imports the real reference module from /root/reference, loads the
weights produced by init_state_dict() into it and stores the reference's own latents / indices / reconstruction in
tests/golden/vqvae_3x64.pt; tests/test_oracle_golden.py checks this file against them.
"""
from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
N_RES = 6


def param_shapes(in_channels: int, embedding_dim: int, num_embeddings: int, hidden_dims: List[int]):
    """state_dict keys -> shapes, in the reference's registration order (models.py:281-343)."""
    S = {}
    ci = in_channels
    k = 0
    for h in hidden_dims:
        S[f"encoder.{k}.0.weight"], S[f"encoder.{k}.0.bias"] = (h, ci, 4, 4), (h,)
        ci, k = h, k + 1
    S[f"encoder.{k}.0.weight"], S[f"encoder.{k}.0.bias"] = (ci, ci, 3, 3), (ci,)
    k += 1
    for _ in range(N_RES):
        S[f"encoder.{k}.conv.0.weight"], S[f"encoder.{k}.conv.2.weight"] = (ci, ci, 3, 3), (ci, ci, 1, 1)
        k += 1
    k += 1  # LeakyReLU
    S[f"encoder.{k}.0.weight"], S[f"encoder.{k}.0.bias"] = (embedding_dim, ci, 1, 1), (embedding_dim,)
    S["vq_layer.embedding.weight"] = (num_embeddings, embedding_dim)
    top = hidden_dims[-1]
    S["decoder.0.0.weight"], S["decoder.0.0.bias"] = (top, embedding_dim, 3, 3), (top,)
    k = 1
    for _ in range(N_RES):
        S[f"decoder.{k}.conv.0.weight"], S[f"decoder.{k}.conv.2.weight"] = (top, top, 3, 3), (top, top, 1, 1)
        k += 1
    k += 1
    rev = hidden_dims[::-1]
    for i in range(len(rev) - 1):
        S[f"decoder.{k}.0.weight"], S[f"decoder.{k}.0.bias"] = (rev[i], rev[i + 1], 4, 4), (rev[i + 1],)
        k += 1
    S[f"decoder.{k}.0.weight"], S[f"decoder.{k}.0.bias"] = (rev[-1], 3, 4, 4), (3,)
    return S


def init_state_dict(seed: int, in_channels: int, embedding_dim: int, num_embeddings: int, hidden_dims: List[int]):
    """Deterministic weights, regenerable anywhere (no checkpoint in the repo): Kaiming-like uniform scales so that the
    activations stay O(1) through the 15 + 15 layers, and a codebook spread over the range of the encoder's latents
    (the reference's default U(-1/K, 1/K) collapses every latent of a random-init encoder onto near-ties)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(in_channels, embedding_dim, num_embeddings, hidden_dims).items():
        if k == "vq_layer.embedding.weight":
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.5
        elif k.endswith(".bias"):
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            if k.startswith("decoder") and len(shp) == 4 and shp[2] == 4:  # ConvTranspose2d: [in][out][4][4], 4 taps per output
                fan_in = shp[0] * 4
            bound = (3.0 / fan_in) ** 0.5 * (0.5 if ".conv.2." in k else 1.4)
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return sd


def state_dict_digest(sd: Dict[str, Tensor]) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def _res_stack(sd, prefix, k, x):
    for _ in range(N_RES):  # ResidualLayer: conv3x3 (no bias) -> ReLU -> conv1x1 (no bias), + x  (models.py:189-201)
        h = F.conv2d(x, sd[f"{prefix}.{k}.conv.0.weight"], None, padding=1)
        h = F.conv2d(F.relu(h), sd[f"{prefix}.{k}.conv.2.weight"], None)
        x = h + x
        k += 1
    return F.leaky_relu(x), k + 1


def encode(sd, x: Tensor, n_hidden: int) -> Tensor:
    """VQVAE.encode (models.py:345-353): [N, C, H, W] -> [N, D, H / 2^n_hidden, W / 2^n_hidden]"""
    k = 0
    for _ in range(n_hidden):
        x = F.leaky_relu(F.conv2d(x, sd[f"encoder.{k}.0.weight"], sd[f"encoder.{k}.0.bias"], stride=2, padding=1))
        k += 1
    x = F.leaky_relu(F.conv2d(x, sd[f"encoder.{k}.0.weight"], sd[f"encoder.{k}.0.bias"], padding=1))
    x, k = _res_stack(sd, "encoder", k + 1, x)
    return F.leaky_relu(F.conv2d(x, sd[f"encoder.{k}.0.weight"], sd[f"encoder.{k}.0.bias"]))


def quantize(sd, latents: Tensor, beta: float = 0.25):
    """VectorQuantizer.forward (models.py:149-185) -> (quantised [N, D, h, w], vq_loss, indices [N*h*w], distances)"""
    E = sd["vq_layer.embedding.weight"]
    lat = latents.permute(0, 2, 3, 1).contiguous()
    flat = lat.view(-1, E.shape[1])
    dist = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(flat, E.t())
    idx = torch.argmin(dist, dim=1)
    q = E[idx].view(lat.shape)
    vq_loss = F.mse_loss(q, lat) * beta + F.mse_loss(q, lat)
    q = lat + (q - lat)  # the straight-through form the reference returns (:180): equal to q only up to fp rounding
    return q.permute(0, 3, 1, 2).contiguous(), vq_loss, idx, dist


def decode(sd, z: Tensor, n_hidden: int) -> Tensor:
    """VQVAE.decode (models.py:355-363)"""
    x = F.leaky_relu(F.conv2d(z, sd["decoder.0.0.weight"], sd["decoder.0.0.bias"], padding=1))
    x, k = _res_stack(sd, "decoder", 1, x)
    for i in range(n_hidden):
        x = F.conv_transpose2d(x, sd[f"decoder.{k}.0.weight"], sd[f"decoder.{k}.0.bias"], stride=2, padding=1)
        x = torch.tanh(x) if i == n_hidden - 1 else F.leaky_relu(x)
        k += 1
    return x
