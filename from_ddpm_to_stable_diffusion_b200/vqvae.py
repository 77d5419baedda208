"""Drop-in for the reference's VQ-VAE latent codec (inference): image -> latent -> (quantise) -> image on the GPU.

Reference: 03_variational_autoencoder/models.py:135-185 (``VectorQuantizer``), :186-201 (``ResidualLayer``), :268-378
(``VQVAE``).  Same constructor, same ``encode`` / ``decode`` / ``forward`` return values, same ``state_dict`` keys and
default initialisation under the same torch seed: the torch sub-modules created here are parameter holders only.
This is the step before and after the denoiser in the repo's "Method 2" (06_tiny_stable_diffusion/03_train_with_vae.py:
24,69 encodes the images and decodes the sampled latents).

Compute path (libtinysd_b200.so, no CPU fallback): activations bf16 channels-last, every convolution on the tcgen05
GEMM core with its activation in the epilogue --
  Conv2d(k=4, s=2, p=1)          tsd_im2col_nhwc + tsd_gemm_fwd(bias, LeakyReLU)
  Conv2d(k=3) / Conv2d(k=1)      tsd_conv3x3_fwd_act / tsd_gemm_fwd (ReLU / LeakyReLU / residual in the epilogue)
  ConvTranspose2d(k=4, s=2, p=1) ONE 3x3 implicit GEMM producing the four output parities as 4 x Cout channels (each
                                 parity uses a 2x2 subset of the 3x3 window; the rest of the packed weight is zero),
                                 then tsd_depth_to_space2 (tsd_d2s_to_nchw_f32 + Tanh for the image layer)
  VectorQuantizer                tsd_vq_nearest (fp32 distances in the reference's formula, first minimum)
Channel counts are padded with zero weights to the GEMM tile (128 output columns, 64 input channels), so any
``hidden_dims`` that are multiples of 8 work.  Training the codec (backward) is outside the hot path and not provided.
"""
from typing import List

import torch
import torch.nn as nn

from . import ops
from ._lib import call, f32

BF16 = torch.bfloat16
F32 = torch.float32


def _up(v, m):
    return (v + m - 1) // m * m


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the compute path lives in libtinysd_b200.so")


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the compute path lives in libtinysd_b200.so")


def _residual_holder(ch):
    h = _Holder()
    h.conv = _Seq(nn.Conv2d(ch, ch, kernel_size=3, stride=1, padding=1, bias=False), nn.ReLU(True),
                  nn.Conv2d(ch, ch, kernel_size=1, stride=1, padding=0, bias=False))
    h.residual = nn.Identity()
    return h


class _VQHolder(_Holder):
    def __init__(self, num_embeddings, embedding_dim, beta):
        super().__init__()
        self.K, self.D, self.beta = num_embeddings, embedding_dim, beta
        self.embedding = nn.Embedding(self.K, self.D)
        self.embedding.weight.data.uniform_(-1 / self.K, 1 / self.K)


class VQVAE(nn.Module):
    N_RES = 6

    def __init__(self, in_channels: int, embedding_dim: int, num_embeddings: int, hidden_dims: List = None,
                 beta: float = 0.25, img_size: int = 64, **kwargs):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.img_size = img_size
        self.beta = beta
        if hidden_dims is None:
            hidden_dims = [64, 128, 256]
        hidden_dims = list(hidden_dims)
        if embedding_dim > 16:
            raise RuntimeError("VQVAE (B200): embedding_dim up to 16 is supported")
        if in_channels > 8 or any(h % 8 for h in hidden_dims):
            raise RuntimeError("VQVAE (B200): in_channels <= 8 and hidden_dims multiples of 8 are supported")
        self.in_channels = in_channels
        self.hidden_dims = hidden_dims
        # ---- parameter tree in the reference's registration order (models.py:281-343)
        mods, ci = [], in_channels
        for h in hidden_dims:
            mods.append(_Seq(nn.Conv2d(ci, h, kernel_size=4, stride=2, padding=1), nn.LeakyReLU()))
            ci = h
        mods.append(_Seq(nn.Conv2d(ci, ci, kernel_size=3, stride=1, padding=1), nn.LeakyReLU()))
        for _ in range(self.N_RES):
            mods.append(_residual_holder(ci))
        mods.append(nn.LeakyReLU())
        mods.append(_Seq(nn.Conv2d(ci, embedding_dim, kernel_size=1, stride=1), nn.LeakyReLU()))
        self.encoder = _Seq(*mods)
        self.vq_layer = _VQHolder(num_embeddings, embedding_dim, beta)
        top = hidden_dims[-1]
        mods = [_Seq(nn.Conv2d(embedding_dim, top, kernel_size=3, stride=1, padding=1), nn.LeakyReLU())]
        for _ in range(self.N_RES):
            mods.append(_residual_holder(top))
        mods.append(nn.LeakyReLU())
        rev = hidden_dims[::-1]
        for i in range(len(rev) - 1):
            mods.append(_Seq(nn.ConvTranspose2d(rev[i], rev[i + 1], kernel_size=4, stride=2, padding=1), nn.LeakyReLU()))
        mods.append(_Seq(nn.ConvTranspose2d(rev[-1], out_channels=3, kernel_size=4, stride=2, padding=1), nn.Tanh()))
        self.decoder = _Seq(*mods)
        self._packed = None
        self._packed_sig = None

    # ------------------------------------------------------------------ nn.Module plumbing
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._packed = None
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._packed = None
        return out

    # ------------------------------------------------------------------ weight packing (one-off torch glue, bf16 GEMM layouts)
    @staticmethod
    def _pack_conv(w, cin_p, cout_p):
        """Conv2d weight [Co][Ci][KH][KW] -> bf16 [cout_p][KH*KW*cin_p], k = (ky*KW + kx)*cin_p + ci, zero padded."""
        co, ci, kh, kw = w.shape
        out = torch.zeros(cout_p, kh * kw, cin_p, device=w.device, dtype=F32)
        out[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
        return out.reshape(cout_p, kh * kw * cin_p).to(BF16).contiguous()

    @staticmethod
    def _pack_convT(w, bias, cin_p, coq, n_p):
        """ConvTranspose2d(k=4, s=2, p=1) weight [Ci][Co][4][4] -> the 3x3 conv weight bf16 [n_p][9*cin_p] that yields the
        four output parities q = py*2 + px as channel blocks [q*coq, q*coq + Co): output row 2i + py takes input row
        i + dy with kernel row ky where (py, dy) -> ky is (0, 0) -> 1, (0, -1) -> 3, (1, +1) -> 0, (1, 0) -> 2."""
        ci, co = w.shape[:2]
        big = torch.zeros(n_p, 3, 3, cin_p, device=w.device, dtype=F32)
        kmap = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}
        for py in (0, 1):
            for px in (0, 1):
                q = py * 2 + px
                for dy, ky in kmap[py].items():
                    for dx, kx in kmap[px].items():
                        big[q * coq:q * coq + co, dy + 1, dx + 1, :ci] = w[:, :, ky, kx].t()
        b = torch.zeros(n_p, device=w.device, dtype=F32)
        for q in range(4):
            b[q * coq:q * coq + co] = bias
        return big.reshape(n_p, 9 * cin_p).to(BF16).contiguous(), b

    @staticmethod
    def _pad_bias(b, n_p):
        out = torch.zeros(n_p, device=b.device, dtype=F32)
        out[:b.shape[0]] = b
        return out

    def _weights(self):
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and sig == self._packed_sig:
            return self._packed
        P = dict(self.named_parameters())
        W = {}
        nh = len(self.hidden_dims)
        # encoder
        ci_p = 8  # the fp32 NCHW image is converted to bf16 NHWC padded to 8 channels
        for i, h in enumerate(self.hidden_dims):
            co_p = _up(h, 128)
            W[f"e{i}.w"] = self._pack_conv(P[f"encoder.{i}.0.weight"], ci_p, co_p)
            W[f"e{i}.b"] = self._pad_bias(P[f"encoder.{i}.0.bias"], co_p)
            ci_p = co_p
        cp = ci_p
        W["e_mid.w"] = self._pack_conv(P[f"encoder.{nh}.0.weight"], cp, cp)
        W["e_mid.b"] = self._pad_bias(P[f"encoder.{nh}.0.bias"], cp)
        for r in range(self.N_RES):
            W[f"e_res{r}.w3"] = self._pack_conv(P[f"encoder.{nh + 1 + r}.conv.0.weight"], cp, cp)
            W[f"e_res{r}.w1"] = self._pack_conv(P[f"encoder.{nh + 1 + r}.conv.2.weight"], cp, cp)
        k_out = nh + 1 + self.N_RES + 1
        W["e_out.w"] = self._pack_conv(P[f"encoder.{k_out}.0.weight"], cp, 128)
        W["e_out.b"] = self._pad_bias(P[f"encoder.{k_out}.0.bias"], 128)
        # decoder
        tp = _up(self.hidden_dims[-1], 128)
        W["d_in.w"] = self._pack_conv(P["decoder.0.0.weight"], 64, tp)  # latent padded to 64 channels
        W["d_in.b"] = self._pad_bias(P["decoder.0.0.bias"], tp)
        for r in range(self.N_RES):
            W[f"d_res{r}.w3"] = self._pack_conv(P[f"decoder.{1 + r}.conv.0.weight"], tp, tp)
            W[f"d_res{r}.w1"] = self._pack_conv(P[f"decoder.{1 + r}.conv.2.weight"], tp, tp)
        rev = self.hidden_dims[::-1]
        ci_p = tp
        base = 1 + self.N_RES + 1
        self._ups = []
        for i in range(len(rev)):
            last = i == len(rev) - 1
            co = 3 if last else rev[i + 1]
            coq = co if last else _up(co, 64)
            n_p = _up(4 * coq, 128)
            W[f"d_up{i}.w"], W[f"d_up{i}.b"] = self._pack_convT(P[f"decoder.{base + i}.0.weight"],
                                                               P[f"decoder.{base + i}.0.bias"], ci_p, coq, n_p)
            self._ups.append((coq, n_p, last))
            ci_p = coq
        W["codebook"] = P["vq_layer.embedding.weight"].detach().float().contiguous()
        self._packed, self._packed_sig = W, sig
        return W

    # ------------------------------------------------------------------ kernels
    @staticmethod
    def _conv3x3(x, n, H, Wd, w, bias, act, residual=None):
        cout = w.shape[0]
        d = torch.empty(n * H * Wd, cout, device=x.device, dtype=BF16)
        call("tsd_conv3x3_fwd_act", x, None, x.shape[1], 0, n, H, Wd, 1, w, cout, bias, None, 0, residual, act, d)
        return d

    @staticmethod
    def _gemm(a, w, bias, act, residual=None):
        M, K = a.shape
        N = w.shape[0]
        d = torch.empty(M, N, device=a.device, dtype=BF16)
        call("tsd_gemm_fwd", a, None, K, 0, M, w, N, bias, None, 1, residual, act, d)
        return d

    def _res_stack(self, x, n, H, Wd, W, prefix):
        """6 x ResidualLayer (models.py:186-201) then LeakyReLU (:302 / :325): conv3x3 (no bias) + ReLU in its epilogue,
        conv1x1 (no bias) + the skip in its epilogue; the stage's closing LeakyReLU rides on the last skip add."""
        for r in range(self.N_RES):
            a = self._conv3x3(x, n, H, Wd, W[f"{prefix}{r}.w3"], None, ops.EPI_RELU)
            x = self._gemm(a, W[f"{prefix}{r}.w1"], None, ops.EPI_LRELU if r == self.N_RES - 1 else 0, residual=x)
        return x

    # ------------------------------------------------------------------ the reference API
    @torch.no_grad()
    def encode(self, x) -> List[torch.Tensor]:
        """[N, C, H, W] fp32 -> [latents [N, D, H/2^k, W/2^k] fp32]  (models.py:345-353)"""
        if not x.is_cuda:
            raise RuntimeError("VQVAE (B200) runs on CUDA only; there is no CPU fallback")
        W = self._weights()
        n, c, H, Wd = x.shape
        assert c == self.in_channels
        cur = ops.nchw_to_nhwc_pad(x.contiguous().float(), 8)
        cin = 8
        for i, _ in enumerate(self.hidden_dims):
            w = W[f"e{i}.w"]
            Ho, Wo = H // 2, Wd // 2
            patch = torch.empty(n * Ho * Wo, w.shape[1], device=x.device, dtype=BF16)
            call("tsd_im2col_nhwc", cur, patch, n, H, Wd, cin, 4, 4, 2, 1, w.shape[1])
            cur = self._gemm(patch, w, W[f"e{i}.b"], ops.EPI_LRELU)
            H, Wd, cin = Ho, Wo, w.shape[0]
        cur = self._conv3x3(cur, n, H, Wd, W["e_mid.w"], W["e_mid.b"], ops.EPI_LRELU)
        cur = self._res_stack(cur, n, H, Wd, W, "e_res")
        zz = self._gemm(cur, W["e_out.w"], W["e_out.b"], ops.EPI_LRELU)
        out = torch.empty(n, self.embedding_dim, H, Wd, device=x.device, dtype=F32)
        call("tsd_nhwc_to_nchw_f32", zz, out, n, H * Wd, self.embedding_dim, zz.shape[1])
        return [out]

    @torch.no_grad()
    def quantize(self, latents):
        """VectorQuantizer.forward (models.py:149-185) -> (quantised latents [N, D, h, w], vq_loss scalar, indices [N*h*w])"""
        W = self._weights()
        z = latents.contiguous().float()
        n, D, h, w = z.shape
        assert D == self.embedding_dim
        idx = torch.empty(n * h * w, device=z.device, dtype=torch.int64)
        zq = torch.empty_like(z)
        loss = torch.empty((), device=z.device, dtype=F32)
        fn = ops._lib.lib().tsd_vq_scratch_floats
        fn.restype = __import__("ctypes").c_int64
        scratch = torch.empty(int(fn(n, h * w)), device=z.device, dtype=F32)
        call("tsd_vq_nearest", z, W["codebook"], idx, zq, scratch, loss, f32(self.beta), n, h * w, D, self.num_embeddings)
        return zq, loss, idx

    @torch.no_grad()
    def decode(self, z) -> torch.Tensor:
        """[B, D, h, w] fp32 -> image [B, 3, h*2^k, w*2^k] fp32 in (-1, 1)  (models.py:355-363)"""
        if not z.is_cuda:
            raise RuntimeError("VQVAE (B200) runs on CUDA only; there is no CPU fallback")
        W = self._weights()
        n, D, H, Wd = z.shape
        assert D == self.embedding_dim
        cur = ops.nchw_to_nhwc_pad(z.contiguous().float(), 64)
        cur = self._conv3x3(cur, n, H, Wd, W["d_in.w"], W["d_in.b"], ops.EPI_LRELU)
        cur = self._res_stack(cur, n, H, Wd, W, "d_res")
        for i, (coq, n_p, last) in enumerate(self._ups):
            big = self._conv3x3(cur, n, H, Wd, W[f"d_up{i}.w"], W[f"d_up{i}.b"], ops.EPI_TANH if last else ops.EPI_LRELU)
            if last:
                out = torch.empty(n, 3, 2 * H, 2 * Wd, device=z.device, dtype=F32)
                call("tsd_d2s_to_nchw_f32", big, out, n, H, Wd, 3, n_p)
                return out
            cur = torch.empty(n * 4 * H * Wd, coq, device=z.device, dtype=BF16)
            call("tsd_depth_to_space2", big, cur, n, H, Wd, coq, n_p)
            H, Wd = 2 * H, 2 * Wd
        raise AssertionError("unreachable")

    def forward(self, x, **kwargs) -> List[torch.Tensor]:
        """[reconstruction, input, vq_loss]  (models.py:365-368)"""
        encoding = self.encode(x)[0]
        quantized, vq_loss, _ = self.quantize(encoding)
        return [self.decode(quantized), x, vq_loss]

    def loss_function(self, *args, **kwargs) -> dict:
        """models.py:370-375 on the forward's outputs (plain torch reductions; a metric, not part of the hot path)."""
        recons, x, vq_loss = args[0], args[1], args[2]
        recons_loss = torch.nn.functional.mse_loss(recons, x)
        return {'loss': recons_loss + vq_loss, 'Reconstruction_Loss': recons_loss, 'VQ_Loss': vq_loss}
