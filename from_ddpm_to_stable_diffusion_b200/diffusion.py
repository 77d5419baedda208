"""Drop-in for the reference's ``diffusion.py``: the class-conditional UNet eps-predictor.

Reference: 06_tiny_stable_diffusion/diffusion.py:183-276 (``Diffusion``).  Same constructor, same
``forward(x, time, context)`` signature, same 425-tensor ``state_dict`` (fp32, OIHW conv weights),
same default initialisation under the same torch seed.  Everything below the module surface is the
sm_100a library: the torch sub-modules created here are *parameter holders only*; their forward
methods are never called.  There is no CPU / eager fallback: a non-CUDA input raises.
"""
from typing import List

import torch
import torch.nn as nn

from . import ops


# --------------------------------------------------------------------------------------------
# Stage table (what diffusion.py:204-261 builds), derived from the constructor arguments.
# entries: ("conv", Cin, Cout, stride) | ("res", Cin, Cout, dropout?) | ("attn", C) | ("up", C)
# --------------------------------------------------------------------------------------------
def _stages(channel_img, multy):
    m = multy
    enc = [
        [("conv", channel_img, m[0], 1)],
        [("res", m[0], m[0], True), ("attn", m[0])],
        [("conv", m[0], m[0], 2)],
        [("res", m[0], m[1], True), ("attn", m[1])],
        [("conv", m[1], m[1], 2)],
        [("res", m[1], m[2], True), ("attn", m[2])],
        [("conv", m[2], m[2], 2)],
        [("res", m[2], m[3], True)],
    ]
    mid = [("res", m[3], m[3], False), ("attn", m[3]), ("res", m[3], m[3], False)]
    dec = [
        [("res", m[3] * 2, m[2], True)],
        [("res", m[2] * 2, m[2], True), ("up", m[2])],
        [("res", m[2] * 2, m[1], True), ("attn", m[1])],
        [("res", m[1] * 2, m[1], True), ("attn", m[1]), ("up", m[1])],
        [("res", m[1] * 2, m[0], True), ("attn", m[0])],
        [("res", m[0] * 2, m[0], True), ("attn", m[0]), ("up", m[0])],
        [("res", m[0] * 2, m[0], True), ("attn", m[0])],
        [("res", m[0] * 2, m[0], True), ("attn", m[0])],
    ]
    return enc, mid, dec


class _Holder(nn.Module):
    """Names parameters like the reference's blocks do; never executed."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the compute path lives in libtinysd_b200.so")


class _Seq(nn.Sequential):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the compute path lives in libtinysd_b200.so")


def _res_holder(ci, co, n_time):
    h = _Holder()
    h.conv_1 = _Seq(nn.GroupNorm(32, ci), nn.Identity(), nn.Conv2d(ci, co, 3, padding=1))
    h.conv_2 = _Seq(nn.GroupNorm(32, co), nn.Identity(), nn.Identity(), nn.Conv2d(co, co, 3, padding=1))
    h.linear_time = _Seq(nn.Identity(), nn.Linear(n_time, co))
    h.residual_layer = nn.Conv2d(ci, co, 1) if ci != co else nn.Identity()
    return h


def _attn_holder(c, d_context):
    h = _Holder()
    h.conv_1 = _Seq(nn.GroupNorm(32, c, eps=1e-6), nn.Conv2d(c, c, 1))
    sa = _Holder()
    sa.in_proj = nn.Linear(c, 3 * c, bias=False)
    sa.out_proj = nn.Linear(c, c)
    h.atten_1 = _Seq(nn.LayerNorm(c), sa)
    h.norm_2 = nn.LayerNorm(c)
    ca = _Holder()
    ca.q_proj = nn.Linear(c, c, bias=False)
    ca.k_proj = nn.Linear(d_context, c, bias=False)
    ca.v_proj = nn.Linear(d_context, c, bias=False)
    ca.out_proj = nn.Linear(c, c)
    h.atten_2 = ca
    h.norm_3 = nn.LayerNorm(c)
    h.linear_1 = nn.Linear(c, 8 * c)
    h.linear_2 = nn.Linear(4 * c, c)
    h.conv_output = nn.Conv2d(c, c, 1)
    return h


def _up_holder(c):
    h = _Holder()
    h.conv = nn.Conv2d(c, c, 3, padding=1)
    return h


class Diffusion(nn.Module):
    N_HEAD = 8

    def __init__(self, channel_img: int, channel_multy: List[int], channel_base: int = 128, num_class: int = 10,
                 dropout: float = 0.0, time_emb_dim: int = 512):
        super().__init__()
        d_model = 256
        assert len(channel_multy) == 4
        multy = [channel_base * i for i in channel_multy]
        for c in multy:
            assert c % 32 == 0
        self.channel_img = channel_img
        self.multy = multy
        self.dropout = float(dropout)
        self.time_emb_dim = time_emb_dim
        self.d_model = d_model
        self.num_class = num_class

        # ---- parameter tree, registered in the reference's order so that the default init consumes
        # the torch RNG identically (same seed -> same weights as the reference module)
        te = _Holder()
        te.mlp = _Seq(nn.Linear(d_model, time_emb_dim), nn.Identity(), nn.Linear(time_emb_dim, time_emb_dim))
        self.time_embedding = te
        self.label_embedding = _Seq(nn.Embedding(num_class + 1, d_model, padding_idx=0),
                                    nn.Linear(d_model, time_emb_dim), nn.Identity(),
                                    nn.Linear(time_emb_dim, time_emb_dim))
        enc, mid, dec = _stages(channel_img, multy)
        self._enc, self._mid, self._dec = enc, mid, dec

        def build(b):
            if b[0] == "conv":
                return nn.Conv2d(b[1], b[2], 3, stride=b[3], padding=1)
            if b[0] == "res":
                return _res_holder(b[1], b[2], time_emb_dim)
            if b[0] == "attn":
                return _attn_holder(b[1], time_emb_dim)
            return _up_holder(b[1])

        self.encoders = nn.ModuleList([_Seq(*[build(b) for b in st]) for st in enc])
        self.bottleneck = _Seq(*[build(b) for b in mid])
        self.decoders = nn.ModuleList([_Seq(*[build(b) for b in st]) for st in dec])
        self.tail = _Seq(nn.GroupNorm(32, multy[0]), nn.Identity(), nn.Conv2d(multy[0], channel_img, 3, padding=1))

        from .engine import UNetEngine
        self._engine = UNetEngine(self)

    # nn.Module plumbing ------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return out

    def forward(self, x, time, context):
        """x [B,C,H,W] fp32, time [B] int64, context [B] int64 (0 = unconditional) -> eps [B,C,H,W] fp32."""
        if not x.is_cuda:
            raise RuntimeError("from_ddpm_to_stable_diffusion_b200.Diffusion runs on CUDA (sm_100a) only; "
                               "there is no CPU fallback")
        return self._engine.apply(x, time, context)
