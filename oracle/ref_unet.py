"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not measured, never imported by the product.

CPU (torch, fp32 or fp64) functional restatement of the reference's tiny-SD hot path, written
against the reference's state_dict key names so the same weights drive both sides:

  * UNet forward            06_tiny_stable_diffusion/diffusion.py:263-276 (blocks :13-180)
  * schedule tables         06_tiny_stable_diffusion/utils.py:105-109, 135-141
  * extract                 06_tiny_stable_diffusion/utils.py:32-39
  * trainer q_sample + MSE  06_tiny_stable_diffusion/utils.py:111-119
  * sampler step            06_tiny_stable_diffusion/utils.py:143-166

Parity status: PINNED.  oracle/make_golden.py imports the real reference from /root/reference,
loads the weights produced by init_state_dict() into it, and stores the reference's own outputs
in tests/golden/*.pt; tests/test_oracle_golden.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# Architecture table (diffusion.py:204-261), derived from the constructor arguments.
# --------------------------------------------------------------------------------------------
def stage_table(channel_img: int, channel_multy: List[int], channel_base: int = 128):
    assert len(channel_multy) == 4
    m = [channel_base * i for i in channel_multy]
    enc = [
        [("conv", "encoders.0.0", channel_img, m[0], 1)],
        [("res", "encoders.1.0", m[0], m[0], True), ("attn", "encoders.1.1", m[0])],
        [("conv", "encoders.2.0", m[0], m[0], 2)],
        [("res", "encoders.3.0", m[0], m[1], True), ("attn", "encoders.3.1", m[1])],
        [("conv", "encoders.4.0", m[1], m[1], 2)],
        [("res", "encoders.5.0", m[1], m[2], True), ("attn", "encoders.5.1", m[2])],
        [("conv", "encoders.6.0", m[2], m[2], 2)],
        [("res", "encoders.7.0", m[2], m[3], True)],
    ]
    mid = [("res", "bottleneck.0", m[3], m[3], False), ("attn", "bottleneck.1", m[3]),
           ("res", "bottleneck.2", m[3], m[3], False)]
    dec = [
        [("res", "decoders.0.0", m[3] * 2, m[2], True)],
        [("res", "decoders.1.0", m[2] * 2, m[2], True), ("up", "decoders.1.1", m[2])],
        [("res", "decoders.2.0", m[2] * 2, m[1], True), ("attn", "decoders.2.1", m[1])],
        [("res", "decoders.3.0", m[1] * 2, m[1], True), ("attn", "decoders.3.1", m[1]), ("up", "decoders.3.2", m[1])],
        [("res", "decoders.4.0", m[1] * 2, m[0], True), ("attn", "decoders.4.1", m[0])],
        [("res", "decoders.5.0", m[0] * 2, m[0], True), ("attn", "decoders.5.1", m[0]), ("up", "decoders.5.2", m[0])],
        [("res", "decoders.6.0", m[0] * 2, m[0], True), ("attn", "decoders.6.1", m[0])],
        [("res", "decoders.7.0", m[0] * 2, m[0], True), ("attn", "decoders.7.1", m[0])],
    ]
    return enc, mid, dec, m


def param_shapes(channel_img: int, channel_multy: List[int], channel_base: int = 128, num_class: int = 10,
                 time_emb_dim: int = 512, d_model: int = 256, d_context: int = 512) -> Dict[str, tuple]:
    """state_dict keys and shapes in the reference's registration order (diffusion.py:184-261)."""
    enc, mid, dec, m = stage_table(channel_img, channel_multy, channel_base)
    out: Dict[str, tuple] = {}

    def lin(k, i, o, bias=True):
        out[k + ".weight"] = (o, i)
        if bias:
            out[k + ".bias"] = (o,)

    def conv(k, i, o, ks):
        out[k + ".weight"] = (o, i, ks, ks)
        out[k + ".bias"] = (o,)

    def norm(k, c):
        out[k + ".weight"] = (c,)
        out[k + ".bias"] = (c,)

    lin("time_embedding.mlp.0", d_model, time_emb_dim)
    lin("time_embedding.mlp.2", time_emb_dim, time_emb_dim)
    out["label_embedding.0.weight"] = (num_class + 1, d_model)
    lin("label_embedding.1", d_model, time_emb_dim)
    lin("label_embedding.3", time_emb_dim, time_emb_dim)

    def block(b):
        kind, key = b[0], b[1]
        if kind == "conv":
            conv(key, b[2], b[3], 3)
        elif kind == "up":
            conv(key + ".conv", b[2], b[2], 3)
        elif kind == "res":
            ci, co = b[2], b[3]
            norm(key + ".conv_1.0", ci)
            conv(key + ".conv_1.2", ci, co, 3)
            norm(key + ".conv_2.0", co)
            conv(key + ".conv_2.3", co, co, 3)
            lin(key + ".linear_time.1", time_emb_dim, co)
            if ci != co:
                conv(key + ".residual_layer", ci, co, 1)
        elif kind == "attn":
            c = b[2]
            norm(key + ".conv_1.0", c)
            conv(key + ".conv_1.1", c, c, 1)
            norm(key + ".atten_1.0", c)
            lin(key + ".atten_1.1.in_proj", c, 3 * c, bias=False)
            lin(key + ".atten_1.1.out_proj", c, c)
            norm(key + ".norm_2", c)
            lin(key + ".atten_2.q_proj", c, c, bias=False)
            lin(key + ".atten_2.k_proj", d_context, c, bias=False)
            lin(key + ".atten_2.v_proj", d_context, c, bias=False)
            lin(key + ".atten_2.out_proj", c, c)
            norm(key + ".norm_3", c)
            lin(key + ".linear_1", c, 8 * c)
            lin(key + ".linear_2", 4 * c, c)
            conv(key + ".conv_output", c, c, 1)

    for stage in enc:
        for b in stage:
            block(b)
    for b in mid:
        block(b)
    for stage in dec:
        for b in stage:
            block(b)
    norm("tail.0", m[0])
    conv("tail.2", m[0], channel_img, 3)
    return out


def init_state_dict(seed: int, channel_img: int, channel_multy: List[int], channel_base: int = 128,
                    num_class: int = 10, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic random-init weights "of that architecture": PyTorch-default-like scales
    (uniform +-1/sqrt(fan_in) for conv/linear weights and biases, N(0,1) embedding with a zero
    padding row, ones/zeros for norms), drawn only from torch.rand on a seeded CPU generator so
    the same tensors can be regenerated on any machine."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    shapes = param_shapes(channel_img, channel_multy, channel_base, num_class)
    norm_like = (".conv_1.0.", ".conv_2.0.", ".atten_1.0.", ".norm_2.", ".norm_3.", "tail.0.")
    for k, shp in shapes.items():
        if any(s in k for s in norm_like):
            # norms: non-trivial affine so that a dropped gamma/beta cannot hide
            u = torch.rand(shp, generator=g, dtype=torch.float64)
            sd[k] = (1.0 + 0.2 * (u - 0.5)) if k.endswith("weight") else 0.2 * (u - 0.5)
        elif k == "label_embedding.0.weight":
            u1 = torch.rand(shp, generator=g, dtype=torch.float64).clamp_min(1e-12)
            u2 = torch.rand(shp, generator=g, dtype=torch.float64)
            w = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)
            w[0] = 0.0  # padding_idx = 0 (diffusion.py:197)
            sd[k] = w
        else:
            if k.endswith("weight"):
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                last_fan_in = fan_in
            else:
                fan_in = last_fan_in
            bound = 1.0 / math.sqrt(fan_in)
            sd[k] = (torch.rand(shp, generator=g, dtype=torch.float64) * 2.0 - 1.0) * bound
        sd[k] = sd[k].to(dtype).contiguous()
    return sd


def state_dict_digest(sd: Dict[str, Tensor]) -> float:
    """Cheap order-sensitive checksum used to verify regenerated weights against the golden file."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += (i + 1) * float(v.double().sum()) + float(v.double().abs().sum())
    return acc


# --------------------------------------------------------------------------------------------
# Blocks
# --------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, dtype, max_period: float = 10000.0) -> Tensor:
    """diffusion.py:23-31.  freqs are built in fp32 exactly like the reference, then promoted."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return emb.to(dtype)


def _mlp(sd, k0, k1, x):
    h = F.linear(x, sd[k0 + ".weight"], sd[k0 + ".bias"])
    return F.linear(F.silu(h), sd[k1 + ".weight"], sd[k1 + ".bias"])


def _res_block(sd, key, x, temb, drop_mask: Optional[Tensor], p_drop: float):
    """diffusion.py:110-115"""
    h = F.group_norm(x, 32, sd[key + ".conv_1.0.weight"], sd[key + ".conv_1.0.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[key + ".conv_1.2.weight"], sd[key + ".conv_1.2.bias"], padding=1)
    tb = F.linear(F.silu(temb), sd[key + ".linear_time.1.weight"], sd[key + ".linear_time.1.bias"])
    h = h + tb[:, :, None, None]
    h = F.group_norm(h, 32, sd[key + ".conv_2.0.weight"], sd[key + ".conv_2.0.bias"], eps=1e-5)
    h = F.silu(h)
    if drop_mask is not None:
        h = h * drop_mask / (1.0 - p_drop)
    h = F.conv2d(h, sd[key + ".conv_2.3.weight"], sd[key + ".conv_2.3.bias"], padding=1)
    if key + ".residual_layer.weight" in sd:
        x = F.conv2d(x, sd[key + ".residual_layer.weight"], sd[key + ".residual_layer.bias"])
    return h + x


def _attn_block(sd, key, x, ctx, n_head: int = 8, use_sdpa: bool = False):
    """diffusion.py:138-158 (SelfAttention :46-58, CrossAttention :69-82)"""
    res_long = x
    n, c, hh, ww = x.shape
    h = F.group_norm(x, 32, sd[key + ".conv_1.0.weight"], sd[key + ".conv_1.0.bias"], eps=1e-6)
    h = F.conv2d(h, sd[key + ".conv_1.1.weight"], sd[key + ".conv_1.1.bias"])
    h = h.reshape(n, c, hh * ww).transpose(1, 2)  # [n, L, c]
    L = hh * ww
    dh = c // n_head
    # self attention
    y = F.layer_norm(h, (c,), sd[key + ".atten_1.0.weight"], sd[key + ".atten_1.0.bias"], eps=1e-5)
    qkv = F.linear(y, sd[key + ".atten_1.1.in_proj.weight"])
    q, k, v = qkv.split(c, dim=-1)
    q = q.reshape(n, L, n_head, dh).transpose(1, 2)
    k = k.reshape(n, L, n_head, dh).transpose(1, 2)
    v = v.reshape(n, L, n_head, dh).transpose(1, 2)
    if use_sdpa:  # the reference's own call (diffusion.py:55); used for the timed CPU baseline
        att = F.scaled_dot_product_attention(q, k, v)
    else:
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1) @ v
    att = att.transpose(1, 2).reshape(n, L, c)
    h = F.linear(att, sd[key + ".atten_1.1.out_proj.weight"], sd[key + ".atten_1.1.out_proj.bias"]) + h
    # cross attention: ctx [n, d_context] is viewed as ONE key/value token per head, so the softmax
    # over a single key is identically 1 and the block reduces to out_proj(v_proj(ctx)) broadcast
    # over positions; it is still evaluated literally here (q, k included) to mirror the reference.
    y = F.layer_norm(h, (c,), sd[key + ".norm_2.weight"], sd[key + ".norm_2.bias"], eps=1e-5)
    q = F.linear(y, sd[key + ".atten_2.q_proj.weight"]).reshape(n, L, n_head, dh).transpose(1, 2)
    k = F.linear(ctx, sd[key + ".atten_2.k_proj.weight"]).reshape(n, 1, n_head, dh).transpose(1, 2)
    v = F.linear(ctx, sd[key + ".atten_2.v_proj.weight"]).reshape(n, 1, n_head, dh).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1) @ v
    att = att.transpose(1, 2).reshape(n, L, c)
    h = F.linear(att, sd[key + ".atten_2.out_proj.weight"], sd[key + ".atten_2.out_proj.bias"]) + h
    # GEGLU feed-forward
    y = F.layer_norm(h, (c,), sd[key + ".norm_3.weight"], sd[key + ".norm_3.bias"], eps=1e-5)
    a, gate = F.linear(y, sd[key + ".linear_1.weight"], sd[key + ".linear_1.bias"]).chunk(2, dim=-1)
    h = F.linear(a * F.gelu(gate), sd[key + ".linear_2.weight"], sd[key + ".linear_2.bias"]) + h
    h = h.transpose(1, 2).reshape(n, c, hh, ww)
    return F.conv2d(h, sd[key + ".conv_output.weight"], sd[key + ".conv_output.bias"]) + res_long


def unet_forward(sd: Dict[str, Tensor], x: Tensor, t: Tensor, labels: Tensor, channel_multy: List[int],
                 channel_base: int = 128, dropout: float = 0.0,
                 drop_masks: Optional[Dict[str, Tensor]] = None, taps: Optional[Dict[str, Tensor]] = None,
                 use_sdpa: bool = False):
    """eps = Diffusion.forward(x, t, labels)  (diffusion.py:263-276), NCHW in / NCHW out.

    drop_masks: optional {res-block key: keep mask [n, c, h, w]} to replay a training-mode dropout
    pattern (diffusion.py:97); None = eval mode.  taps: optional dict that receives the output of
    every block keyed by its state_dict prefix (for per-block parity checks)."""
    dtype = sd["tail.2.weight"].dtype
    channel_img = sd["tail.2.weight"].shape[0]
    enc, mid, dec, _ = stage_table(channel_img, channel_multy, channel_base)
    x = x.to(dtype)
    emb = F.embedding(labels, sd["label_embedding.0.weight"], padding_idx=0)  # row 0 gets no gradient (:197)
    ctx = _mlp(sd, "label_embedding.1", "label_embedding.3", emb)
    temb = _mlp(sd, "time_embedding.mlp.0", "time_embedding.mlp.2",
                timestep_embedding(t, sd["time_embedding.mlp.0.weight"].shape[1], dtype))

    def run(block, x):
        kind, key = block[0], block[1]
        if kind == "conv":
            x = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=block[4], padding=1)
        elif kind == "up":
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[key + ".conv.weight"], sd[key + ".conv.bias"], padding=1)
        elif kind == "res":
            mask = drop_masks.get(key) if (drop_masks is not None and block[4]) else None
            x = _res_block(sd, key, x, temb, mask, dropout)
        else:
            x = _attn_block(sd, key, x, ctx, use_sdpa=use_sdpa)
        if taps is not None:
            taps[key] = x
        return x

    skips = []
    for stage in enc:
        for b in stage:
            x = run(b, x)
        skips.append(x)
    for b in mid:
        x = run(b, x)
    for stage in dec:
        x = torch.cat((x, skips.pop()), dim=1)
        for b in stage:
            x = run(b, x)
    x = F.group_norm(x, 32, sd["tail.0.weight"], sd["tail.0.bias"], eps=1e-5)
    return F.conv2d(F.silu(x), sd["tail.2.weight"], sd["tail.2.bias"], padding=1)


# --------------------------------------------------------------------------------------------
# DDPM process
# --------------------------------------------------------------------------------------------
def make_schedule(beta_1: float, beta_T: float, T: int) -> Dict[str, Tensor]:
    """utils.py:105-109 and :135-141.  Same torch calls in the same order and dtypes (fp32 linspace,
    then .double()), because any other construction differs in the last bits."""
    betas = torch.linspace(beta_1, beta_T, T).double()
    alphas = 1.0 - betas
    alphas_bar = torch.cumprod(alphas, dim=0)
    alphas_bar_prev = F.pad(alphas_bar, [1, 0], value=1)[:T]
    coeff1 = torch.sqrt(1.0 / alphas)
    return {
        "betas": betas,
        "sqrt_alphas_bar": torch.sqrt(alphas_bar),
        "sqrt_one_minus_alphas_bar": torch.sqrt(1.0 - alphas_bar),
        "coeff1": coeff1,
        "coeff2": coeff1 * (1.0 - alphas) / torch.sqrt(1.0 - alphas_bar),
        "posterior_var": betas * (1.0 - alphas_bar_prev) / (1.0 - alphas_bar),
    }


def sampler_variance(sched: Dict[str, Tensor]) -> Tensor:
    """utils.py:149: var = cat(posterior_var[1:2], betas[1:])"""
    return torch.cat([sched["posterior_var"][1:2], sched["betas"][1:]])


def extract(v: Tensor, t: Tensor, x_shape) -> Tensor:
    """utils.py:32-39"""
    out = torch.gather(v, index=t, dim=0).float()
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


def q_sample(sched, x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """utils.py:115-116"""
    return (extract(sched["sqrt_alphas_bar"], t, x0.shape) * x0
            + extract(sched["sqrt_one_minus_alphas_bar"], t, x0.shape) * noise)


def trainer_loss(sd, sched, x0, labels, t, noise, channel_multy, channel_base=128, dropout=0.0, drop_masks=None,
                 use_sdpa=False):
    """utils.py:111-119 with t and noise injected: returns the un-reduced MSE [B,C,H,W]."""
    x_t = q_sample(sched, x0, t, noise)
    pred = unet_forward(sd, x_t, t, labels, channel_multy, channel_base, dropout, drop_masks, use_sdpa=use_sdpa)
    return (pred - noise.to(pred.dtype)) ** 2


def sampler_step(sd, sched, x_t, labels, time_step: int, noise: Optional[Tensor], w: float,
                 channel_multy, channel_base=128, use_sdpa=False):
    """One iteration of utils.py:159-166 (two forwards, CFG combine, posterior mean, + sigma*z).
    Returns (x_{t-1}, eps_cond, eps_uncond)."""
    B = x_t.shape[0]
    t = torch.full((B,), time_step, dtype=torch.long)
    var = extract(sampler_variance(sched), t, x_t.shape)
    eps_c = unet_forward(sd, x_t, t, labels, channel_multy, channel_base, use_sdpa=use_sdpa).float()
    eps_u = unet_forward(sd, x_t, t, torch.zeros_like(labels), channel_multy, channel_base, use_sdpa=use_sdpa).float()
    eps = (1.0 + w) * eps_c - w * eps_u
    mean = extract(sched["coeff1"], t, x_t.shape) * x_t - extract(sched["coeff2"], t, x_t.shape) * eps
    z = noise if time_step > 0 else 0
    return mean + torch.sqrt(var) * z, eps_c, eps_u
