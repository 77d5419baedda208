"""Thin tensor-level wrappers over the C ABI (include/tinysd_b200.h).

PyTorch is used here only as the device allocator and stream provider; every computation below is
a kernel in libtinysd_b200.so.  Activations are bf16 channels-last, stored as 2-D [n*h*w, c].
"""
import torch

from . import _lib
from ._lib import call, f32, i64, u64

BF16 = torch.bfloat16
F32 = torch.float32


def _chk(t, dtype):
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.device, t.is_contiguous(), t.dtype, dtype)
    return t


def empty_bf16(*shape, like):
    return torch.empty(shape, device=like.device, dtype=BF16)


# ------------------------------------------------------------------ dense contractions (tcgen05)
def _gn_part_for(d, M, n_out):
    """fp32 [M / 32][n_out][2] buffer for the GroupNorm partial sums the producing kernel leaves behind; attached to the
    output tensor so that gn_stats() on it needs no pass over the data (lost, harmlessly, when the tensor is copied)."""
    part = torch.empty(M // 32, n_out, 2, device=d.device, dtype=F32)
    d._gn_part = part
    return part


def gemm(a0, w, n_out, a1=None, bias=None, row_bias=None, rows_per_sample=1, residual=None, geglu=False, gn=False,
         out=None, gn_part=None):
    """gn: also produce the GroupNorm partial sums of the output (plain epilogue, M % 64 == 0).  out / gn_part: write
    into caller-provided (row-slices of) buffers instead of allocating."""
    M, c0 = a0.shape
    c1 = a1.shape[1] if a1 is not None else 0
    d = empty_bf16(M, n_out // 2 if geglu else n_out, like=a0) if out is None else _chk(out, BF16)
    if gn_part is not None:
        call("tsd_gemm_fwd_gn", _chk(a0, BF16), a1, c0, c1, M, _chk(w, BF16), n_out, bias, row_bias, rows_per_sample,
             residual, 0, d, _chk(gn_part, F32))
        return d
    if gn and not geglu and M % 64 == 0:
        call("tsd_gemm_fwd_gn", _chk(a0, BF16), a1, c0, c1, M, _chk(w, BF16), n_out, bias, row_bias, rows_per_sample,
             residual, 0, d, _gn_part_for(d, M, n_out))
        return d
    call("tsd_gemm_fwd", _chk(a0, BF16), a1, c0, c1, M, _chk(w, BF16), n_out, bias, row_bias, rows_per_sample,
         residual, 1 if geglu else 0, d)
    return d


def gemm_geglu_bwd(a, w_geglu, bias_geglu, dgg, dbias=None):
    """dh8 [M, 8C] of the GEGLU feed-forward input, pre-activations recomputed (no stored h8); dbias (+=) optional."""
    M, K = a.shape
    N = w_geglu.shape[0]
    dh8 = empty_bf16(M, N, like=a)
    call("tsd_gemm_geglu_bwd", _chk(a, BF16), M, K, _chk(w_geglu, BF16), N, bias_geglu, _chk(dgg, BF16), dh8, dbias)
    return dh8


def conv3x3(x0, n_img, H, W, w, cout, x1=None, stride=1, bias=None, row_bias=None, rows_per_sample=0, residual=None,
            gn=False):
    """gn: also produce the GroupNorm partial sums of the output (output pixels per image a multiple of 64)"""
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    hw_out = (H // stride) * (W // stride)
    d = empty_bf16(n_img * hw_out, cout, like=x0)
    if gn and hw_out % 64 == 0:
        call("tsd_conv3x3_fwd_gn", _chk(x0, BF16), x1, c0, c1, n_img, H, W, stride, _chk(w, BF16), cout, bias, row_bias,
             rows_per_sample, residual, 0, d, _gn_part_for(d, n_img * hw_out, cout))
        return d
    call("tsd_conv3x3_fwd", _chk(x0, BF16), x1, c0, c1, n_img, H, W, stride, _chk(w, BF16), cout, bias, row_bias,
         rows_per_sample, residual, d)
    return d


def gemm_dgrad(dy, w, k_in, residual=None):
    M, N = dy.shape
    dx = empty_bf16(M, k_in, like=dy)
    call("tsd_gemm_dgrad", _chk(dy, BF16), M, N, _chk(w, BF16), k_in, residual, dx)
    return dx


def conv3x3_dgrad(dy, n_img, H, W, w, cin, residual=None):
    cout = dy.shape[1]
    dx = empty_bf16(n_img * H * W, cin, like=dy)
    call("tsd_conv3x3_dgrad", _chk(dy, BF16), n_img, H, W, cout, _chk(w, BF16), cin, residual, dx)
    return dx


def gemm_wgrad(dy, x0, dw, x1=None):
    M, N = dy.shape
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    call("tsd_gemm_wgrad", _chk(dy, BF16), _chk(x0, BF16), x1, c0, c1, M, N, _chk(dw, F32))


def conv3x3_wgrad(dy, x0, n_img, H, W, dw_packed, x1=None, stride=1):
    cout = dy.shape[1]
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    call("tsd_conv3x3_wgrad", _chk(dy, BF16), _chk(x0, BF16), x1, c0, c1, n_img, H, W, stride, cout, _chk(dw_packed, F32))


# ------------------------------------------------------------------ norms
GN_FROM_PARTS = bool(int(__import__("os").environ.get("TSD_GN_EPILOGUE_STATS", "1")))


def gn_stats(x0, n_img, hw, eps, scratch, x1=None):
    """(mean, rstd) per (image, group).  When the kernels that wrote x0 (and x1) left their per-channel partial sums
    behind (gemm / conv3x3 with gn=True), the statistics come from those -- no pass over the tensors."""
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    stats = torch.empty(n_img, 32, 2, device=x0.device, dtype=F32)
    p0 = getattr(x0, "_gn_part", None)
    p1 = getattr(x1, "_gn_part", None) if x1 is not None else None
    if GN_FROM_PARTS and p0 is not None and (x1 is None or p1 is not None) and hw % 64 == 0 \
            and (c1 == 0 or c0 % ((c0 + c1) // 32) == 0):
        call("tsd_gn_stats_from_parts", p0, p1, c0, c1, n_img, hw, f32(eps), stats)
        return stats
    call("tsd_gn_stats", _chk(x0, BF16), x1, c0, c1, n_img, hw, f32(eps), scratch, stats)
    return stats


def gn_apply(x0, n_img, hw, stats, gamma, beta, silu, x1=None, drop_p=0.0, seed=0, rng=None):
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    out = empty_bf16(n_img * hw, c0 + c1, like=x0)
    call("tsd_gn_apply", x0, x1, c0, c1, n_img, hw, stats, gamma, beta, int(silu), f32(drop_p), u64(seed), out, rng)
    return out


def gn_bwd(dy, x0, n_img, hw, stats, gamma, beta, silu, dgamma, dbeta, x1=None, drop_p=0.0, seed=0, radd=None, rng=None,
           colsum_out=None, colsum_total=None):
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    ab = torch.empty(n_img, c0 + c1, 2, device=x0.device, dtype=F32)
    dx0 = empty_bf16(n_img * hw, c0, like=x0)
    dx1 = empty_bf16(n_img * hw, c1, like=x0) if c1 else None
    call("tsd_gn_bwd", _chk(dy, BF16), x0, x1, c0, c1, n_img, hw, stats, gamma, beta, int(silu), f32(drop_p), u64(seed),
         ab, radd, dx0, dx1, dgamma, dbeta, rng, colsum_out, colsum_total)
    return dx0, dx1


def ln_fwd(x, gamma, beta, eps=1e-5):
    M, C = x.shape
    out = torch.empty_like(x)
    call("tsd_ln_fwd", _chk(x, BF16), M, C, gamma, beta, f32(eps), out)
    return out


def ln_bwd(dy, x, gamma, dgamma, dbeta, eps=1e-5, radd=None, rows_per_sample=0, colsum_out=None, colsum_total=None):
    M, C = x.shape
    dx = torch.empty_like(x)
    call("tsd_ln_bwd", _chk(dy, BF16), x, M, C, gamma, f32(eps), radd, dx, dgamma, dbeta, int(rows_per_sample), colsum_out,
         colsum_total)
    return dx


# ------------------------------------------------------------------ attention
def attn_fwd(qkv, B, L, C, heads=8, need_lse=False):
    out = empty_bf16(B * L, C, like=qkv)
    lse = torch.empty(B, heads, L, device=qkv.device, dtype=F32) if need_lse else None
    ws = torch.empty(B * heads, device=qkv.device, dtype=F32)  # max |k|^2 per (sample, head): score bound
    call("tsd_attn_fwd_ws", _chk(qkv, BF16), out, lse, ws, B, L, C, heads)
    return out, lse


def attn_bwd(qkv, out, dout, lse, B, L, C, heads=8):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, heads, L, device=qkv.device, dtype=F32)
    # fp32 scratch for the one-pass backward (head_dim 16, L % 256 == 0; head_dim 32, L % 128 == 0): dQ partials are reduced there
    dh = C // heads
    ws = torch.empty(B * L * C, device=qkv.device, dtype=F32) \
        if ((dh == 16 and L % 256 == 0) or (dh == 32 and L % 128 == 0 and L >= 256)) else None
    call("tsd_attn_bwd_ws", qkv, out, _chk(dout, BF16), lse, delta, dqkv, ws, B, L, C, heads)
    return dqkv


# ------------------------------------------------------------------ elementwise
def add(a, b):
    out = torch.empty_like(a)
    call("tsd_add_bf16", _chk(a, BF16), _chk(b, BF16), out, i64(a.numel()))
    return out


def geglu_fwd(h8):
    M, H2 = h8.shape
    out = empty_bf16(M, H2 // 2, like=h8)
    call("tsd_geglu_fwd", h8, out, i64(M), H2 // 2)
    return out


def geglu_bwd(h8, dout, dbias=None):
    """dh8 of the GEGLU; dbias (fp32 [8C]) += column sums of dh8 when given (bias gradient of the C -> 8C linear)"""
    dh8 = torch.empty_like(h8)
    call("tsd_geglu_bwd", h8, _chk(dout, BF16), dh8, i64(h8.shape[0]), h8.shape[1] // 2, dbias)
    return dh8


def upsample2_fwd(x, n_img, H, W):
    C = x.shape[1]
    out = empty_bf16(n_img * 4 * H * W, C, like=x)
    call("tsd_upsample2_fwd", x, out, n_img, H, W, C)
    return out


def upsample2_bwd(dout, n_img, H, W):
    C = dout.shape[1]
    din = empty_bf16(n_img * H * W, C, like=dout)
    call("tsd_upsample2_bwd", _chk(dout, BF16), din, n_img, H, W, C)
    return din


def zero_stuff2(x, n_img, H, W):
    C = x.shape[1]
    out = empty_bf16(n_img * 4 * H * W, C, like=x)
    call("tsd_zero_stuff2", _chk(x, BF16), out, n_img, H, W, C)
    return out


def colsum(x, n_samples, rows_per_sample, total=None, out=None):
    """[n_samples*rows_per_sample, C] bf16 -> fp32 [n_samples, C]; total[C] += the sum over all samples (optional).
    ``out`` (zero-filled fp32 [n_samples, C]) may be supplied by the caller."""
    C = x.shape[1]
    if out is None:
        out = torch.zeros(n_samples, C, device=x.device, dtype=F32)
    call("tsd_colsum", _chk(x, BF16), n_samples, rows_per_sample, C, out, None if total is None else _chk(total, F32))
    return out


def reduce_rows_into(src, dst):
    """dst[c] += sum_n src[n][c]"""
    call("tsd_reduce_rows_f32", _chk(src, F32), src.shape[0], src.shape[1], _chk(dst, F32))


def bias_grad(dy, n_samples, rows_per_sample, db):
    """db[c] += sum over all rows of dy; returns the per-sample sums."""
    return colsum(dy, n_samples, rows_per_sample, total=db)


# ------------------------------------------------------------------ conditioning (fp32, small)
def small_linear(x, w, bias=None, silu_in=False):
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=F32)
    call("tsd_small_linear_fwd", _chk(x, F32), _chk(w, F32), bias, out, M, N, K, int(silu_in))
    return out


def small_linear_bwd(dy, x, w, dx, dw, db, silu_in=False, accumulate_dx=False):
    M, N = dy.shape
    K = x.shape[1]
    call("tsd_small_linear_bwd", _chk(dy, F32), _chk(x, F32), _chk(w, F32), dx, dw, db, M, N, K, int(silu_in),
         int(accumulate_dx))


def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _int_array(vals):
    import ctypes
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def small_linear_many(xs, ws, biases=None, silu_in=False):
    """[f(x_i) w_i^T + b_i] for up to 16 layers that share the batch dimension, one launch (fp32)."""
    outs = []
    for c in range(0, len(xs), 16):
        x, w = xs[c:c + 16], ws[c:c + 16]
        b = None if biases is None else biases[c:c + 16]
        M = x[0].shape[0]
        out = [torch.empty(M, wi.shape[0], device=wi.device, dtype=F32) for wi in w]
        for xi, wi in zip(x, w):
            _chk(xi, F32), _chk(wi, F32)
        call("tsd_small_linear_many_fwd", len(x), _ptr_array(x), _ptr_array(w), None if b is None else _ptr_array(b),
             _ptr_array(out), _int_array([wi.shape[0] for wi in w]), _int_array([wi.shape[1] for wi in w]), M, int(silu_in))
        outs += out
    return outs


def small_linear_many_bwd(dys, xs, ws, dxs, dws, dbs, silu_in=False, accumulate_dx=False):
    """Backward of small_linear_many.  dxs / dws / dbs: lists (entries may be None; dx entries may alias when
    accumulate_dx).  Batch-sized M goes through one launch per direction, tiny M through the per-layer kernels."""
    M = dys[0].shape[0]
    n = len(dys)
    dxs = dxs or [None] * n
    dws = dws or [None] * n
    dbs = dbs or [None] * n
    if M < 32:
        for i in range(n):
            small_linear_bwd(dys[i], xs[i], ws[i], dxs[i], dws[i], dbs[i], silu_in=silu_in, accumulate_dx=accumulate_dx)
        return
    for c in range(0, n, 16):
        sl = slice(c, c + 16)
        call("tsd_small_linear_many_bwd", len(dys[sl]), _ptr_array(dys[sl]), _ptr_array(xs[sl]), _ptr_array(ws[sl]),
             _ptr_array(dxs[sl]), _ptr_array(dws[sl]), _ptr_array(dbs[sl]), _int_array([w.shape[0] for w in ws[sl]]),
             _int_array([w.shape[1] for w in ws[sl]]), M, int(silu_in), int(accumulate_dx))


def timestep_embedding(t, freqs):
    M = t.shape[0]
    half = freqs.shape[0]
    emb = torch.empty(M, 2 * half, device=t.device, dtype=F32)
    call("tsd_timestep_embedding", _chk(t, torch.int64), _chk(freqs, F32), emb, M, half)
    return emb


def embedding_fwd(idx, table):
    M, D = idx.shape[0], table.shape[1]
    out = torch.empty(M, D, device=table.device, dtype=F32)
    call("tsd_embedding_fwd", _chk(idx, torch.int64), _chk(table, F32), out, M, D)
    return out


def embedding_bwd(idx, dy, dtable, padding_idx=0):
    call("tsd_embedding_bwd", idx, _chk(dy, F32), _chk(dtable, F32), idx.shape[0], dtable.shape[1], padding_idx)


# ------------------------------------------------------------------ image-side convs and DDPM process
def head_conv_fwd(x, w, bias):
    n, ci, H, W = x.shape
    co = w.shape[0]
    out = torch.empty(n * H * W, co, device=x.device, dtype=BF16)
    call("tsd_head_conv_fwd", _chk(x, F32), _chk(w, F32), bias, out, n, ci, H, W, co)
    return out


def im2col_head(x, KP=128):
    n, ci, H, W = x.shape
    patch = torch.empty(n * H * W, KP, device=x.device, dtype=BF16)
    call("tsd_im2col_head", _chk(x, F32), patch, n, ci, H, W, KP)
    return patch


def nchw_to_nhwc_pad(src, CP=64):
    n, co, H, W = src.shape
    out = torch.empty(n * H * W, CP, device=src.device, dtype=BF16)
    call("tsd_nchw_to_nhwc_pad", _chk(src, F32), out, n, co, H * W, CP)
    return out


def tail_conv_fwd(a, w, bias, n, H, W, out=None):
    co = w.shape[0]
    if out is None:
        out = torch.empty(n, co, H, W, device=a.device, dtype=F32)
    call("tsd_tail_conv_fwd", _chk(a, BF16), _chk(w, F32), bias, out, n, H, W, a.shape[1], co)
    return out


def tail_conv_dgrad(dy, a, w, n, H, W):
    da = torch.empty_like(a)
    call("tsd_tail_conv_dgrad", _chk(dy, F32), w, da, n, H, W, a.shape[1], w.shape[0])
    return da


def q_sample(x0, t, sqrt_ab, sqrt_1mab, seed=0, offset=0, noise=None, rng=None):
    n = x0.shape[0]
    x_t = torch.empty_like(x0)
    noise_out = torch.empty_like(x0) if noise is None else None
    call("tsd_q_sample", _chk(x0, F32), _chk(t, torch.int64), sqrt_ab, sqrt_1mab, noise, u64(seed), u64(offset), x_t,
         noise_out, n, i64(x0.numel() // n), rng)
    return x_t, (noise if noise is not None else noise_out)


def draw_timesteps(n, T, seed, rng, device):
    """t ~ U{0..T-1} per sample (utils.py:112) from the device Philox stream -> int64 [n]"""
    t = torch.empty(n, device=device, dtype=torch.int64)
    call("tsd_draw_timesteps", t, n, int(T), u64(seed), rng)
    return t


def counter_add_u64(counter, delta):
    call("tsd_counter_add_u64", _chk(counter, torch.int64), u64(delta))


def mse_fwd(pred, noise):
    loss = torch.empty_like(pred)
    call("tsd_mse_fwd", _chk(pred, F32), _chk(noise, F32), loss, i64(pred.numel()))
    return loss


def mse_bwd(pred, noise, gout):
    dpred = torch.empty_like(pred)
    call("tsd_mse_bwd", pred, noise, _chk(gout, F32), dpred, i64(pred.numel()))
    return dpred


def sampler_update(x, eps, step_ptr, c1, c2, sigma, w, x_out, nan_flag, noise=None, seed=0, clip_last=True, dup=False,
                   rng=None):
    total = eps.numel() // 2
    call("tsd_sampler_update", _chk(x, F32), _chk(eps, F32), step_ptr, c1, c2, sigma, f32(w), noise, u64(seed), x_out,
         nan_flag, i64(total), int(clip_last), int(dup), rng, eps.shape[0] // 2)


def tail_conv_sample(a, w, bias, x, B, H, W, step_ptr, c1, c2, sigma, wcfg, nan_flag, noise=None, seed=0, eps_out=None,
                     clip_last=True, rng=None):
    call("tsd_tail_conv_sample", _chk(a, BF16), _chk(w, F32), bias, _chk(x, F32), step_ptr, c1, c2, sigma, f32(wcfg), noise,
         u64(seed), nan_flag, eps_out, B, H, W, a.shape[1], w.shape[0], int(clip_last), rng)


def step_add(step_ptr, delta):
    call("tsd_step_add", step_ptr, delta)


def gather_row(table, step_ptr, out):
    call("tsd_gather_row_f32", _chk(table, F32), step_ptr, table.shape[1], out)


# ------------------------------------------------------------------ weight packing / optimiser
EPI_NONE, EPI_GEGLU, EPI_LRELU, EPI_RELU, EPI_TANH = range(5)  # TSD_EPI_* in the header
PACK_LINEAR, PACK_LINEAR_GEGLU, PACK_CONV3X3, PACK_CONV3X3_DGRAD, PACK_GEGLU_BIAS = range(5)  # TSD_PACK_* in the header


def pack_table(rows, device):
    """Device copy of the tsd_pack_many descriptor table; rows = [(src_ptr, dst_ptr, begin, rows, cols, kind)]."""
    import numpy as np
    dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("begin", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("kind", "<i4"),
                   ("pad", "<i4")])
    assert dt.itemsize == 40
    arr = np.zeros(len(rows), dtype=dt)
    for i, (s, d, b, r, c, k) in enumerate(rows):
        arr[i] = (s, d, b, r, c, k, 0)
    return torch.from_numpy(arr.view(np.uint8).copy()).to(device)


def pack_many(table, n_desc, total):
    call("tsd_pack_many", table, int(n_desc), i64(total))


def pack_linear(w, geglu=False):
    rows = w.shape[0]
    cols = w.numel() // rows
    out = torch.empty(rows, cols, device=w.device, dtype=BF16)
    call("tsd_pack_linear", _chk(w, F32), out, rows, cols, int(geglu))
    return out


def pack_geglu_bias(b):
    out = torch.empty_like(b)
    call("tsd_pack_geglu_bias", _chk(b, F32), out, b.shape[0])
    return out


def pack_conv3x3(w):
    co, ci = w.shape[:2]
    out = torch.empty(co, 9 * ci, device=w.device, dtype=BF16)
    call("tsd_pack_conv3x3", _chk(w, F32), out, co, ci)
    return out


def pack_conv3x3_dgrad(w):
    co, ci = w.shape[:2]
    out = torch.empty(ci, 9 * co, device=w.device, dtype=BF16)
    call("tsd_pack_conv3x3_dgrad", _chk(w, F32), out, co, ci)
    return out


def unpack_conv3x3_grad(src_packed, dst_oihw):
    co, ci = dst_oihw.shape[:2]
    call("tsd_unpack_conv3x3_grad", _chk(src_packed, F32), _chk(dst_oihw, F32), co, ci)


def add_cols(dst, src, rows, cols, ldd, lds):
    call("tsd_add_cols_f32", _chk(dst, F32), _chk(src, F32), rows, cols, ldd, lds)


def ema_update(ema, p, decay):
    call("tsd_ema_update", _chk(ema, F32), _chk(p, F32), i64(p.numel()), f32(decay), f32(1.0 - decay))


_sumsq_scratch = {}


def sumsq(g, out, scratch=None):
    """out[0] += sum g^2, reduced in a fixed order (bit-identical on every data-parallel rank)"""
    if scratch is None:
        scratch = _sumsq_scratch.get(g.device)
        if scratch is None:
            fn = _lib.lib().tsd_sumsq_scratch_floats
            fn.restype = __import__("ctypes").c_int64
            scratch = _sumsq_scratch[g.device] = torch.zeros(int(fn()), device=g.device, dtype=F32)
    call("tsd_sumsq_f32", _chk(g, F32), i64(g.numel()), out, scratch)


def adamw_clip(p, g, m, v, lr, beta1, beta2, eps, wd, step, max_norm, sumsq_buf, write_clipped_grad=True):
    call("tsd_adamw_clip", p, g, m, v, i64(p.numel()), f32(lr), f32(beta1), f32(beta2), f32(eps), f32(wd), int(step),
         f32(max_norm), sumsq_buf, int(write_clipped_grad))


def adamw_clip_dev(p, g, m, v, lr_dev, step_dev, beta1, beta2, eps, wd, max_norm, sumsq_buf, write_clipped_grad=True):
    """clip + AdamW with lr (fp32 [1]) and the step count (int32 [1], already incremented) read on the device"""
    call("tsd_adamw_clip_dev", p, g, m, v, i64(p.numel()), _chk(lr_dev, F32), _chk(step_dev, torch.int32), f32(beta1),
         f32(beta2), f32(eps), f32(wd), f32(max_norm), sumsq_buf, int(write_clipped_grad))


def gn_scratch_floats(n_img):
    fn = _lib.lib().tsd_gn_scratch_floats
    fn.restype = __import__("ctypes").c_int64
    return int(fn(int(n_img)))


# ------------------------------------------------------------------ image input / output (csrc/imageio.cu)
def _host_floats(vals):
    import ctypes
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def u8_to_f32_norm(img_u8_nhwc, mean, std):
    """uint8 [N,H,W,C] -> fp32 [N,C,H,W] = ((x / 255) - mean) / std  (ToTensor + Normalize, utils.py:21-25)."""
    if img_u8_nhwc.dtype != torch.uint8 or not img_u8_nhwc.is_cuda or not img_u8_nhwc.is_contiguous():
        raise RuntimeError("u8_to_f32_norm: expected a contiguous CUDA uint8 tensor [N,H,W,C]")
    N, H, W, C = img_u8_nhwc.shape
    out = torch.empty(N, C, H, W, device=img_u8_nhwc.device, dtype=F32)
    call("tsd_u8_to_f32_norm", img_u8_nhwc, out, N, C, H, W, _host_floats(mean), _host_floats(std))
    return out


def denorm_grid_u8(x, nrow, padding, mean, std):
    """fp32 [N,C,H,W] -> uint8 [GH,GW,C|3]: denormalize + make_grid + save_image's uint8 conversion."""
    N, C, H, W = x.shape
    if N == 1:
        padding = 0  # torchvision.utils.make_grid returns a single image unframed
    xmaps = min(nrow, N)
    ymaps = (N + xmaps - 1) // xmaps
    GH, GW = (H + padding) * ymaps + padding, (W + padding) * xmaps + padding
    out = torch.empty(GH, GW, 3 if C == 1 else C, device=x.device, dtype=torch.uint8)
    call("tsd_denorm_grid_u8", _chk(x, F32), out, N, C, H, W, int(nrow), int(padding), _host_floats(mean), _host_floats(std))
    return out
