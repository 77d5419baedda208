"""Kernel schedule of the UNet forward / backward (host side, Python like the reference).

Replaces Diffusion.forward (06_tiny_stable_diffusion/diffusion.py:263-276) and what torch autograd
derives from it (02_train_direct.py:71).  One autograd node spans the whole network: forward saves
the activations it needs in a plain Python record, backward walks the blocks in reverse and writes
parameter gradients straight into ``p.grad`` (views of one flat fp32 buffer).

Data layout in HBM: activations bf16 channels-last as [n*h*w, c]; GEMM weights packed bf16
[n_out][k] (3x3: k = tap*cin + ci); norm affine, biases, conditioning path and x_t / eps fp32.
"""
import math

import torch

from . import ops
from .rng import DeviceRng

F32 = torch.float32


class _ZeroArena:
    """Zero-initialised fp32 scratch of one backward pass, carved out of a single buffer that is cleared by ONE memset
    (instead of one ``torch.zeros`` per weight-gradient partial: ~100 small fills per step)."""

    def __init__(self):
        self.buf = None
        self.off = 0
        self.high = 0

    def begin(self, dev):
        want = max(self.high, 1 << 20)
        if self.buf is None or self.buf.device != dev or self.buf.numel() < want:
            self.buf = torch.zeros(want + want // 8, device=dev, dtype=F32)
        else:
            self.buf.zero_()
        self.off = 0
        self.high = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        n4 = (n + 3) // 4 * 4  # keep every slice 16-byte aligned
        self.high += n4
        if self.buf is not None and self.off + n4 <= self.buf.numel():
            out = self.buf[self.off:self.off + n].view(*shape)
            self.off += n4
            return out
        return torch.zeros(*shape, device=self.buf.device, dtype=F32)  # first pass only: the arena grows next time


class _Rec:
    """Saved tensors of one block (forward -> backward)."""
    __slots__ = ("kind", "key", "d")

    def __init__(self, kind, key, **d):
        self.kind, self.key, self.d = kind, key, d

    def __getattr__(self, k):
        try:
            return self.d[k]
        except KeyError:
            raise AttributeError(k)


class UNetEngine:
    def __init__(self, model):
        # weak-ish back reference (the engine is owned by the model)
        object.__setattr__(self, "_model_ref", model)
        self._packs = {}           # 'train' / 'infer' / 'dgrad' -> packed bf16 weights + descriptor table (persistent)
        self._epoch = 0
        self._generation = 0       # bumped whenever a cached device buffer (packed weights, scratch) is REPLACED:
        self._scratch = None       # captured CUDA graphs that baked the old pointers must be dropped (sampling.py)
        self._freqs = None
        self._flat_grad = None
        self._grad_views = None
        self._fn = None
        self._arena = _ZeroArena()
        # dropout masks: Philox keyed by (seed of the layer, call counter on the device, global sample index, element)
        self.rng = DeviceRng(salt=0xD20F)
        self.section_hook = None   # callable(section) fired by backward() when a parameter section's grads are final
        # training GEGLU without the 8C-wide pre-activation tensor: forward through the fused epilogue, backward by
        # recomputing the pre-activations inside the GEMM that applies the activation gradient (TSD_GEGLU_FUSED_TRAIN=0:
        # store h8 and run the stand-alone geglu kernels)
        import os
        self.fused_geglu_train = bool(int(os.environ.get("TSD_GEGLU_FUSED_TRAIN", "1")))

    @property
    def seed(self):
        return self.rng.seed

    @seed.setter
    def seed(self, v):
        self.rng.seed = int(v)

    # ------------------------------------------------------------------ parameters
    @property
    def model(self):
        return self._model_ref

    def invalidate(self):
        self._packs = {}
        self._flat_grad = None
        self._grad_views = None
        self._scratch = None
        self._freqs = None
        self._generation += 1

    def bump(self):
        """Call after parameters were modified through raw pointers (fused optimiser)."""
        self._epoch += 1

    def params(self):
        return dict(self.model.named_parameters())

    def _signature(self, P):
        return (self._epoch,) + tuple((p.data_ptr(), p._version) for p in P.values())

    def _block_list(self):
        m = self.model
        out = []
        for i, st in enumerate(m._enc):
            for j, b in enumerate(st):
                out.append((f"encoders.{i}.{j}", b))
        for j, b in enumerate(m._mid):
            out.append((f"bottleneck.{j}", b))
        for i, st in enumerate(m._dec):
            for j, b in enumerate(st):
                out.append((f"decoders.{i}.{j}", b))
        return out

    # ------------------------------------------------------------------ packed bf16 weights
    def _pack_entries(self, P, which):
        """[(name, source parameter, kind, rows, cols, dst shape, dst dtype)] of one packing set.
        'train' / 'infer': forward GEMM weights (the C -> 8C linear of the attention blocks is packed plain when the
        tape is saved -- h8 is kept for the backward -- and with interleaved value / gate rows plus a permuted bias for
        the fused GEGLU epilogue otherwise); 'dgrad': data-gradient weights of the 3x3 convs ([ci][mirrored tap][co])."""
        BF = torch.bfloat16
        E = []

        def lin(name, key, kind=ops.PACK_LINEAR):
            w = P[key]
            rows = w.shape[0]
            E.append((name, w, kind, rows, w.numel() // rows, (rows, w.numel() // rows), BF))

        def conv(name, key, kind):
            w = P[key]
            co, ci = w.shape[:2]
            shape = (co, 9 * ci) if kind == ops.PACK_CONV3X3 else (ci, 9 * co)
            E.append((name, w, kind, co, ci, shape, BF))

        ck = ops.PACK_CONV3X3_DGRAD if which == "dgrad" else ops.PACK_CONV3X3
        for key, b in self._block_list():
            if b[0] == "conv":
                if b[1] % 64 == 0:  # the head conv (3/4 input channels) runs on CUDA cores from fp32
                    conv(key, key + ".weight", ck)
            elif b[0] == "up":
                conv(key + ".conv", key + ".conv.weight", ck)
            elif b[0] == "res":
                conv(key + ".conv_1.2", key + ".conv_1.2.weight", ck)
                conv(key + ".conv_2.3", key + ".conv_2.3.weight", ck)
                if b[1] != b[2] and which != "dgrad":
                    lin(key + ".residual_layer", key + ".residual_layer.weight")
            elif which != "dgrad":
                for n in ("conv_1.1", "atten_1.1.in_proj", "atten_1.1.out_proj", "linear_2", "conv_output"):
                    lin(key + "." + n, f"{key}.{n}.weight")
                if which == "train":  # plain: the B operand of the data gradient dl3 = dh8 W1
                    lin(key + ".linear_1", key + ".linear_1.weight")
                if which == "infer" or self._geglu_fused(b[1]):
                    lin(key + ".linear_1.geglu", key + ".linear_1.weight", ops.PACK_LINEAR_GEGLU)
                    b1 = P[key + ".linear_1.bias"]
                    E.append((key + ".linear_1.geglu_bias", b1, ops.PACK_GEGLU_BIAS, b1.shape[0], 1, (b1.shape[0],), F32))
        return E

    def _geglu_fused(self, C):
        """training GEGLU through the fused forward epilogue + recomputing backward GEMM (8C <= 2048: its per-thread
        bias-gradient partials cover 16 n-blocks); wider blocks keep the stored pre-activations"""
        return self.fused_geglu_train and 8 * C <= 2048

    def _pack_set(self, P, which):
        """Packed copies of one set, refreshed by ONE kernel launch when any parameter changed.  The destination buffers
        and the descriptor table are allocated once and never move (captured CUDA graphs keep using them)."""
        st = self._packs.get(which)
        sig = self._signature(P)
        if st is not None and sig != st["sig"] and st["ptrs"] != tuple(w.data_ptr() for w in st["srcs"]):
            st = None  # a parameter was re-homed behind our back: the descriptor table points at the old storage
        if st is None:
            E = self._pack_entries(P, which)
            dev = E[0][1].device
            W, rows, begin = {}, [], 0
            for name, w, kind, r, c, shape, dt in E:
                W[name] = torch.empty(shape, device=dev, dtype=dt)
                rows.append((w.data_ptr(), W[name].data_ptr(), begin, r, c, kind))
                begin += W[name].numel()
            st = dict(W=W, table=ops.pack_table(rows, dev), n=len(rows), total=begin, sig=None,
                      srcs=[w for _, w, *_ in E], ptrs=tuple(w.data_ptr() for _, w, *_ in E))
            self._packs[which] = st
            self._generation += 1
        if sig != st["sig"]:
            ops.pack_many(st["table"], st["n"], st["total"])
            st["sig"] = sig
        return st["W"]

    def packed(self, P, save=False):
        return self._pack_set(P, "train" if save else "infer")

    def packed_dgrad(self, P):
        return self._pack_set(P, "dgrad")

    def _get_scratch(self, n_img, dev):
        if self._scratch is None or self._scratch_imgs < n_img or self._scratch.device != dev:
            need = ops.gn_scratch_floats(n_img)  # per-CTA GroupNorm partials, sized by the library for this device
            self._scratch = torch.zeros(max(need, 4096), device=dev, dtype=F32)
            self._scratch_imgs = n_img
            self._generation += 1
        return self._scratch

    def _get_freqs(self, dev):
        if self._freqs is None or self._freqs.device != dev:
            half = self.model.d_model // 2
            # the reference's own expression (diffusion.py:26), evaluated on the host in fp32
            self._freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half) / half).to(dev)
        return self._freqs

    # ------------------------------------------------------------------ autograd entry
    def apply(self, x, t, labels):
        P = self.params()
        plist = list(P.values())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
        if not need_grad:
            with torch.no_grad():
                eps, _ = self.forward(x, t, labels, save=False)
            return eps
        return _UNetFn.apply(self, x, t, labels, *plist)

    # ------------------------------------------------------------------ conditioning (diffusion.py:265-266)
    def conditioning(self, P, t, labels, save):
        m = self.model
        rec = {}
        t_freq = ops.timestep_embedding(t, self._get_freqs(t.device))
        h1 = ops.small_linear(t_freq, P["time_embedding.mlp.0.weight"], P["time_embedding.mlp.0.bias"])
        temb = ops.small_linear(h1, P["time_embedding.mlp.2.weight"], P["time_embedding.mlp.2.bias"], silu_in=True)
        emb = ops.embedding_fwd(labels, P["label_embedding.0.weight"])
        c1 = ops.small_linear(emb, P["label_embedding.1.weight"], P["label_embedding.1.bias"])
        ctx = ops.small_linear(c1, P["label_embedding.3.weight"], P["label_embedding.3.bias"], silu_in=True)
        if save:
            rec.update(t_freq=t_freq, h1=h1, temb=temb, emb=emb, c1=c1, ctx=ctx, labels=labels)
        return temb, ctx, rec

    def time_bias(self, P, key, temb):
        """linear_time = Linear(SiLU(temb)) (diffusion.py:101-104, 112) -> [B, Cout] fp32"""
        return ops.small_linear(temb, P[key + ".linear_time.1.weight"], P[key + ".linear_time.1.bias"], silu_in=True)

    def cross_bias(self, P, key, ctx):
        """CrossAttention with one key/value token == out_proj(v_proj(ctx)) per sample (diffusion.py:71-82)"""
        vv = ops.small_linear(ctx, P[key + ".atten_2.v_proj.weight"])
        cb = ops.small_linear(vv, P[key + ".atten_2.out_proj.weight"], P[key + ".atten_2.out_proj.bias"])
        return vv, cb

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, labels, save, tb_override=None, cb_override=None, eps_out=None, taps=None, sample_tail=None,
                shared_prefix=False):
        """Returns (eps fp32 NCHW, tape).  tb_override / cb_override: precomputed per-block conditioning
        (sampling: one shared time row for the whole batch, constant label vectors).

        shared_prefix (sampling with classifier-free guidance): rows [0, n/2) and [n/2, n) hold the same x_t and differ
        only in the label, and the label first enters through the cross-attention vector of the first attention block
        (diffusion.py:141-148).  Everything before it -- head conv, first ResBlock, and that block's GroupNorm,
        conv_1, LayerNorm, in_proj and the L = H*W self-attention -- is computed once on n/2 rows and duplicated."""
        m = self.model
        P = self.params()
        W = self.packed(P, save)
        n, ci, H, Wd = x.shape
        assert ci == m.channel_img
        dev = x.device
        scratch = self._get_scratch(n, dev)
        training = m.training
        tape = []
        x = x.contiguous().float()
        n_full = n
        drop_pos = None
        if training and m.dropout > 0.0:
            # one stream position per forward, snapshotted so that the backward regenerates the same masks whatever
            # happens to the counter in between (a second forward before the backward, graph replays)
            drop_pos = self.rng.advance(dev).clone()
        layer_no = [0]
        first_attn = None
        if shared_prefix:
            assert not save and tb_override is not None and cb_override is not None and n % 2 == 0
            first_attn = next((f"encoders.{i}.{j}" for i, st in enumerate(m._enc) for j, b in enumerate(st)
                               if b[0] == "attn"), None)
            if first_attn is not None:
                n = n_full // 2  # rows computed until the label enters
        skips = []

        if tb_override is None:
            temb, ctx, crec = self.conditioning(P, t.contiguous(), labels.contiguous(), save)
            # every per-block conditioning vector of this forward in three launches: the 14 linear_time rows, the 10
            # cross-attention v_proj and out_proj vectors (separate parameters, same batch)
            blocks = self._block_list()
            rkeys = [k for k, b in blocks if b[0] == "res"]
            akeys = [k for k, b in blocks if b[0] == "attn"]
            tbs = dict(zip(rkeys, ops.small_linear_many([temb] * len(rkeys), [P[k + ".linear_time.1.weight"] for k in rkeys],
                                                        [P[k + ".linear_time.1.bias"] for k in rkeys], silu_in=True)))
            vvs = ops.small_linear_many([ctx] * len(akeys), [P[k + ".atten_2.v_proj.weight"] for k in akeys])
            cbs = ops.small_linear_many(vvs, [P[k + ".atten_2.out_proj.weight"] for k in akeys],
                                        [P[k + ".atten_2.out_proj.bias"] for k in akeys])
            vvs, cbs = dict(zip(akeys, vvs)), dict(zip(akeys, cbs))
        else:
            temb = ctx = None
            crec = {}

        def run_res(key, b, x0, x1, h, w):
            ci_, co = b[1], b[2]
            hw = h * w
            if tb_override is None:
                tb, rps = tbs[key], hw
            else:
                tb, rps = tb_override[key]
            p_drop = m.dropout if (b[3] and training) else 0.0
            st1 = ops.gn_stats(x0, n, hw, 1e-5, scratch, x1=x1)
            a1 = ops.gn_apply(x0, n, hw, st1, P[key + ".conv_1.0.weight"], P[key + ".conv_1.0.bias"], True, x1=x1)
            # row_bias row = output row / rps: per image in training, one shared time row when sampling
            # (gn=True: the epilogue also leaves the GroupNorm partial sums of its output behind -- the statistics of the
            # next GroupNorm then need no pass over the tensor)
            hmid = ops.conv3x3(a1, n, h, w, W[key + ".conv_1.2"], co, bias=P[key + ".conv_1.2.bias"],
                               row_bias=tb, rows_per_sample=rps, gn=True)
            st2 = ops.gn_stats(hmid, n, hw, 1e-5, scratch)
            seed = 0
            layer_no[0] += 1
            if p_drop > 0.0:
                seed = (self.rng.seed * 1000003 + layer_no[0]) & 0xFFFFFFFFFFFFFFFF
            a2 = ops.gn_apply(hmid, n, hw, st2, P[key + ".conv_2.0.weight"], P[key + ".conv_2.0.bias"], True,
                              drop_p=p_drop, seed=seed, rng=drop_pos)
            if ci_ != co:
                sc = ops.gemm(x0, W[key + ".residual_layer"], co, a1=x1, bias=P[key + ".residual_layer.bias"])
            else:
                sc = x0
            out = ops.conv3x3(a2, n, h, w, W[key + ".conv_2.3"], co, bias=P[key + ".conv_2.3.bias"], residual=sc, gn=True)
            if save:
                tape.append(_Rec("res", key, b=b, x0=x0, x1=x1, st1=st1, a1=a1, hmid=hmid, st2=st2, a2=a2,
                                 p_drop=p_drop, seed=seed, pos=drop_pos, h=h, w=w))
            return out

        def run_attn(key, b, x0, h, w):
            nonlocal n
            C = b[1]
            L = h * w
            if cb_override is None:
                vv, cb = vvs[key], cbs[key]
            else:
                vv, cb = None, cb_override[key]
            st = ops.gn_stats(x0, n, L, 1e-6, scratch)
            g = ops.gn_apply(x0, n, L, st, P[key + ".conv_1.0.weight"], P[key + ".conv_1.0.bias"], False)
            t0 = ops.gemm(g, W[key + ".conv_1.1"], C, bias=P[key + ".conv_1.1.bias"])
            l1 = ops.ln_fwd(t0, P[key + ".atten_1.0.weight"], P[key + ".atten_1.0.bias"])
            qkv = ops.gemm(l1, W[key + ".atten_1.1.in_proj"], 3 * C)
            o, lse = ops.attn_fwd(qkv, n, L, C, m.N_HEAD, need_lse=save)
            widen = key == first_attn
            if widen:
                # the cross-attention vector is the first label-dependent term: from here on all n_full rows exist.  o,
                # t0 and the block input x0 are not duplicated: the two halves of t2 (and of the block output below) are
                # produced by two GEMM calls that read the same n rows; only the skip tensors are copied
                nh = n
                n = n_full
                skips[:] = [torch.cat([sk, sk]) for sk in skips]
                t2 = ops.empty_bf16(n * L, C, like=o)
                for hf in range(2):
                    ops.gemm(o, W[key + ".atten_1.1.out_proj"], C, bias=P[key + ".atten_1.1.out_proj.bias"],
                             row_bias=cb[hf * nh:(hf + 1) * nh], rows_per_sample=L, residual=t0,
                             out=t2[hf * nh * L:(hf + 1) * nh * L])
            else:
                t2 = ops.gemm(o, W[key + ".atten_1.1.out_proj"], C, bias=P[key + ".atten_1.1.out_proj.bias"],
                              row_bias=cb, rows_per_sample=L, residual=t0)
            l3 = ops.ln_fwd(t2, P[key + ".norm_3.weight"], P[key + ".norm_3.bias"])
            if save and not self._geglu_fused(C):
                h8 = ops.gemm(l3, W[key + ".linear_1"], 8 * C, bias=P[key + ".linear_1.bias"])
                gg = ops.geglu_fwd(h8)
            else:
                h8 = None
                gg = ops.gemm(l3, W[key + ".linear_1.geglu"], 8 * C, bias=W[key + ".linear_1.geglu_bias"], geglu=True)
            t3 = ops.gemm(gg, W[key + ".linear_2"], C, bias=P[key + ".linear_2.bias"], residual=t2)
            if widen:  # both halves add the same (un-duplicated) block input
                out = ops.empty_bf16(n * L, C, like=t3)
                part = ops._gn_part_for(out, n * L, C) if L % 64 == 0 else None
                for hf in range(2):
                    rows = slice(hf * nh * L, (hf + 1) * nh * L)
                    ops.gemm(t3[rows], W[key + ".conv_output"], C, bias=P[key + ".conv_output.bias"], residual=x0,
                             out=out[rows], gn_part=None if part is None else part[hf * nh * L // 32:(hf + 1) * nh * L // 32])
            else:
                out = ops.gemm(t3, W[key + ".conv_output"], C, bias=P[key + ".conv_output.bias"], residual=x0, gn=True)
            if save:
                tape.append(_Rec("attn", key, b=b, x0=x0, st=st, g=g, t0=t0, l1=l1, qkv=qkv, o=o, lse=lse, t2=t2, l3=l3,
                                 h8=h8, gg=gg, t3=t3, vv=vv, h=h, w=w))
            return out

        def run_block(key, b, x0, x1, h, w):
            """returns (out, h, w)"""
            if b[0] == "conv":
                if b[1] % 64 != 0:  # head conv on the fp32 NCHW image
                    out = ops.head_conv_fwd(x if n == n_full else x[:n], P[key + ".weight"], P[key + ".bias"])
                    if save:
                        tape.append(_Rec("head", key, b=b))
                    return out, h, w
                out = ops.conv3x3(x0, n, h, w, W[key], b[2], stride=b[3], bias=P[key + ".bias"], gn=True)
                if save:
                    tape.append(_Rec("conv", key, b=b, x0=x0, h=h, w=w))
                return out, h // b[3], w // b[3]
            if b[0] == "up":
                u = ops.upsample2_fwd(x0, n, h, w)
                out = ops.conv3x3(u, n, 2 * h, 2 * w, W[key + ".conv"], b[1], bias=P[key + ".conv.bias"], gn=True)
                if save:
                    tape.append(_Rec("up", key, b=b, u=u, h=h, w=w))
                return out, 2 * h, 2 * w
            if b[0] == "res":
                return run_res(key, b, x0, x1, h, w), h, w
            return run_attn(key, b, x0, h, w), h, w

        h, w = H, Wd
        cur = None
        for i, st in enumerate(m._enc):
            for j, b in enumerate(st):
                cur, h, w = run_block(f"encoders.{i}.{j}", b, cur, None, h, w)
                if taps is not None:
                    taps[f"encoders.{i}.{j}"] = (cur, h, w)
            skips.append(cur)
            if save:
                tape.append(_Rec("skip_push", ""))
        for j, b in enumerate(m._mid):
            cur, h, w = run_block(f"bottleneck.{j}", b, cur, None, h, w)
            if taps is not None:
                taps[f"bottleneck.{j}"] = (cur, h, w)
        for i, st in enumerate(m._dec):
            skip = skips.pop()
            if save:
                tape.append(_Rec("skip_pop", ""))
            for j, b in enumerate(st):
                cur, h, w = run_block(f"decoders.{i}.{j}", b, cur, skip if j == 0 else None, h, w)
                if taps is not None:
                    taps[f"decoders.{i}.{j}"] = (cur, h, w)
        # tail: GroupNorm -> SiLU -> conv3x3 C -> channel_img (diffusion.py:257-261)
        hw = h * w
        stt = ops.gn_stats(cur, n, hw, 1e-5, scratch)
        at = ops.gn_apply(cur, n, hw, stt, P["tail.0.weight"], P["tail.0.bias"], True)
        if sample_tail is not None:
            # sampling: final conv + CFG combine + posterior update + Philox noise in one kernel (x updated in place)
            ops.tail_conv_sample(at, P["tail.2.weight"], P["tail.2.bias"], x, n // 2, h, w, **sample_tail)
            return None, None
        eps = ops.tail_conv_fwd(at, P["tail.2.weight"], P["tail.2.bias"], n, h, w, out=eps_out)
        if save:
            tape.append(_Rec("tail", "tail", xin=cur, st=stt, a=at, h=h, w=w))
            return eps, dict(tape=tape, crec=crec, x=x, n=n, P=P, W=W, scratch=scratch)
        return eps, None

    # ------------------------------------------------------------------ gradients
    def _grad_buffers(self, P):
        """p.grad tensors as views into one flat fp32 buffer (zeroed when freshly attached)."""
        plist = list(P.values())
        dev = plist[0].device
        total = sum(p.numel() for p in plist)
        if self._flat_grad is None or self._flat_grad.device != dev or self._flat_grad.numel() != _aligned_total(plist):
            self._flat_grad = torch.zeros(_aligned_total(plist), device=dev, dtype=F32)
            views, off = {}, 0
            for k, p in P.items():
                views[k] = self._flat_grad[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
            self._grad_views = views
        G = {}
        fresh = all(p.grad is None for p in plist)
        if fresh:
            self._flat_grad.zero_()
        for k, p in P.items():
            v = self._grad_views[k]
            if p.grad is None:
                if not fresh:
                    v.zero_()
                p.grad = v
                G[k] = v
            elif p.grad.data_ptr() == v.data_ptr():
                G[k] = v
            else:  # foreign grad tensor: accumulate through a temporary
                G[k] = torch.zeros_like(p)
        return G, total

    def _bias_grad(self, dy, n_samples, rows_per_sample, db):
        """db[c] += column sums of dy; returns the per-sample sums (their zeroed destination comes from the arena)."""
        return ops.colsum(dy, n_samples, rows_per_sample, total=db, out=self._arena.take(n_samples, dy.shape[1]))

    def backward(self, saved, deps):
        m = self.model
        P, W, n, scratch = saved["P"], saved["W"], saved["n"], saved["scratch"]
        x = saved["x"]
        tape = saved["tape"]
        crec = saved["crec"]
        G, _ = self._grad_buffers(P)
        D = self.packed_dgrad(P)
        saved["D"] = D
        dev = x.device
        arena = self._arena
        arena.begin(dev)
        d_temb = arena.take(n, m.time_emb_dim)
        d_ctx = arena.take(n, m.time_emb_dim)
        temb, ctx = crec["temb"], crec["ctx"]
        packed_wgrads = []  # (packed fp32 grad, OIHW grad view) to unpack at the end

        def conv_wgrad(dy, x0, key_w, hh, ww, x1=None, stride=1):
            gw = G[key_w]
            tmp = arena.take(gw.shape[0], 9 * gw.shape[1])
            ops.conv3x3_wgrad(dy, x0, n, hh, ww, tmp, x1=x1, stride=stride)
            ops.unpack_conv3x3_grad(tmp, gw)

        dcur = None
        skip_grads = []
        pend_time, pend_cross = [], []  # (dtb, key) / (dcb, vv, key): conditioning backward, batched per section

        def flush_conditioning():
            if pend_time:
                ks = [k for _, k in pend_time]
                ops.small_linear_many_bwd([d for d, _ in pend_time], [temb] * len(ks),
                                          [P[k + ".linear_time.1.weight"] for k in ks], [d_temb] * len(ks),
                                          [G[k + ".linear_time.1.weight"] for k in ks],
                                          [G[k + ".linear_time.1.bias"] for k in ks], silu_in=True, accumulate_dx=True)
                pend_time.clear()
            if pend_cross:
                ks = [k for _, _, k in pend_cross]
                dvvs = [torch.empty_like(v) for _, v, _ in pend_cross]
                # cb = out_proj(v_proj(ctx)) + bias   (norm_2, q_proj, k_proj get exact zeros)
                ops.small_linear_many_bwd([d for d, _, _ in pend_cross], [v for _, v, _ in pend_cross],
                                          [P[k + ".atten_2.out_proj.weight"] for k in ks], dvvs,
                                          [G[k + ".atten_2.out_proj.weight"] for k in ks],
                                          [G[k + ".atten_2.out_proj.bias"] for k in ks])
                ops.small_linear_many_bwd(dvvs, [ctx] * len(ks), [P[k + ".atten_2.v_proj.weight"] for k in ks],
                                          [d_ctx] * len(ks), [G[k + ".atten_2.v_proj.weight"] for k in ks], None,
                                          accumulate_dx=True)
                pend_cross.clear()

        section = "decoders"  # the tail sits right behind the decoders in the flat gradient buffer
        for rec in reversed(tape):
            kind, key = rec.kind, rec.key
            if key:
                sec = key.split(".", 1)[0]
                if sec in ("bottleneck", "encoders") and sec != section:
                    flush_conditioning()  # the linear_time / cross-attention gradients of the section just left
                    if self.section_hook is not None:
                        self.section_hook(section)  # every gradient of that section is final now
                    section = sec
            if kind == "tail":
                hh, ww = rec.h, rec.w
                # tail conv backward: data gradient on CUDA cores (N = 3), weight gradient on the tensor cores through
                # the generic 3x3 wgrad path with d(eps) padded to 64 channels
                deps_c = deps.contiguous().float()
                da = ops.tail_conv_dgrad(deps_c, rec.a, P["tail.2.weight"], n, hh, ww)
                co_img = P["tail.2.weight"].shape[0]
                dyp = ops.nchw_to_nhwc_pad(deps_c, 64)
                tmpw = arena.take(64, 9 * rec.a.shape[1])
                ops.conv3x3_wgrad(dyp, rec.a, n, hh, ww, tmpw)
                ops.unpack_conv3x3_grad(tmpw[:co_img], G["tail.2.weight"])
                tmpb = arena.take(64)
                self._bias_grad(dyp, n, hh * ww, tmpb)
                ops.add_cols(G["tail.2.bias"], tmpb, 1, co_img, co_img, 64)
                dcur, _ = ops.gn_bwd(da, rec.xin, n, hh * ww, rec.st, P["tail.0.weight"], P["tail.0.bias"], True,
                                     G["tail.0.weight"], G["tail.0.bias"])
            elif kind == "skip_pop":
                # forward popped a skip here: the block list after it consumed cat(cur, skip); dcur currently
                # holds (dx0, dx1) from that stage's first ResBlock
                dcur, dskip = dcur
                skip_grads.append(dskip)
            elif kind == "skip_push":
                # forward pushed `cur` as a skip: its gradient gains the matching decoder's concat gradient
                dskip = skip_grads.pop()
                dcur = ops.add(dcur, dskip) if dcur is not None else dskip
            elif kind == "res":
                dcur = self._res_bwd(rec, dcur, P, W, G, n, pend_time, conv_wgrad, D)
            elif kind == "attn":
                dcur = self._attn_bwd(rec, dcur, P, W, G, n, pend_cross)
            elif kind == "conv":
                b, hh, ww = rec.b, rec.h, rec.w
                s = b[3]
                self._bias_grad(dcur, n, (hh // s) * (ww // s), G[key + ".bias"])
                conv_wgrad(dcur, rec.x0, key + ".weight", hh, ww, stride=s)
                if s == 2:
                    zs = ops.zero_stuff2(dcur, n, hh // 2, ww // 2)
                    dcur = ops.conv3x3(zs, n, hh, ww, D[key], b[1])
                else:
                    dcur = ops.conv3x3(dcur, n, hh, ww, D[key], b[1])
            elif kind == "up":
                hh, ww = rec.h, rec.w
                self._bias_grad(dcur, n, 4 * hh * ww, G[key + ".conv.bias"])
                conv_wgrad(dcur, rec.u, key + ".conv.weight", 2 * hh, 2 * ww)
                du = ops.conv3x3(dcur, n, 2 * hh, 2 * ww, D[key + ".conv"], rec.b[1])
                dcur = ops.upsample2_bwd(du, n, hh, ww)
            elif kind == "head":
                # head conv weight gradient on the tensor cores: dW[co][c*9+tap] = dY^T * im2col(x)
                gw = G[key + ".weight"]
                patch = ops.im2col_head(x, 128)
                tmpw = arena.take(gw.shape[0], 128)
                ops.gemm_wgrad(dcur, patch, tmpw)
                ops.add_cols(gw, tmpw, gw.shape[0], gw.shape[1] * 9, gw.shape[1] * 9, 128)
                self._bias_grad(dcur, n, dcur.shape[0] // n, G[key + ".bias"])
                dcur = None
        assert not skip_grads
        flush_conditioning()
        # conditioning backward (time / label MLPs, embedding)
        dh1 = torch.empty_like(crec["h1"])
        ops.small_linear_bwd(d_temb, crec["h1"], P["time_embedding.mlp.2.weight"], dh1, G["time_embedding.mlp.2.weight"],
                             G["time_embedding.mlp.2.bias"], silu_in=True)
        ops.small_linear_bwd(dh1, crec["t_freq"], P["time_embedding.mlp.0.weight"], None, G["time_embedding.mlp.0.weight"],
                             G["time_embedding.mlp.0.bias"])
        dc1 = torch.empty_like(crec["c1"])
        ops.small_linear_bwd(d_ctx, crec["c1"], P["label_embedding.3.weight"], dc1, G["label_embedding.3.weight"],
                             G["label_embedding.3.bias"], silu_in=True)
        demb = torch.empty_like(crec["emb"])
        ops.small_linear_bwd(dc1, crec["emb"], P["label_embedding.1.weight"], demb, G["label_embedding.1.weight"],
                             G["label_embedding.1.bias"])
        ops.embedding_bwd(crec["labels"], demb, G["label_embedding.0.weight"], padding_idx=0)
        # foreign grad tensors
        for k, p in P.items():
            if p.grad is not None and p.grad.data_ptr() != G[k].data_ptr():
                p.grad.add_(G[k])
        if self.section_hook is not None:
            self.section_hook("rest")  # encoders + the conditioning MLPs: everything that is left

    def _res_bwd(self, rec, dout, P, W, G, n, pend_time, conv_wgrad, D):
        key, b = rec.key, rec.b
        ci, co = b[1], b[2]
        hh, ww = rec.h, rec.w
        hw = hh * ww
        # conv_2 (+ shortcut bias share the same column sums of dout)
        per = self._bias_grad(dout, n, hw, G[key + ".conv_2.3.bias"])
        conv_wgrad(dout, rec.a2, key + ".conv_2.3.weight", hh, ww)
        da2 = ops.conv3x3(dout, n, hh, ww, D[key + ".conv_2.3"], co)
        # time bias: per-sample column sums of dh feed linear_time; their sum over samples is conv_1's bias grad -- both
        # come out of the GroupNorm backward kernel that writes dh
        dtb = self._arena.take(n, co)
        dh, _ = ops.gn_bwd(da2, rec.hmid, n, hw, rec.st2, P[key + ".conv_2.0.weight"], P[key + ".conv_2.0.bias"], True,
                           G[key + ".conv_2.0.weight"], G[key + ".conv_2.0.bias"], drop_p=rec.p_drop, seed=rec.seed, rng=rec.pos,
                           colsum_out=dtb, colsum_total=G[key + ".conv_1.2.bias"])
        pend_time.append((dtb, key))  # linear_time backward: batched with the other blocks of this section
        conv_wgrad(dh, rec.a1, key + ".conv_1.2.weight", hh, ww)
        da1 = ops.conv3x3(dh, n, hh, ww, D[key + ".conv_1.2"], ci)
        if ci != co:
            ops.reduce_rows_into(per, G[key + ".residual_layer.bias"])
            ops.gemm_wgrad(dout, rec.x0, G[key + ".residual_layer.weight"].view(co, ci), x1=rec.x1)
            radd = ops.gemm_dgrad(dout, W[key + ".residual_layer"], ci)
        else:
            radd = dout
        dx0, dx1 = ops.gn_bwd(da1, rec.x0, n, hw, rec.st1, P[key + ".conv_1.0.weight"], P[key + ".conv_1.0.bias"], True,
                              G[key + ".conv_1.0.weight"], G[key + ".conv_1.0.bias"], x1=rec.x1, radd=radd)
        return (dx0, dx1) if rec.x1 is not None else dx0

    def _attn_bwd(self, rec, dout, P, W, G, n, pend_cross):
        key, b = rec.key, rec.b
        C = b[1]
        L = rec.h * rec.w
        k = key
        # conv_output (1x1) + long residual
        self._bias_grad(dout, n, L, G[k + ".conv_output.bias"])
        ops.gemm_wgrad(dout, rec.t3, G[k + ".conv_output.weight"].view(C, C))
        dt3 = ops.gemm_dgrad(dout, W[k + ".conv_output"], C)
        # linear_2 (+ short residual to t2)
        self._bias_grad(dt3, n, L, G[k + ".linear_2.bias"])
        ops.gemm_wgrad(dt3, rec.gg, G[k + ".linear_2.weight"])
        dgg = ops.gemm_dgrad(dt3, W[k + ".linear_2"], 4 * C)
        if rec.h8 is None:  # pre-activations recomputed inside the GEMM whose epilogue applies the activation gradient
            dh8 = ops.gemm_geglu_bwd(rec.l3, W[k + ".linear_1.geglu"], W[k + ".linear_1.geglu_bias"], dgg,
                                     dbias=G[k + ".linear_1.bias"])
        else:
            dh8 = ops.geglu_bwd(rec.h8, dgg, dbias=G[k + ".linear_1.bias"])  # bias gradient as a by-product
        ops.gemm_wgrad(dh8, rec.l3, G[k + ".linear_1.weight"])
        dl3 = ops.gemm_dgrad(dh8, W[k + ".linear_1"], C)
        # out_proj (+ cross-attention vector + residual t0): the per-sample column sums of dt2 are the gradient of the
        # cross-attention vector, their total the out_proj bias gradient -- by-products of the LayerNorm backward
        dcb = self._arena.take(n, C)
        dt2 = ops.ln_bwd(dl3, rec.t2, P[k + ".norm_3.weight"], G[k + ".norm_3.weight"], G[k + ".norm_3.bias"], radd=dt3,
                         rows_per_sample=L, colsum_out=dcb, colsum_total=G[k + ".atten_1.1.out_proj.bias"])
        ops.gemm_wgrad(dt2, rec.o, G[k + ".atten_1.1.out_proj.weight"])
        do = ops.gemm_dgrad(dt2, W[k + ".atten_1.1.out_proj"], C)
        dqkv = ops.attn_bwd(rec.qkv, rec.o, do, rec.lse, n, L, C, self.model.N_HEAD)
        ops.gemm_wgrad(dqkv, rec.l1, G[k + ".atten_1.1.in_proj.weight"])
        dl1 = ops.gemm_dgrad(dqkv, W[k + ".atten_1.1.in_proj"], C)
        dt0 = ops.ln_bwd(dl1, rec.t0, P[k + ".atten_1.0.weight"], G[k + ".atten_1.0.weight"], G[k + ".atten_1.0.bias"],
                         radd=dt2, rows_per_sample=L, colsum_total=G[k + ".conv_1.1.bias"])
        # conv_1.1 (1x1) and GroupNorm (eps 1e-6, no activation), + long residual
        ops.gemm_wgrad(dt0, rec.g, G[k + ".conv_1.1.weight"].view(C, C))
        dg = ops.gemm_dgrad(dt0, W[k + ".conv_1.1"], C)
        dx, _ = ops.gn_bwd(dg, rec.x0, n, L, rec.st, P[k + ".conv_1.0.weight"], P[k + ".conv_1.0.bias"], False,
                           G[k + ".conv_1.0.weight"], G[k + ".conv_1.0.bias"], radd=dout)
        pend_cross.append((dcb, rec.vv, k))  # degenerate cross attention backward: batched per section
        return dx


def _aligned_total(plist):
    return sum((p.numel() + 3) // 4 * 4 for p in plist)


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, t, labels, *params):
        eps, saved = engine.forward(x, t, labels, save=True)
        ctx.engine = engine
        ctx.saved = saved
        return eps

    @staticmethod
    def backward(ctx, deps):
        ctx.engine.backward(ctx.saved, deps)
        ctx.saved = None
        return (None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 4)
