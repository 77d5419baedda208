"""Reverse-diffusion loop of SamplerDDPM as a replayed CUDA graph.

Reference: SamplerDDPM.forward / p_mean_variance, 06_tiny_stable_diffusion/utils.py:147-171.  What
changes versus the reference's loop, not its arithmetic:
  * the two forwards of a step (label and label 0, utils.py:151-152) run as one 2B-row batch, and the part of the
    network that comes before the first label-dependent term (head conv, first ResBlock, the first attention
    block up to and including its 64x64 self-attention) is evaluated once for both (engine.forward shared_prefix);
  * everything that does not depend on x_t is hoisted out of the loop: the time MLP and every
    ResBlock's linear_time for all T steps ([T, sum Cout] table), the label MLP and the ten
    cross-attention vectors (constant over the loop);
  * the per-step scalars (coeff1, coeff2, sigma) and the time rows are indexed on the device by a
    step counter, so one captured graph replays for every step; Philox noise is keyed by (seed, sampler call
    counter, step, GLOBAL sample index), the call counter and the shard offset living in device memory too: every
    call draws a fresh z sequence (the reference's ``randn_like``, utils.py:163) and a batch sharded over ranks
    draws what a single rank would;
  * the per-step host sync of utils.py:167 becomes a device flag checked once after the loop.

The captured graph bakes pointers (packed weights, GroupNorm scratch, conditioning tables) and scalars (w, seed): the
plan keeps strong references to those buffers and drops the graph whenever the engine replaces one of them
(``engine._generation``) or a baked scalar changes.
"""
import torch

from . import ops

F32 = torch.float32
_ALL_ROWS = 1 << 30  # rows_per_sample for a row bias shared by the whole batch


class SamplingPlan:
    def __init__(self, sampler, shape, device):
        self.sampler = sampler
        self.shape = tuple(shape)
        self.device = device
        B, C, H, W = self.shape
        self.x2 = torch.empty(2 * B, C, H, W, device=device, dtype=F32)
        self.eps = torch.empty(2 * B, C, H, W, device=device, dtype=F32)
        self.step = torch.zeros(1, device=device, dtype=torch.int32)
        self.nan_flag = torch.zeros(1, device=device, dtype=torch.int32)
        self.c1, self.c2, self.sigma = sampler._f32_tables(device)
        self.graph = None
        self.graph_sig = None
        self._held = None  # buffers whose addresses the captured graph uses (kept alive with it)
        self.tb_table = None
        self.cond_sig = None
        self.fused_tail = getattr(sampler, "fused_tail", True)
        # the two forwards of a step see the same x_t: their label-independent prefix is computed once
        self.shared_prefix = getattr(sampler, "shared_prefix", True)

    def matches(self, x_T, model):
        return tuple(x_T.shape) == self.shape and x_T.device == self.device and model is self.sampler.model

    # ------------------------------------------------------------------ hoisted conditioning
    def _prepare(self, labels):
        model = self.sampler.model
        eng = model._engine
        P = eng.params()
        dev = self.device
        B = self.shape[0]
        T = self.sampler.T
        blocks = eng._block_list()
        sig = eng._signature(P)
        if self.tb_table is None or self.cond_sig != sig:
            t_all = torch.arange(T, device=dev, dtype=torch.int64)
            temb_all, _, _ = eng.conditioning(P, t_all, torch.zeros(T, device=dev, dtype=torch.int64), save=False)
            rows, self.tb_off = [], {}
            off = 0
            for key, b in blocks:
                if b[0] == "res":
                    rows.append(eng.time_bias(P, key, temb_all))
                    self.tb_off[key] = (off, b[2])
                    off += b[2]
            self.tb_table = torch.cat(rows, dim=1).contiguous()  # [T, sum Cout]
            self.tb_cur = torch.empty(off, device=dev, dtype=F32)
            self.tb_override = {k: (self.tb_cur[o:o + c], _ALL_ROWS) for k, (o, c) in self.tb_off.items()}
            self.cond_sig = sig
            self.graph = None
        labels2 = torch.cat([labels.to(dev).long(), torch.zeros_like(labels, device=dev).long()]).contiguous()
        _, ctx, _ = eng.conditioning(P, torch.zeros(2 * B, device=dev, dtype=torch.int64), labels2, save=False)
        cb = {}
        for key, b in blocks:
            if b[0] == "attn":
                cb[key] = eng.cross_bias(P, key, ctx)[1]
        if getattr(self, "cb_override", None) is None:
            self.cb_override = cb
        else:  # keep the captured graph's pointers: copy in place
            for k in cb:
                self.cb_override[k].copy_(cb[k])

    def _graph_signature(self):
        """Everything a captured step bakes in besides the plan's own buffers."""
        eng = self.sampler.model._engine
        P = eng.params()
        # touch the caches first so that a pending refresh happens here, outside any capture
        W = eng.packed(P, save=False)
        scratch = eng._get_scratch(2 * self.shape[0], self.device)
        sig = (float(self.sampler.w), int(self.sampler.seed), bool(self.fused_tail), bool(self.shared_prefix),
               eng._generation, self.cond_sig)
        return sig, (W, scratch, P)

    def _one_step(self, noise=None):
        model = self.sampler.model
        pos = self.sampler.rng.tensor(self.device)
        ops.gather_row(self.tb_table, self.step, self.tb_cur)
        if self.fused_tail:
            tail = dict(step_ptr=self.step, c1=self.c1, c2=self.c2, sigma=self.sigma, wcfg=float(self.sampler.w),
                        nan_flag=self.nan_flag, noise=noise, seed=self.sampler.seed, clip_last=True, rng=pos)
            model._engine.forward(self.x2, None, None, save=False, tb_override=self.tb_override,
                                  cb_override=self.cb_override, sample_tail=tail, shared_prefix=self.shared_prefix)
        else:
            model._engine.forward(self.x2, None, None, save=False, tb_override=self.tb_override,
                                  cb_override=self.cb_override, eps_out=self.eps, shared_prefix=self.shared_prefix)
            ops.sampler_update(self.x2, self.eps, self.step, self.c1, self.c2, self.sigma, float(self.sampler.w), self.x2,
                               self.nan_flag, noise=noise, seed=self.sampler.seed, clip_last=True, dup=True, rng=pos)
        ops.step_add(self.step, -1)

    def run(self, x_T, labels, steps=None, noise_fn=None):
        B = self.shape[0]
        T = self.sampler.T
        steps = list(reversed(range(T))) if steps is None else list(steps)
        assert all(0 <= s < T for s in steps)
        contiguous_desc = all(steps[i] - 1 == steps[i + 1] for i in range(len(steps) - 1))
        self._prepare(labels)
        self.sampler.rng.advance(self.device)  # a new z sequence for this call (outside the captured step)
        x = x_T.contiguous().float()
        self.x2[:B].copy_(x)
        self.x2[B:].copy_(x)
        self.nan_flag.zero_()
        model = self.sampler.model
        use_graph = (self.sampler.use_cuda_graph and noise_fn is None and contiguous_desc and not model.training
                     and len(steps) > 2)
        if not use_graph:
            for s in steps:
                self.step.fill_(s)
                self._one_step(None if noise_fn is None else noise_fn(s).contiguous().float())
        else:
            self.step.fill_(steps[0])
            remaining = len(steps)
            sig, held = self._graph_signature()
            if self.graph is not None and sig != self.graph_sig:
                self.graph = None  # a baked pointer or scalar changed: capture again
            if self.graph is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._one_step()  # a real step; also warms every lazy initialisation before capture
                torch.cuda.current_stream().wait_stream(side)
                remaining -= 1
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._one_step()
                self.graph = g
                self.graph_sig, self._held = sig, held
            for _ in range(remaining):
                self.graph.replay()
        if int(self.nan_flag.item()) != 0:
            raise AssertionError("nan in tensor.")
        return self.x2[:B].clone()
