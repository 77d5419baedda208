"""Caller-side step body of the reference training loop as two fused passes.

Reference: 02_train_direct.py:72-73 -- ``clip_grad_norm_(params, grad_clip)`` then ``AdamW.step()``
(torch defaults: betas (0.9, 0.999), eps 1e-8; lr and weight_decay as given at :52).  Parameters and
gradients live in flat fp32 buffers, so the global norm is one reduction kernel and clip + AdamW one
sweep; for data-parallel training the flat gradient buffer is what gets all-reduced (NCCL, sum), in
buckets that follow the order in which the backward pass finishes them.

It subclasses ``torch.optim.Optimizer`` so that the reference's ``CosineWarmupScheduler``
(utils.py:75-93) and any other ``LRScheduler`` can drive ``param_groups[0]['lr']`` unchanged, and it
can keep the exponential moving average of the weights that the reference's (unused) ``EMA`` helper
(utils.py:42-72) maintains, updated in the same sweep.

The learning rate and the step count are read by the kernel from device memory, so ``step()`` can be
captured in a CUDA graph (training.GraphedTrainStep) and still follow the schedule on every replay.
"""
import torch
import torch.distributed as dist

from . import ops


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=0.0, ema_decay=None):
        self.model = model
        self.engine = model._engine
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        super().__init__([p for p in model.parameters()], defaults)
        self.ema_decay = ema_decay
        self._lr_on_device = None
        self._flatten()

    # ------------------------------------------------------------------ flat buffers
    def _flatten(self):
        """Re-home every parameter into one flat fp32 buffer (same 16-byte-aligned offsets as the grad buffer)."""
        P = self.engine.params()
        plist = list(P.values())
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedClipAdamW (B200) keeps its state on the GPU: move the model to CUDA first")
        total = sum((p.numel() + 3) // 4 * 4 for p in plist)
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self._offsets = {}
        for k, p in P.items():
            v = flat[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            self._offsets[k] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4
        self.flat_p = flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.ema = flat.clone() if self.ema_decay is not None else None
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)  # number of optimiser steps taken
        self.lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self._lr_on_device = None
        self._host_steps = 0
        self.engine.invalidate()

    def _check_homed(self):
        """The parameters must still be views of flat_p (a later ``model.to()`` / ``.float()`` re-homes ``p.data``)."""
        lo = self.flat_p.data_ptr()
        hi = lo + self.flat_p.numel() * 4
        plist = self.param_groups[0]["params"]
        for p in (plist[0], plist[-1]):  # Module._apply moves every parameter, so the two ends tell
            if not (lo <= p.data_ptr() < hi):
                self._rehome()
                return

    def _rehome(self):
        """Parameters were moved (``Module._apply``): copy their current values into the flat buffer and re-attach."""
        P = self.engine.params()
        for k, p in P.items():
            off, n = self._offsets[k]
            v = self.flat_p[off:off + n].view_as(p)
            v.copy_(p.data.to(self.flat_p.device, torch.float32))
            p.data = v
        self.engine.invalidate()

    @property
    def step_count(self):
        return int(self.step_dev.item())

    def sync_lr(self):
        """Upload ``param_groups[0]['lr']`` when the host value changed (an LRScheduler stepped).  Call it outside a
        graph capture; ``step()`` does it itself in eager mode."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.lr_dev.fill_(lr)
            self._lr_on_device = lr

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    # ------------------------------------------------------------------ data parallel
    def buckets(self):
        """[(begin, end)] element ranges of the flat gradient buffer in the order the backward pass completes them:
        decoders + tail first, then the bottleneck, then the encoders and the conditioning MLPs (whose gradients
        accumulate until the very end).  Keys: 'decoders', 'bottleneck', 'rest'."""
        first = {}
        for k, (off, n) in self._offsets.items():
            sec = k.split(".", 1)[0]
            first.setdefault(sec, off)
        total = self.flat_p.numel()
        b0, d0 = first.get("bottleneck", total), first.get("decoders", total)
        return {"decoders": (d0, total), "bottleneck": (b0, d0), "rest": (0, b0)}

    def all_reduce_grads(self, group=None):
        """Data-parallel exchange: sum the flat gradient buffer over ranks (NCCL over NVLink)."""
        g = self.engine._flat_grad
        if g is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)

    def overlap_all_reduce(self, group=None):
        """Arms bucketed gradient exchange overlapped with the backward pass: as the backward walk leaves a parameter
        section (engine.section_hook), that section's slice of the flat gradient is all-reduced on a side stream while
        the main stream keeps computing.  Returns a ``finish()`` callable that makes the main stream wait for the last
        bucket (call it after ``loss.backward()`` and before ``step()``).  No-op for a single process."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            self.engine.section_hook = None
            return lambda: None
        ranges = self.buckets()
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream()
        pending = []

        def hook(section):
            g = self.engine._flat_grad
            b, e = ranges[section]
            if e <= b:
                return
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dist.all_reduce(g[b:e], op=dist.ReduceOp.SUM, group=group)
            pending.append(section)

        def finish():
            torch.cuda.current_stream().wait_stream(side)
            self.engine.section_hook = None
            pending.clear()

        self.engine.section_hook = hook
        return finish

    def grad_norm(self):
        return float(self.sumsq.sqrt().item())

    # ------------------------------------------------------------------ EMA / checkpoints
    def ema_state_dict(self):
        """The shadow weights (reference EMA.shadow, utils.py:49-58) keyed like model.state_dict()."""
        if self.ema is None:
            raise RuntimeError("FusedClipAdamW was built without ema_decay")
        return {k: self.ema[off:off + n].view_as(p).clone()
                for (k, p), (off, n) in zip(self.engine.params().items(), self._offsets.values())}

    def state_dict(self):
        """Everything a resumed run needs to continue bit-identically (02_train_direct.py:40-50 only saves the model;
        SURVEY 8f-4): Adam moments, step count, EMA shadow, hyper-parameters, keyed per parameter name so the file does
        not depend on the flat layout."""
        P = self.engine.params()
        per = {}
        for k, p in P.items():
            off, n = self._offsets[k]
            per[k] = {"exp_avg": self.m[off:off + n].view_as(p).clone(), "exp_avg_sq": self.v[off:off + n].view_as(p).clone()}
            if self.ema is not None:
                per[k]["ema"] = self.ema[off:off + n].view_as(p).clone()
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return {"state": per, "step": self.step_count, "param_groups": groups, "ema_decay": self.ema_decay}

    def load_state_dict(self, sd):
        P = self.engine.params()
        self._check_homed()
        for k, p in P.items():
            off, n = self._offsets[k]
            st = sd["state"][k]
            self.m[off:off + n].view_as(p).copy_(st["exp_avg"])
            self.v[off:off + n].view_as(p).copy_(st["exp_avg_sq"])
            if self.ema is not None and "ema" in st:
                self.ema[off:off + n].view_as(p).copy_(st["ema"])
        self.step_dev.fill_(int(sd["step"]))
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})
        self._lr_on_device = None

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, closure=None):
        g = self.engine._flat_grad
        if g is None:
            raise RuntimeError("FusedClipAdamW.step() called before any backward pass")
        grp = self.param_groups[0]
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._check_homed()
            self.sync_lr()
        ops.step_add(self.step_dev, 1)
        self.sumsq.zero_()
        ops.sumsq(g, self.sumsq)
        ops.adamw_clip_dev(self.flat_p, g, self.m, self.v, self.lr_dev, self.step_dev, grp["betas"][0], grp["betas"][1],
                           grp["eps"], grp["weight_decay"], grp["max_norm"], self.sumsq)
        if self.ema is not None:
            ops.ema_update(self.ema, self.flat_p, self.ema_decay)
        self.engine.bump()  # parameters changed through raw pointers: packed bf16 copies are stale
