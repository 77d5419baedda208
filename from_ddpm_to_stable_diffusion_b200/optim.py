"""Caller-side step body of the reference training loop as two fused passes.

Reference: 02_train_direct.py:72-73 -- ``clip_grad_norm_(params, grad_clip)`` then ``AdamW.step()``
(torch defaults: betas (0.9, 0.999), eps 1e-8; lr and weight_decay as given at :52).  Parameters and
gradients live in flat fp32 buffers, so the global norm is one reduction kernel and clip + AdamW one
sweep; for data-parallel training the flat gradient buffer is what gets all-reduced (NCCL, sum).

It subclasses ``torch.optim.Optimizer`` so that the reference's ``CosineWarmupScheduler``
(utils.py:75-93) and any other ``LRScheduler`` can drive ``param_groups[0]['lr']`` unchanged, and it
can keep the exponential moving average of the weights that the reference's (unused) ``EMA`` helper
(utils.py:42-72) maintains, updated in the same sweep.
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
        self.step_count = 0
        self.ema_decay = ema_decay
        self._flatten()

    def _flatten(self):
        """Re-home every parameter into one flat fp32 buffer (same 16-byte-aligned offsets as the grad buffer)."""
        P = self.engine.params()
        plist = list(P.values())
        dev = plist[0].device
        total = sum((p.numel() + 3) // 4 * 4 for p in plist)
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for p in plist:
            v = flat[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            off += (p.numel() + 3) // 4 * 4
        self.flat_p = flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.ema = flat.clone() if self.ema_decay is not None else None
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.engine.invalidate()

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def all_reduce_grads(self, group=None):
        """Data-parallel exchange: sum the flat gradient buffer over ranks (NCCL over NVLink)."""
        g = self.engine._flat_grad
        if g is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)

    def grad_norm(self):
        return float(self.sumsq.sqrt().item())

    def ema_state_dict(self):
        """The shadow weights (reference EMA.shadow, utils.py:49-58) keyed like model.state_dict()."""
        if self.ema is None:
            raise RuntimeError("FusedClipAdamW was built without ema_decay")
        out, off = {}, 0
        for k, p in self.engine.params().items():
            out[k] = self.ema[off:off + p.numel()].view_as(p).clone()
            off += (p.numel() + 3) // 4 * 4
        return out

    @torch.no_grad()
    def step(self, closure=None):
        g = self.engine._flat_grad
        if g is None:
            raise RuntimeError("FusedClipAdamW.step() called before any backward pass")
        grp = self.param_groups[0]
        self.step_count += 1
        self.sumsq.zero_()
        ops.sumsq(g, self.sumsq)
        ops.adamw_clip(self.flat_p, g, self.m, self.v, grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"],
                       grp["weight_decay"], self.step_count, grp["max_norm"], self.sumsq)
        if self.ema is not None:
            ops.ema_update(self.ema, self.flat_p, self.ema_decay)
        self.engine.bump()  # parameters changed through raw pointers: packed bf16 copies are stale
