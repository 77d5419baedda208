"""The callers either side of the denoiser: the reference's training-loop glue as B200-native pieces.

Reference: 06_tiny_stable_diffusion/utils.py:10-29 (``means``/``stds``, ``denormalize``, the loader's ToTensor +
Normalize), :42-72 (``EMA``), :75-93 (``CosineWarmupScheduler``); 02_train_direct.py:13-27 (``generate``: sample,
denormalize, save_image grid) and :64-74 (the step body).  Same names, arguments and behaviour; the device work is
done by the library's kernels (csrc/imageio.cu, csrc/optim.cu) and there is no CPU fallback for it.
"""
import math

import numpy as np
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR, LRScheduler

from . import ops

means = [0.485, 0.456, 0.406]
stds = [0.229, 0.224, 0.225]


# ------------------------------------------------------------------ image input / output
def normalize_u8(images_u8_nhwc):
    """What ``animal_faces_loader``'s transform does after the resize (utils.py:21-25), on the device:
    uint8 [N,H,W,3] -> fp32 [N,3,H,W], ``(x / 255 - mean) / std``; bit-exact with ToTensor + Normalize."""
    return ops.u8_to_f32_norm(images_u8_nhwc, means, stds)


def denormalize(tensor):
    """utils.py:14-18 -- ``tensor * std + mean`` (kept as torch glue for API parity; ``image_grid_u8`` fuses it)."""
    device = tensor.device
    mean = torch.tensor(means).view(1, 3, 1, 1).to(device)
    std = torch.tensor(stds).view(1, 3, 1, 1).to(device)
    return tensor * std + mean


def image_grid_u8(img_sample, nrow, padding=0):
    """``save_image(denormalize(img_sample), nrow=nrow, padding=padding)`` up to the PNG encoder
    (02_train_direct.py:24-27): returns the uint8 [GH,GW,3] grid torchvision would hand to PIL."""
    return ops.denorm_grid_u8(img_sample.contiguous().float(), nrow, padding, means, stds)


# ------------------------------------------------------------------ learning-rate schedule (host logic)
class CosineWarmupScheduler(LRScheduler):
    """Same constructor and stepping behaviour as utils.py:75-93, stepped once per epoch (02_train_direct.py:83).

    Two phases share the optimizer: while this scheduler's own epoch counter is below ``warmup_epochs`` the lr climbs
    linearly from the base lr towards ``max_lr``; from then on only the wrapped ``CosineAnnealingLR`` (period
    ``total_epochs - warmup_epochs``) is advanced, and its recursive update continues from whatever lr is in the
    optimizer.  The hand-over step asks the cosine scheduler for its update at its epoch 0, so -- exactly like the
    reference -- the peak is ``2 / (1 + cos(pi / T))`` times the last warm-up value and ``max_lr`` itself is never
    reached (tests/golden/callers.pt pins this behaviour)."""

    def __init__(self, optimizer, warmup_epochs, max_lr, total_epochs, last_epoch=-1):
        self.warmup_epochs, self.max_lr = warmup_epochs, max_lr
        # built first, as in the reference: its constructor records initial_lr and leaves the base lr in place
        self.cosine_scheduler = CosineAnnealingLR(optimizer, T_max=total_epochs - warmup_epochs)
        super().__init__(optimizer, last_epoch)

    def _in_warmup(self):
        return self.last_epoch < self.warmup_epochs

    def get_lr(self):
        if not self._in_warmup():
            return self.cosine_scheduler.get_lr()
        frac_num, frac_den = self.last_epoch, self.warmup_epochs
        return [b + (self.max_lr - b) * frac_num / frac_den for b in self.base_lrs]

    def step(self, epoch=None, metrics=None):
        if self._in_warmup():
            return super().step(epoch)
        self.cosine_scheduler.step(epoch)


# ------------------------------------------------------------------ exponential moving average of the weights
class EMA:
    """utils.py:42-72 with the shadow weights in one flat fp32 buffer: ``update`` is a single kernel sweep when the
    parameters are flat too (after ``FusedClipAdamW``), else one launch per tensor; results are bit-identical to the
    reference's ``(1 - decay) * param + decay * shadow``."""

    def __init__(self, model, decay):
        self.model = model
        self.decay = decay
        self.backup = {}
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named or not named[0][1].is_cuda:
            raise RuntimeError("EMA (B200) keeps its shadow weights on the GPU: move the model to CUDA first")
        self._views = {}
        base = self._common_base(named)
        if base is not None:
            # parameters already live in one flat buffer (FusedClipAdamW): mirror its layout, update = one sweep
            self._flat = base.clone()
            self._layout_of = base.data_ptr()
            for n, p in named:
                off = (p.data.data_ptr() - base.data_ptr()) // 4
                self._views[n] = (self._flat[off:off + p.numel()].view_as(p), off, p.numel())
        else:
            sizes = [(p.numel() + 3) // 4 * 4 for _, p in named]
            self._flat = torch.zeros(sum(sizes), device=named[0][1].device, dtype=torch.float32)
            self._layout_of = None
            off = 0
            for (n, p), sz in zip(named, sizes):
                v = self._flat[off:off + p.numel()].view_as(p)
                v.copy_(p.data)
                self._views[n] = (v, off, sz)
                off += sz

    @staticmethod
    def _common_base(named):
        """The flat fp32 buffer all parameters are carved out of (one storage, contiguous slices), or None."""
        st = named[0][1].data.untyped_storage()
        if st.nbytes() % 16 != 0 or len(named) < 2:
            return None
        for _, p in named:
            d = p.data
            if d.dtype != torch.float32 or not d.is_contiguous() or d.untyped_storage().data_ptr() != st.data_ptr():
                return None
        return torch.empty(0, dtype=torch.float32, device=named[0][1].device).set_(st, 0, (st.nbytes() // 4,))

    @property
    def shadow(self):
        return {n: v for n, (v, _, _) in self._views.items()}

    def _flat_params(self, named):
        """The flat parameter buffer when the shadow mirrors its layout (and the parameters still live there)."""
        if self._layout_of is None:
            return None
        base = self._common_base(named)
        if base is None or base.data_ptr() != self._layout_of or base.numel() != self._flat.numel():
            return None
        return base

    @torch.no_grad()
    def update(self):
        named = [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]
        flat = self._flat_params(named)
        if flat is not None:
            ops.ema_update(self._flat, flat, self.decay)
            return
        for n, p in named:  # parameters not flattened: per-tensor sweeps over 4-element-padded copies
            v, off, sz = self._views[n]
            if p.numel() == sz and p.data.is_contiguous():
                ops.ema_update(self._flat[off:off + sz], p.data.view(-1), self.decay)
            else:
                tmp = torch.zeros(sz, device=p.device, dtype=torch.float32)
                tmp[:p.numel()] = p.data.reshape(-1)
                ops.ema_update(self._flat[off:off + sz], tmp, self.decay)

    def apply_shadow(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.backup[n] = p.data.clone()
                p.data.copy_(self._views[n][0])  # in place: flat-buffer views and packed-weight tracking stay valid
        self._bump()

    def restore(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                p.data.copy_(self.backup[n])
        self._bump()

    def _bump(self):
        eng = getattr(self.model, "_engine", None)
        if eng is not None:
            eng.bump()


# ------------------------------------------------------------------ one optimisation step (02_train_direct.py:64-74)
def train_step(trainer, optimizer, images, labels, train_rand=0.0, grad_clip=None, rng=np.random):
    """zero_grad; labels + 1 (0 is the unconditional class); whole-batch label drop with probability ``train_rand``;
    ``loss = trainer(images, labels).sum() / bs ** 2``; backward; clip_grad_norm_ + optimizer step; returns the loss
    tensor (the reference reads it with ``.item()``).  ``optimizer`` is either ``FusedClipAdamW`` (clip and AdamW in one
    sweep, ``grad_clip`` taken from its ``max_norm``) or any torch optimizer (then ``grad_clip`` is applied here)."""
    optimizer.zero_grad()
    bs = images.shape[0]
    labels = labels + 1
    if rng.rand() < train_rand:
        labels = torch.zeros_like(labels)
    loss = trainer(images, labels).sum() / bs ** 2.
    loss.backward()
    if hasattr(optimizer, "all_reduce_grads"):
        optimizer.all_reduce_grads()
    elif grad_clip is not None:
        torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), grad_clip)
    optimizer.step()
    return loss


@torch.no_grad()
def generate_grid(sampler, num_class, nrow, img_channel, img_size, device, x_T=None):
    """02_train_direct.py:13-27 up to the PNG encoder: ``nrow`` samples of every class, denormalised, as one uint8 grid."""
    values = torch.arange(1, num_class + 1).repeat_interleave(nrow).to(device)
    if x_T is None:
        x_T = torch.randn(size=[num_class * nrow, img_channel, img_size, img_size], device=device)
    img_sample = sampler(x_T, values)
    return image_grid_u8(img_sample, nrow=nrow, padding=0)
