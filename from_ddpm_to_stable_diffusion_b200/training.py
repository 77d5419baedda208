"""The callers either side of the denoiser: the reference's training-loop glue as B200-native pieces.

Reference: 06_tiny_stable_diffusion/utils.py:10-29 (``means``/``stds``, ``denormalize``, the loader's ToTensor +
Normalize), :42-72 (``EMA``), :75-93 (``CosineWarmupScheduler``); 02_train_direct.py:13-27 (``generate``: sample,
denormalize, save_image grid) and :64-74 (the step body).  Same names, arguments and behaviour; the device work is
done by the library's kernels (csrc/imageio.cu, csrc/optim.cu) and there is no CPU fallback for it.
"""

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import CosineAnnealingLR, LRScheduler

from . import ops
from .parallel import set_shard, world_info

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
    # data parallel: every rank holds bs images of a world * bs batch; the reference's normalisation sum / B^2 uses the
    # GLOBAL batch so that the summed (all-reduced) gradient equals the single-process one (parallel.dp_loss_scale)
    rank, world = world_info()
    bs = images.shape[0] * world
    set_shard(trainer, rank * images.shape[0])
    labels = labels + 1
    if rng.rand() < train_rand:
        labels = torch.zeros_like(labels)
    finish = optimizer.overlap_all_reduce() if hasattr(optimizer, "overlap_all_reduce") else None
    loss = trainer(images, labels).sum() / bs ** 2.
    loss.backward()
    if finish is not None:
        finish()
    elif grad_clip is not None:
        torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), grad_clip)
    optimizer.step()
    return loss


def _launch_count():
    import ctypes

    from . import _lib
    fn = _lib.lib().tsd_launch_count
    fn.restype = ctypes.c_ulonglong
    return int(fn())


class GraphedTrainStep:
    """The whole training iteration of 02_train_direct.py:64-74 as replayed CUDA graphs.

    ``step(images, labels)`` does what ``train_step`` does -- label shift / whole-batch label drop, q_sample,
    UNet forward, noise-MSE normalised by the global batch squared, backward, gradient all-reduce, clip + AdamW
    (+ EMA) -- but the ~1 100 kernel launches of one iteration are captured once and replayed, so the host issues a
    handful of calls per step.  What makes the capture replayable: timesteps, q_sample noise and dropout masks are
    drawn on the device from Philox streams whose call counters live in device memory (rng.DeviceRng), the learning
    rate and the AdamW step count are read from device memory (optim.FusedClipAdamW), inputs are copied into static
    buffers, and the label-drop decision (host RNG, as in the reference) is a device flag.

    ``micro_batches`` > 1 accumulates gradients over equal slices of the local batch (BASELINE configs[3]: global
    batch 2048 on 2 / 4 GPUs); the exchange happens once, after the last slice.  With ``overlap=True`` the all-reduce
    runs in buckets on a side stream inside the captured backward (NCCL captures into the graph); otherwise it is
    issued eagerly between the backward graph and the optimiser graph.
    """

    def __init__(self, trainer, optimizer, train_rand=0.0, micro_batches=1, overlap=True, group=None, rng=np.random):
        if not hasattr(optimizer, "sync_lr"):
            raise RuntimeError("GraphedTrainStep needs optim.FusedClipAdamW (learning rate and step count on the device)")
        self.trainer, self.opt = trainer, optimizer
        self.train_rand, self.micro, self.overlap, self.group, self.host_rng = train_rand, int(micro_batches), overlap, group, rng
        self.rank, self.world = world_info()
        self._graphs = None
        self._shape = None

    # -------------------------------------------------------------- pieces of one iteration
    def _fwd_bwd(self):
        if torch.cuda.is_current_stream_capturing():
            # the optimiser inside the graph changes the parameters on every replay: the refresh of the packed bf16
            # weights must be part of the captured work, whatever the host-side change tracking believes right now
            self.trainer.model._engine.bump()
        labels = (self._y + 1) * self._keep
        loss = self.trainer(self._x, labels).sum() * self._scale
        loss.backward()
        self.loss.add_(loss.detach())

    def _fwd_bwd_exchange(self):
        finish = self.opt.overlap_all_reduce(self.group)
        self._fwd_bwd()
        finish()

    def _build(self, images):
        dev = self.trainer.sqrt_alphas_bar.device
        B = images.shape[0]
        assert B % self.micro == 0, "local batch must be divisible by micro_batches"
        mb = B // self.micro
        self._shape = tuple(images.shape)
        self._mb = mb
        self._x = torch.empty((mb,) + tuple(images.shape[1:]), device=dev, dtype=torch.float32)
        self._y = torch.zeros(mb, device=dev, dtype=torch.int64)
        self._keep = torch.ones(1, device=dev, dtype=torch.int64)
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)
        self._scale = 1.0 / float(B * self.world) ** 2
        eng = self.trainer.model._engine
        # ---- warm-up: one eager forward + backward on a side stream (lazy initialisation, flat gradient buffer, arena
        # sizing, NCCL communicator), with the random streams put back so that the first replay draws what an eager
        # first step would
        saved = (self.trainer.rng.state_dict(), eng.rng.state_dict())
        # gradient accumulation: the micro-batches of one step share one call number (their samples are told apart by
        # the global sample index), so the counters are advanced once per step by __call__, not by every replay
        self._own_counters = self.micro > 1
        self.trainer.rng.frozen = eng.rng.frozen = self._own_counters
        self._x.normal_()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.opt.zero_grad()
            for _ in range(2):  # the second pass runs with the arena at its final size
                if self.world > 1 and self.overlap:
                    self._fwd_bwd_exchange()  # also creates the side stream and the NCCL communicator before capture
                else:
                    self._fwd_bwd()
            if self.world > 1:
                dist.all_reduce(self.loss.clone(), group=self.group)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.trainer.rng.load_state_dict(saved[0])
        eng.rng.load_state_dict(saved[1])
        torch.cuda.empty_cache()
        # ---- capture.  Gradients stay attached to the flat buffer (no zero_grad inside: zeroing is one eager memset)
        mode = "thread_local" if self.world > 1 else "global"
        pool = None
        graphs = {}

        self.launches = {}  # kernels of this library per replay of each graph (for bench.py's gpu_launches)

        def capture(name, fn):
            nonlocal pool
            g = torch.cuda.CUDAGraph()
            n0 = _launch_count()
            with torch.cuda.graph(g, pool=pool, capture_error_mode=mode):
                fn()
            self.launches[name] = _launch_count() - n0
            pool = g.pool()
            graphs[name] = g

        exchange_in_graph = self.world > 1 and self.overlap
        if self.world == 1 and self.micro == 1:
            capture("full", lambda: (self._fwd_bwd(), self.opt.step()))
        else:
            if self.micro > 1 or not exchange_in_graph:
                capture("fb", self._fwd_bwd)
            if exchange_in_graph:
                capture("fb_last", self._fwd_bwd_exchange)
            capture("opt", self.opt.step)
        self._graphs = graphs
        self._exchange_in_graph = exchange_in_graph
        self._held = self._engine_buffers()

    def _engine_buffers(self):
        """The engine-owned buffers whose addresses the captured graphs use.  Holding them keeps them alive; comparing
        identities tells when the engine replaced one (load_state_dict, .to(), a re-homed parameter buffer) and the
        graphs must be captured again."""
        eng = self.trainer.model._engine
        return (eng._packs.get("train"), eng._packs.get("dgrad"), eng._flat_grad, eng._arena.buf, eng._scratch,
                self.opt.flat_p)

    def _stale(self):
        return any(a is not b for a, b in zip(self._held, self._engine_buffers()))

    def launches_per_step(self):
        """Library kernels executed by one call (replayed graph nodes; torch glue and NCCL not counted)."""
        L = self.launches
        if "full" in L:
            return L["full"]
        fb_last = L["fb_last"] if self._exchange_in_graph else L["fb"]
        return (self.micro - 1) * L.get("fb", 0) + fb_last + L["opt"]

    # -------------------------------------------------------------- the call
    def __call__(self, images, labels):
        if self._graphs is None or tuple(images.shape) != self._shape or self._stale():
            self._graphs = None  # drop the old graphs (and their memory pool) before capturing again
            self._build(images)
        eng = self.trainer.model._engine
        g = self._graphs
        keep = 0 if self.host_rng.rand() < self.train_rand else 1
        self._keep.fill_(keep)
        self.opt.sync_lr()
        eng._flat_grad.zero_()
        self.loss.zero_()
        mb, B = self._mb, self._shape[0]
        if self._own_counters:
            dev = self._x.device
            self.trainer.rng.advance(dev, force=True)
            eng.rng.advance(dev, force=True)
        for i in range(self.micro):
            self._x.copy_(images[i * mb:(i + 1) * mb], non_blocking=True)
            self._y.copy_(labels[i * mb:(i + 1) * mb], non_blocking=True)
            set_shard(self.trainer, self.rank * B + i * mb)
            last = i == self.micro - 1
            if "full" in g:
                g["full"].replay()
            elif last and self._exchange_in_graph:
                g["fb_last"].replay()
            else:
                g["fb"].replay()
        if "full" not in g:
            if not self._exchange_in_graph:
                self.opt.all_reduce_grads(self.group)
            g["opt"].replay()
        eng.bump()  # the replay changed the parameters behind the host's back: packed copies held by eager paths are stale
        return self.loss


@torch.no_grad()
def generate_grid(sampler, num_class, nrow, img_channel, img_size, device, x_T=None):
    """02_train_direct.py:13-27 up to the PNG encoder: ``nrow`` samples of every class, denormalised, as one uint8 grid."""
    values = torch.arange(1, num_class + 1).repeat_interleave(nrow).to(device)
    if x_T is None:
        x_T = torch.randn(size=[num_class * nrow, img_channel, img_size, img_size], device=device)
    img_sample = sampler(x_T, values)
    return image_grid_u8(img_sample, nrow=nrow, padding=0)


# ------------------------------------------------------------------ checkpoint / resume (02_train_direct.py:40-50,85-88)
def training_state(trainer, optimizer):
    """Everything needed to resume a run bit-identically.  The reference saves only ``diffusion.state_dict()``
    (02_train_direct.py:85-88) and restarts Adam and the random streams on resume; this adds the optimiser moments,
    the step count, the EMA shadow and the positions of the device Philox streams (timesteps, q_sample noise, dropout).
    ``state['model']`` alone is the reference's ``ckpt_XXX.pth`` content (same 425 keys)."""
    eng = trainer.model._engine
    return {"model": {k: v.detach().clone() for k, v in trainer.model.state_dict().items()},
            "optimizer": optimizer.state_dict(),
            "rng": {"trainer": trainer.rng.state_dict(), "dropout": eng.rng.state_dict()}}


def load_training_state(trainer, optimizer, state):
    trainer.model.load_state_dict(state["model"], strict=False)
    if hasattr(optimizer, "_check_homed"):
        optimizer._check_homed()
    optimizer.load_state_dict(state["optimizer"])
    trainer.rng.load_state_dict(state["rng"]["trainer"])
    trainer.model._engine.rng.load_state_dict(state["rng"]["dropout"])
    trainer.model._engine.bump()
