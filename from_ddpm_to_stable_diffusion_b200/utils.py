"""Drop-in for the DDPM part of the reference's ``utils.py``.

Reference: 06_tiny_stable_diffusion/utils.py:32-39 (``extract``), :96-119 (``TrainerDDPM``),
:122-171 (``SamplerDDPM``).  Same constructor arguments, same registered fp64 buffers, same return
values.  The schedule tables are built with the reference's own torch expressions on the host
(fp32 ``linspace`` then ``.double()``), never re-derived on the device, so ``extract`` is bit-exact.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .rng import DeviceRng


def extract(v, t, x_shape):
    """Gather schedule coefficients at timesteps t and reshape to [B, 1, 1, ...] (utils.py:32-39)."""
    device = t.device
    out = torch.gather(v, index=t, dim=0).float().to(device)
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


class _MseFn(torch.autograd.Function):
    """loss = (pred - noise)^2, un-reduced (utils.py:118), with its gradient as one kernel."""

    @staticmethod
    def forward(ctx, pred, noise):
        ctx.save_for_backward(pred, noise)
        return ops.mse_fwd(pred.contiguous(), noise)

    @staticmethod
    def backward(ctx, gout):
        pred, noise = ctx.saved_tensors
        return ops.mse_bwd(pred, noise, gout.expand_as(pred).contiguous().float()), None


class TrainerDDPM(nn.Module):
    def __init__(self, model, beta_1, beta_T, T):
        super().__init__()
        self.model = model
        self.T = T
        self.register_buffer('betas', torch.linspace(beta_1, beta_T, T).double())
        alphas = 1. - self.betas
        alphas_bar = torch.cumprod(alphas, dim=0)
        self.register_buffer('sqrt_alphas_bar', torch.sqrt(alphas_bar))
        self.register_buffer('sqrt_one_minus_alphas_bar', torch.sqrt(1. - alphas_bar))
        self._tables = None
        # timesteps and q_sample noise: Philox keyed by (seed, call counter on the device, GLOBAL sample index); the
        # seed derives from torch.initial_seed(), so torch.manual_seed() selects the stream as in the reference
        self.rng = DeviceRng(salt=0x5EED0001)

    @property
    def seed(self):
        return self.rng.seed

    @seed.setter
    def seed(self, v):
        self.rng.seed = int(v)

    def _f32_tables(self, device):
        # extract() gathers the f64 table then casts to fp32; casting the table first gives the same bits
        if self._tables is None or self._tables[0].device != device:
            self._tables = (self.sqrt_alphas_bar.float().to(device).contiguous(),
                            self.sqrt_one_minus_alphas_bar.float().to(device).contiguous())
        return self._tables

    def forward(self, x_0, labels, t=None, noise=None):
        """Returns the un-reduced noise-MSE [B,C,H,W] (utils.py:111-119).  ``t`` / ``noise`` may be injected
        for parity tests; by default t ~ U{0..T-1} per sample and noise ~ N(0,1) come from the on-device Philox
        stream (the noise fused into the q_sample kernel), a function of (seed, call number, global sample index)."""
        if not x_0.is_cuda:
            raise RuntimeError("TrainerDDPM (B200) runs on CUDA only; there is no CPU fallback")
        dev = x_0.device
        pos = self.rng.advance(dev)
        if t is None:
            t = ops.draw_timesteps(x_0.shape[0], self.T, self.rng.seed, pos, dev)
        sa, sb = self._f32_tables(dev)
        x_0 = x_0.contiguous().float()
        x_t, noise = ops.q_sample(x_0, t.contiguous(), sa, sb, seed=self.rng.seed, rng=pos,
                                  noise=None if noise is None else noise.contiguous().float())
        pred_noise = self.model(x_t, t, labels)
        return _MseFn.apply(pred_noise, noise)


class SamplerDDPM(nn.Module):
    def __init__(self, model, beta_1, beta_T, T, w=0.):
        super().__init__()
        self.model = model
        self.T = T
        self.w = w
        self.register_buffer('betas', torch.linspace(beta_1, beta_T, T).double())
        alphas = 1. - self.betas
        alphas_bar = torch.cumprod(alphas, dim=0)
        alphas_bar_prev = F.pad(alphas_bar, [1, 0], value=1)[:T]
        self.register_buffer('coeff1', torch.sqrt(1. / alphas))
        self.register_buffer('coeff2', self.coeff1 * (1. - alphas) / torch.sqrt(1. - alphas_bar))
        self.register_buffer('posterior_var', self.betas * (1. - alphas_bar_prev) / (1. - alphas_bar))
        # per-step z: Philox keyed by (seed, call counter, time step, GLOBAL sample index): two sampler calls never
        # reuse a noise sequence and a batch sharded over ranks draws what one rank would
        self.rng = DeviceRng(salt=0x5EED0002)
        self.use_cuda_graph = True
        self.fused_tail = True  # final conv + CFG + posterior update + noise as one kernel
        self._plan = None

    @property
    def seed(self):
        return self.rng.seed

    @seed.setter
    def seed(self, v):
        self.rng.seed = int(v)

    # device tables -----------------------------------------------------------------------------
    def _f32_tables(self, device):
        var = torch.cat([self.posterior_var[1:2], self.betas[1:]])  # "fixed-large" variance (utils.py:149)
        c1 = self.coeff1.float().to(device).contiguous()
        c2 = self.coeff2.float().to(device).contiguous()
        sigma = torch.sqrt(var.float()).to(device).contiguous()  # sqrt taken in fp32 after the cast, as utils.py:166
        return c1, c2, sigma

    def forward(self, x_T, labels, steps=None, noise_fn=None):
        """Runs the T-step reverse process (utils.py:157-171) and returns clip(x_0, -1, 1).

        steps: optional iterable of time steps to run (default reversed(range(T))); used by benchmarks
        to time a prefix.  noise_fn(time_step) -> z tensor injects the per-step noise for parity tests."""
        from .sampling import SamplingPlan
        if not x_T.is_cuda:
            raise RuntimeError("SamplerDDPM (B200) runs on CUDA only; there is no CPU fallback")
        with torch.no_grad():
            plan = self._plan
            if plan is None or not plan.matches(x_T, self.model):
                plan = SamplingPlan(self, x_T.shape, x_T.device)
                self._plan = plan
            return plan.run(x_T, labels, steps=steps, noise_fn=noise_fn)
