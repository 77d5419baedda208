"""The training iteration as the caller runs it (02_train_direct.py:64-74) on the CUDA path: parity at the benchmark
batch size, the captured-graph iteration against the eager one, micro-batch accumulation and resume parity."""
import pytest
import torch

from oracle import ref_unet as R

pytestmark = pytest.mark.gpu
MULTY = [1, 2, 2, 2]
BETAS = (0.0015, 0.0195, 1000)


def _model(cuda, dropout=0.0, seed=0):
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    sd = R.init_state_dict(seed, 3, MULTY, 128, 3)
    m = Diffusion(3, MULTY, 128, num_class=3, dropout=dropout)
    m.load_state_dict(sd)
    return m.to(cuda), sd


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


class _NoDrop:
    @staticmethod
    def rand():
        return 1.0


def test_benchmark_batch_256_matches_oracle(cuda):
    """BASELINE configs[1] size (batch 256, 3x64x64: 1 M-row GEMMs, 2 GB tensors, 16 key blocks reducing into dQ):
    eps of samples {0, 127, 255} against the oracle run on those images, and loss + a set of parameter gradients of
    the whole batch against oracle autograd accumulated over chunks of 8 samples."""
    from from_ddpm_to_stable_diffusion_b200 import TrainerDDPM
    m, sd = _model(cuda)
    m.train()  # dropout p = 0: deterministic
    B = 256
    g = torch.Generator().manual_seed(2026)
    x0 = torch.randn(B, 3, 64, 64, generator=g)
    y = torch.randint(0, 4, (B,), generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn(B, 3, 64, 64, generator=g)
    sched = R.make_schedule(*BETAS)
    # ---- forward: three samples of the 256-row batch vs the oracle on those three images alone
    x_t = R.q_sample(sched, x0, t, noise)
    with torch.no_grad():
        eps = m(x_t.to(cuda), t.to(cuda), y.to(cuda)).cpu()
        pick = torch.tensor([0, 127, 255])
        ref = R.unet_forward(sd, x_t[pick], t[pick], y[pick], MULTY)
    for j, i in enumerate(pick.tolist()):
        e = _rel(eps[i], ref[j])
        assert e < 2e-2, f"sample {i}: eps rel-L2 {e}"
    # ---- training step: loss and gradients of the whole batch
    trainer = TrainerDDPM(m, *BETAS).to(cuda)
    loss = trainer(x0.to(cuda), y.to(cuda), t=t.to(cuda), noise=noise.to(cuda)).sum() / B ** 2
    loss.backward()
    torch.cuda.synchronize()
    watch = ["encoders.1.0.conv_1.2.weight", "encoders.1.1.atten_1.1.in_proj.weight", "encoders.1.1.linear_1.weight",
             "encoders.3.1.atten_1.1.in_proj.weight", "bottleneck.1.atten_1.1.out_proj.weight",
             "decoders.7.1.atten_1.1.in_proj.weight", "decoders.7.0.conv_2.3.weight", "decoders.6.1.linear_2.weight",
             "decoders.5.2.conv.weight", "encoders.2.0.weight", "encoders.0.0.weight", "tail.2.weight", "tail.0.weight",
             "time_embedding.mlp.0.weight", "label_embedding.0.weight", "decoders.7.0.linear_time.1.weight",
             "decoders.7.1.atten_2.v_proj.weight", "encoders.1.1.norm_3.weight", "decoders.4.0.residual_layer.weight"]
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref = 0.0
    for c in range(0, B, 8):
        sl = slice(c, c + 8)
        lr_ = R.trainer_loss(sdg, sched, x0[sl], y[sl], t[sl], noise[sl], MULTY, use_sdpa=True).sum() / B ** 2
        lr_.backward()
        loss_ref += lr_.item()
    assert abs(loss.item() - loss_ref) / abs(loss_ref) < 1e-3, (loss.item(), loss_ref)
    P = dict(m.named_parameters())
    gn_ref = sum(float(v.grad.double().pow(2).sum()) for v in sdg.values() if v.grad is not None) ** 0.5
    gn_got = sum(float(p.grad.double().pow(2).sum()) for p in P.values()) ** 0.5
    assert abs(gn_got - gn_ref) / gn_ref < 1e-3, (gn_got, gn_ref)
    for k in watch:
        gr, gg = sdg[k].grad, P[k].grad.float().cpu()
        cos = (gr.flatten() @ gg.flatten() / (gr.norm() * gg.norm() + 1e-30)).item()
        assert cos > 0.995, (k, cos)


def _setup(cuda, dropout, ema=None):
    from from_ddpm_to_stable_diffusion_b200 import TrainerDDPM
    from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW
    torch.manual_seed(1234)  # the device Philox seeds derive from torch.initial_seed()
    m, _ = _model(cuda, dropout=dropout)
    m.train()
    tr = TrainerDDPM(m, *BETAS).to(cuda)
    opt = FusedClipAdamW(m, lr=1e-4, weight_decay=1e-5, max_norm=1.0, ema_decay=ema)
    return m, tr, opt


def _batches(n, B, S=32):
    g = torch.Generator().manual_seed(77)
    return [(torch.randn(B, 3, S, S, generator=g), torch.randint(0, 3, (B,), generator=g)) for _ in range(n)]


def test_graphed_iteration_matches_eager(cuda):
    """GraphedTrainStep (whole iteration captured) == training.train_step, step after step: same device
    random streams (timesteps, noise, dropout), same LR schedule through the device scalar, same AdamW step count."""
    from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep, train_step
    data = _batches(4, 8)
    lrs = [1e-4, 1e-4, 5e-5, 2e-5]
    m1, tr1, opt1 = _setup(cuda, 0.1, ema=0.99)
    eager = []
    for (x, y), lr in zip(data, lrs):
        opt1.param_groups[0]["lr"] = lr
        eager.append(train_step(tr1, opt1, x.to(cuda), y.to(cuda), train_rand=0.0, rng=_NoDrop).item())
    m2, tr2, opt2 = _setup(cuda, 0.1, ema=0.99)
    stepper = GraphedTrainStep(tr2, opt2, train_rand=0.0, rng=_NoDrop)
    graphed = []
    for (x, y), lr in zip(data, lrs):
        opt2.param_groups[0]["lr"] = lr
        graphed.append(stepper(x.pin_memory(), y.pin_memory()).item())
    assert opt1.step_count == opt2.step_count == 4
    assert tr1.rng.calls == tr2.rng.calls and m1._engine.rng.calls == m2._engine.rng.calls
    assert len(set(eager)) == 4  # every step drew new timesteps / noise
    # dQ's arrival-order sums make two runs of the same step differ in the last bits; at lr 1e-4 the loss moves by 15 %
    # per step, which amplifies that to ~1e-4 relative after one step (stale weights would show as ~1e-1)
    for a, b in zip(eager, graphed):
        assert abs(a - b) / abs(a) < 1e-3, (eager, graphed)
    # dQ partials are summed by the L2 in arrival order, so the two runs agree to fp32 re-association, not bit for bit;
    # Adam turns noise-level gradients into +-lr steps, so a few weights differ by O(lr) after four steps
    assert _rel(opt2.flat_p, opt1.flat_p) < 2e-3
    assert _rel(opt2.m, opt1.m) < 2e-2 and _rel(opt2.ema, opt1.ema) < 2e-4
    # an eager consumer after the replays sees the updated weights (its packed copies are refreshed): the same forward
    # through a fresh module that loaded m2's current state_dict gives the same bits
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    m2.eval()
    m3 = Diffusion(3, MULTY, 128, num_class=3, dropout=0.1)
    m3.load_state_dict({k: v.detach().cpu() for k, v in m2.state_dict().items()})
    m3 = m3.to(cuda).eval()
    x = data[0][0][:2].to(cuda)
    tt = torch.tensor([10, 900], device=cuda)
    yy = torch.tensor([1, 2], device=cuda)
    with torch.no_grad():
        assert torch.equal(m2(x, tt, yy), m3(x, tt, yy))


def test_micro_batch_accumulation(cuda):
    """Global batch in k micro-batches (BASELINE configs[3] on fewer GPUs) == the same batch in one pass: the random
    streams are keyed by global sample index, the loss is normalised by the full batch."""
    from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep
    (x, y), = _batches(1, 8)
    outs = []
    for k in (1, 4):
        m, tr, opt = _setup(cuda, 0.1)
        loss = GraphedTrainStep(tr, opt, micro_batches=k, rng=_NoDrop)(x.to(cuda), y.to(cuda)).item()
        outs.append((loss, m._engine._flat_grad.clone(), opt.flat_p.clone()))
    assert abs(outs[0][0] - outs[1][0]) / abs(outs[0][0]) < 1e-5, (outs[0][0], outs[1][0])
    assert _rel(outs[1][1], outs[0][1]) < 1e-2  # clipped gradients (written back by the optimiser sweep); bf16-level
    assert _rel(outs[1][2], outs[0][2]) < 2e-3  # one AdamW step of +-lr per weight on re-associated gradients


def test_label_drop_flag(cuda):
    """Whole-batch label drop (02_train_direct.py:68-69) inside the captured iteration follows the host decision."""
    from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep

    class Seq:
        def __init__(self, vals):
            self.vals = list(vals)

        def rand(self):
            return self.vals.pop(0)

    (x, y), = _batches(1, 4)
    losses = {}
    for name, r in (("keep", 0.9), ("drop", 0.0)):
        m, tr, opt = _setup(cuda, 0.0)
        losses[name] = GraphedTrainStep(tr, opt, train_rand=0.5, rng=Seq([r]))(x.to(cuda), y.to(cuda)).item()
    m, tr, opt = _setup(cuda, 0.0)
    from from_ddpm_to_stable_diffusion_b200.training import train_step
    ref_drop = train_step(tr, opt, x.to(cuda), y.to(cuda), train_rand=0.5, rng=Seq([0.0])).item()
    assert abs(losses["drop"] - ref_drop) / abs(ref_drop) < 1e-4
    assert abs(losses["drop"] - losses["keep"]) / abs(ref_drop) > 1e-4  # the label does change the prediction


def test_resume_is_bit_compatible(cuda, tmp_path):
    """Train 3 steps, save model + optimiser (moments, step, EMA) + random-stream positions, reload into fresh objects:
    step 4 matches the uninterrupted run (02_train_direct.py:40-50,85-88)."""
    from from_ddpm_to_stable_diffusion_b200.training import load_training_state, train_step, training_state
    data = _batches(4, 4)
    m1, tr1, opt1 = _setup(cuda, 0.1, ema=0.9)
    for x, y in data[:3]:
        train_step(tr1, opt1, x.to(cuda), y.to(cuda), rng=_NoDrop)
    path = tmp_path / "ckpt_003.pth"
    torch.save(training_state(tr1, opt1), path)
    loss_a = train_step(tr1, opt1, data[3][0].to(cuda), data[3][1].to(cuda), rng=_NoDrop).item()
    # a new process would rebuild everything from the constructor arguments and the file
    torch.manual_seed(999)  # different default seeds: the checkpoint must carry the stream identity
    from from_ddpm_to_stable_diffusion_b200 import Diffusion, TrainerDDPM
    from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW
    m2 = Diffusion(3, MULTY, 128, num_class=3, dropout=0.1).to(cuda).train()
    tr2 = TrainerDDPM(m2, *BETAS).to(cuda)
    opt2 = FusedClipAdamW(m2, lr=1e-4, weight_decay=1e-5, max_norm=1.0, ema_decay=0.9)
    load_training_state(tr2, opt2, torch.load(path, weights_only=False))
    assert opt2.step_count == 3
    loss_b = train_step(tr2, opt2, data[3][0].to(cuda), data[3][1].to(cuda), rng=_NoDrop).item()
    assert abs(loss_a - loss_b) / abs(loss_a) < 1e-6, (loss_a, loss_b)
    # (dQ is summed in arrival order: the resumed step equals the uninterrupted one up to fp32 re-association)
    assert _rel(opt2.flat_p, opt1.flat_p) < 1e-3 and _rel(opt2.m, opt1.m) < 2e-2 and _rel(opt2.ema, opt1.ema) < 3e-4
    # the reference's own checkpoint content is the 425-key model state_dict
    assert len(torch.load(path, weights_only=False)["model"]) == 425


def test_trainer_draws_fresh_numbers_per_call_and_per_global_sample(cuda):
    """q_sample noise / timesteps: new on every call, and a function of the GLOBAL sample index (a rank holding
    samples [4, 8) of a batch draws what a single process draws for those samples)."""
    from from_ddpm_to_stable_diffusion_b200 import ops
    from from_ddpm_to_stable_diffusion_b200.rng import DeviceRng
    r = DeviceRng(salt=1, seed=42)
    pos = r.advance(cuda)
    t_full = ops.draw_timesteps(8, 1000, r.seed, pos, cuda)
    x0 = torch.zeros(8, 3, 16, 16, device=cuda)
    ones = torch.ones(1000, device=cuda)
    _, n_full = ops.q_sample(x0, t_full, ones, ones, seed=r.seed, rng=pos)
    r.set_sample0(4)
    t_half = ops.draw_timesteps(4, 1000, r.seed, pos, cuda)
    _, n_half = ops.q_sample(x0[:4], t_half, ones, ones, seed=r.seed, rng=pos)
    assert torch.equal(t_half, t_full[4:]) and torch.equal(n_half, n_full[4:])
    r.set_sample0(0)
    r.advance(cuda)
    t2 = ops.draw_timesteps(8, 1000, r.seed, pos, cuda)
    _, n2 = ops.q_sample(x0, t2, ones, ones, seed=r.seed, rng=pos)
    assert not torch.equal(t2, t_full) and not torch.equal(n2, n_full)
    assert 0 <= int(t_full.min()) and int(t_full.max()) < 1000
    big = ops.draw_timesteps(200000, 1000, r.seed, pos, cuda).float()
    assert abs(big.mean().item() - 499.5) < 3.0 and abs(n_full.std().item() - 1.0) < 0.05


def test_graph_is_recaptured_when_the_engine_replaces_its_buffers(cuda):
    """load_state_dict() between two graphed steps drops the engine's packed-weight buffers: the stepper must notice and
    capture again instead of replaying launches that point at freed memory; the step then equals an eager step from the
    same state."""
    from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep, train_step
    data = _batches(3, 4)
    m1, tr1, opt1 = _setup(cuda, 0.0)
    step = GraphedTrainStep(tr1, opt1, rng=_NoDrop)
    step(data[0][0].to(cuda), data[0][1].to(cuda))
    sd = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    held = step._held
    m1.load_state_dict(sd)  # same values, but the engine invalidates every derived buffer
    junk = [torch.randn(1 << 22, device=cuda) for _ in range(8)]  # recycle whatever was released
    la = step(data[1][0].to(cuda), data[1][1].to(cuda)).item()
    assert step._held[0] is not held[0]  # captured again on the new buffers
    del junk
    m2, tr2, opt2 = _setup(cuda, 0.0)
    train_step(tr2, opt2, data[0][0].to(cuda), data[0][1].to(cuda), rng=_NoDrop)
    lb = train_step(tr2, opt2, data[1][0].to(cuda), data[1][1].to(cuda), rng=_NoDrop).item()
    assert abs(la - lb) / abs(lb) < 1e-3, (la, lb)
