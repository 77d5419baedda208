"""UNet forward / backward and DDPM steps on the B200 vs the CPU oracle (oracle/ref_unet.py)."""
import pytest
import torch

from oracle import ref_unet as R

pytestmark = pytest.mark.gpu

MULTY = [1, 2, 2, 2]


def _model(cuda, channel_img=3, num_class=3, dropout=0.0, seed=0):
    from from_ddpm_to_stable_diffusion_b200 import Diffusion
    sd = R.init_state_dict(seed, channel_img, MULTY, 128, num_class)
    m = Diffusion(channel_img, MULTY, 128, num_class=num_class, dropout=dropout)
    m.load_state_dict(sd)
    return m.to(cuda).eval(), sd


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _inputs(B, C, H, seed=1234, num_class=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, H, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, num_class + 1, (B,), generator=g)
    return x, t, y


def test_forward_blocks_64(cuda):
    """Every block output against the fp32 oracle (bf16 tolerance: rel-L2 <= 2e-2 per block, eps <= 2e-2), and the
    SURVEY 8c gate: eps error vs the fp64 oracle <= 1.5x the error of torch's own bf16 autocast on the same inputs."""
    m, sd = _model(cuda)
    x, t, y = _inputs(2, 3, 64)
    taps_ref = {}
    with torch.no_grad():
        ref = R.unet_forward(sd, x, t, y, MULTY, taps=taps_ref)
        ref64 = R.unet_forward({k: v.double() for k, v in sd.items()}, x.double(), t, y, MULTY)
        sdc = {k: v.to(cuda) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            tbf = R.unet_forward(sdc, x.to(cuda), t.to(cuda), y.to(cuda), MULTY).float().cpu()
        taps = {}
        eps, _ = m._engine.forward(x.to(cuda), t.to(cuda), y.to(cuda), save=False, taps=taps)
    torch.cuda.synchronize()
    worst = 0.0
    for k, (v, h, w) in taps.items():
        got = v.float().view(2, h, w, -1).permute(0, 3, 1, 2)
        e = _rel_l2(got, taps_ref[k])
        worst = max(worst, e)
        assert e < 2e-2, f"block {k}: rel-L2 {e}"
    e = _rel_l2(eps, ref)
    assert e < 2e-2, f"eps rel-L2 {e} (worst block {worst})"
    e64, ebf = _rel_l2(eps, ref64), _rel_l2(tbf, ref64)
    print(f"eps rel-L2 vs fp64: ours {e64:.3e}, torch bf16 autocast {ebf:.3e}, worst block {worst:.3e}")
    assert e64 <= 1.5 * ebf, f"eps error {e64} exceeds 1.5x torch-bf16's own error {ebf}"


def test_forward_latent_16(cuda):
    m, sd = _model(cuda, channel_img=4, num_class=10)
    x, t, y = _inputs(16, 4, 16, num_class=10)
    with torch.no_grad():
        ref = R.unet_forward(sd, x, t, y, MULTY)
        eps = m(x.to(cuda), t.to(cuda), y.to(cuda))
    assert _rel_l2(eps, ref) < 2e-2


def test_backward_32(cuda):
    """Trainer loss + parameter gradients vs oracle autograd (fp32 CPU), gates from SURVEY 8c."""
    from from_ddpm_to_stable_diffusion_b200 import TrainerDDPM
    m, sd = _model(cuda)
    m.train()  # dropout p = 0 -> deterministic
    B = 4
    x0, t, y = _inputs(B, 3, 32, seed=77)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(B, 3, 32, 32, generator=g)
    sched = R.make_schedule(0.0015, 0.0195, 1000)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref = R.trainer_loss(sdg, sched, x0, y, t, noise, MULTY).sum() / B ** 2
    loss_ref.backward()
    trainer = TrainerDDPM(m, 0.0015, 0.0195, 1000).to(cuda)
    loss = trainer(x0.to(cuda), y.to(cuda), t=t.to(cuda), noise=noise.to(cuda)).sum() / B ** 2
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()) < 1e-3, (loss.item(), loss_ref.item())
    dead = (".norm_2.", ".atten_2.q_proj.", ".atten_2.k_proj.")
    cos = {}
    gn_ref = gn_got = 0.0
    for k, p in m.named_parameters():
        gr = sdg[k].grad
        gg = p.grad.float().cpu()
        if any(d in k for d in dead):
            assert gg.abs().max().item() <= 1e-6, k
            continue
        gn_ref += gr.double().pow(2).sum().item()
        gn_got += gg.double().pow(2).sum().item()
        if k == "label_embedding.0.weight":
            assert gg[0].abs().max().item() == 0.0
        denom = gr.norm() * gg.norm()
        cos[k] = (gr.flatten() @ gg.flatten() / denom).item() if denom > 0 else 1.0
    vals = sorted(cos.values())
    print(f"grad cosine: min {vals[0]:.5f} median {vals[len(vals) // 2]:.6f}; grad-norm rel err "
          f"{abs(gn_got ** 0.5 - gn_ref ** 0.5) / gn_ref ** 0.5:.2e}")
    # SURVEY 8c gates: per-tensor cosine >= 0.995, median >= 0.9995, loss and global grad norm within 0.1 %
    bad = {k: v for k, v in cos.items() if v < 0.995}
    assert not bad, f"low-cosine grads: {sorted(bad.items(), key=lambda kv: kv[1])[:10]}"
    assert vals[len(vals) // 2] > 0.9995, f"median cosine {vals[len(vals) // 2]}"
    assert abs(gn_got ** 0.5 - gn_ref ** 0.5) / gn_ref ** 0.5 < 1e-3, (gn_got ** 0.5, gn_ref ** 0.5)


def test_sampler_step_teacher_forced(cuda):
    """x_{t-1} from the same x_t and the same z (utils.py:159-166) at several t, CFG w = 1.8."""
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM
    m, sd = _model(cuda)
    w = 1.8
    sampler = SamplerDDPM(m, 0.0015, 0.0195, 1000, w=w).to(cuda)
    sched = R.make_schedule(0.0015, 0.0195, 1000)
    B = 2
    g = torch.Generator().manual_seed(9)
    y = torch.tensor([1, 3])
    for ts in (999, 500, 1, 0):
        x_t = torch.randn(B, 3, 64, 64, generator=g) * (3.0 if ts < 500 else 1.0)
        z = torch.randn(B, 3, 64, 64, generator=g)
        with torch.no_grad():
            ref, ec, eu = R.sampler_step(sd, sched, x_t, y, ts, z, w, MULTY)
        if ts == 0:
            ref = ref.clip(-1, 1)
        got = sampler(x_t.to(cuda), y.to(cuda), steps=[ts], noise_fn=lambda s: z.to(cuda))
        c2 = float(sched["coeff2"][ts])
        tol = c2 * (1 + 2 * w) * 2.5e-2 * max(1.0, float(ec.abs().max())) + 1e-5 * float(x_t.abs().max())
        err = (got.cpu() - ref).abs().max().item()
        assert err <= tol, f"t={ts}: max abs err {err} > tol {tol}"


def test_sampler_graph_matches_eager(cuda):
    """The replayed CUDA graph (device step counter, Philox keyed by step) == eager stepping."""
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM
    m, _ = _model(cuda)
    B = 2
    g = torch.Generator().manual_seed(11)
    xT = torch.randn(B, 3, 64, 64, generator=g).to(cuda)
    y = torch.tensor([2, 1]).to(cuda)
    steps = list(range(999, 993, -1))
    s1 = SamplerDDPM(m, 0.0015, 0.0195, 1000, w=1.8).to(cuda)
    a = s1(xT, y, steps=steps)
    s2 = SamplerDDPM(m, 0.0015, 0.0195, 1000, w=1.8).to(cuda)
    s2.use_cuda_graph = False
    b = s2(xT, y, steps=steps)
    # same kernels on the same data (the forward path has no atomics): replay and eager stepping agree bit for bit
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), (a - b).abs().max().item()


def test_fused_clip_adamw_matches_torch(cuda):
    """tsd_sumsq_f32 + tsd_adamw_clip == clip_grad_norm_(1.0) + torch.optim.AdamW.step() (02_train_direct.py:72-73)."""
    from from_ddpm_to_stable_diffusion_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 1000003
    p0 = torch.randn(n + 1, device=cuda, generator=g)[:n].clone()
    flat = torch.zeros((n + 3) // 4 * 4, device=cuda)
    flat[:n] = p0
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    m = torch.zeros_like(flat)
    v = torch.zeros_like(flat)
    ss = torch.zeros(1, device=cuda)
    for step in range(1, 4):
        grad = torch.randn(n, device=cuda, generator=g) * 3.0
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        gflat = torch.zeros_like(flat)
        gflat[:n] = grad
        ss.zero_()
        ops.sumsq(gflat, ss)
        assert abs(ss.sqrt().item() - grad.norm().item()) / grad.norm().item() < 1e-5
        ops.adamw_clip(flat, gflat, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, ss)
        torch.cuda.synchronize()
        assert (flat[:n] - ref.data).abs().max().item() < 2e-6, step


def test_data_parallel_gradient_additivity(cuda):
    """Two 'ranks' (half batches, loss scaled by the GLOBAL batch) sum to the single-process gradient -- what the
    NCCL all-reduce of the flat gradient buffer relies on (run sequentially on one GPU)."""
    from from_ddpm_to_stable_diffusion_b200 import TrainerDDPM
    from from_ddpm_to_stable_diffusion_b200.parallel import dp_loss_scale, shard_range
    m, _ = _model(cuda)
    m.train()
    B = 4
    x0, t, y = _inputs(B, 3, 32, seed=5)
    noise = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(6))
    tr = TrainerDDPM(m, 0.0015, 0.0195, 1000).to(cuda)

    def grads(sl):
        for p in m.parameters():
            p.grad = None
        loss = tr(x0[sl].to(cuda), y[sl].to(cuda), t=t[sl].to(cuda), noise=noise[sl].to(cuda)).sum() * dp_loss_scale(B)
        loss.backward()
        return m._engine._flat_grad.clone()

    full = grads(slice(0, B))
    parts = sum(grads(slice(*shard_range(B, r, 2))) for r in range(2))
    rel = ((parts - full).norm() / full.norm()).item()
    # dQ partials are summed by the L2 in arrival order: an fp32 ulp there occasionally flips the bf16 rounding of a
    # dq element (2^-9 relative), which the layers below carry on -- the same batch run twice differs by ~1e-3 as well
    assert rel < 1e-2, rel


def test_fused_sampling_tail_matches_unfused(cuda):
    """tail conv + CFG + posterior update fused (tsd_tail_conv_sample) == tail conv kernel followed by the update kernel."""
    from from_ddpm_to_stable_diffusion_b200 import SamplerDDPM
    m, _ = _model(cuda)
    g = torch.Generator().manual_seed(21)
    xT = torch.randn(2, 3, 64, 64, generator=g).to(cuda)
    z = torch.randn(2, 3, 64, 64, generator=g).to(cuda)
    y = torch.tensor([3, 1]).to(cuda)
    outs = []
    for fused in (True, False):
        s = SamplerDDPM(m, 0.0015, 0.0195, 1000, w=1.8).to(cuda)
        s.fused_tail = fused
        outs.append(s(xT, y, steps=[700, 699], noise_fn=lambda ts: z))
    assert (outs[0] - outs[1]).abs().max().item() < 2e-3
