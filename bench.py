#!/usr/bin/env python
"""Benchmark of the tiny-SD DDPM hot path on B200 (contract: see the task's bench.py section).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference path

Workload (BASELINE.json configs[1]): one DDPM training step of the tiny UNet on synthetic 3x64x64
images, batch 256 per GPU, bf16 activations / fp32 accumulate, dropout 0.1, AdamW + grad clip -- the
reference's step body 02_train_direct.py:64-74.  A "step" = zero_grad, label shift/drop, trainer (timesteps +
q_sample + UNet forward + noise-MSE), backward, [bucketed gradient all-reduce], clip + AdamW, replayed as a
captured CUDA graph (training.GraphedTrainStep).  The same line carries, under "sampling", the other half of
BASELINE's metric: DDPM 64x64 images/s over the reverse process (T=1000, CFG w=1.8; the full 1000 steps are run
and timed at N=1), plus the latent-space figure (configs[4]) and, as context, the reference's own modules run
eagerly on the same GPU ("gpu_eager_baseline") and on the host CPU ("cpu_baseline").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(channel_img=3, channel_multy=[1, 2, 2, 2], channel_base=128, num_class=3, dropout=0.1,
           T=1000, beta_1=0.0015, beta_T=0.0195, w=1.8, lr=2e-6, weight_decay=1e-5, grad_clip=1.0, img=64)
FWD_GF_PER_SAMPLE = 62.60  # necessary forward work, SURVEY 8(d)
BWD_TC_TRAFFIC_B256 = 2.712e9  # dram__bytes_read.sum + dram__bytes_write.sum of one attn_bwd_tc_kernel launch at batch 256
TRAIN_GF_PER_SAMPLE = 187.8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for i, nm in enumerate(names):
                    if r[2 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference legs
def _reference_modules():
    """The reference's own diffusion.py / utils.py, staged under oracle/_ref by oracle/make_ref.py (None if absent)."""
    try:
        from oracle import make_ref
        if make_ref.available():
            return make_ref.load_tiny_sd()
    except Exception:
        pass
    return None


def reference_train_sample(device, n_img, steps, warmup, autocast=None):
    """The UNMODIFIED reference (Diffusion + TrainerDDPM + the step body of 02_train_direct.py:64-74) on `device`."""
    D, U = _reference_modules()
    torch.manual_seed(0)
    model = D.Diffusion(channel_img=CFG["channel_img"], channel_base=CFG["channel_base"], num_class=CFG["num_class"],
                        channel_multy=CFG["channel_multy"], dropout=CFG["dropout"]).to(device)
    opt = torch.optim.AdamW(model.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"])
    trainer = U.TrainerDDPM(model, CFG["beta_1"], CFG["beta_T"], CFG["T"]).to(device)
    g = torch.Generator().manual_seed(1234)
    x0 = torch.randn(n_img, CFG["channel_img"], CFG["img"], CFG["img"], generator=g).to(device)
    y = torch.randint(0, CFG["num_class"], (n_img,), generator=g).to(device)

    def step():
        opt.zero_grad()
        labels = y + 1
        if autocast is not None:
            with torch.autocast(device.type, dtype=autocast):
                loss = trainer(x0, labels).sum() / n_img ** 2.
        else:
            loss = trainer(x0, labels).sum() / n_img ** 2.
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), CFG["grad_clip"])
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    if device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()  # .item() synchronises every step, as the reference loop does
    dt = (time.perf_counter() - t0) / steps
    return n_img / dt, dt


def port_train_sample(n_img, steps, warmup):
    """Fallback when oracle/_ref is absent: the oracle port (functional restatement) of the same step on the CPU."""
    from oracle import ref_unet as R
    sd = R.init_state_dict(0, CFG["channel_img"], CFG["channel_multy"], CFG["channel_base"], CFG["num_class"])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=CFG["lr"], weight_decay=CFG["weight_decay"])
    sched = R.make_schedule(CFG["beta_1"], CFG["beta_T"], CFG["T"])
    g = torch.Generator().manual_seed(1234)
    x0 = torch.randn(n_img, CFG["channel_img"], CFG["img"], CFG["img"], generator=g)
    y = torch.randint(1, CFG["num_class"] + 1, (n_img,), generator=g)

    def step():
        opt.zero_grad()
        t = torch.randint(CFG["T"], (n_img,))
        noise = torch.randn_like(x0)
        loss = R.trainer_loss(params, sched, x0, y, t, noise, CFG["channel_multy"], CFG["channel_base"],
                              use_sdpa=True).sum() / n_img ** 2
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), CFG["grad_clip"])
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_img / dt, dt


def cpu_train_sample(n_img, steps, warmup, threads):
    """(samples/s, s/step, kind): the reference path on the host cores, on a bounded sample of the workload."""
    torch.set_num_threads(threads)
    if _reference_modules() is not None:
        v, dt = reference_train_sample(torch.device("cpu"), n_img, steps, warmup)
        return v, dt, "reference"
    v, dt = port_train_sample(n_img, steps, warmup)
    return v, dt, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_img = 8
    steps = max(1, min(args.steps, 5))
    warmup = 1 if args.warmup > 0 else 0
    val, dt, kind = cpu_train_sample(n_img, steps, warmup, threads)
    what = ("the unmodified reference modules (oracle/_ref: diffusion.py + utils.py, step body of 02_train_direct.py:64-74)"
            if kind == "reference" else "oracle port of the reference path")
    line = {
        "impl": "reference", "metric": "train_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"tiny UNet DDPM training step 3x64x64 on the host CPU, {what}, "
                               f"bounded sample: batch {n_img} per step", "global_batch": n_img, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": kind,
                         "sample": f"{steps} training step(s) of batch {n_img} (fwd+bwd+clip+AdamW), torch CPU fp32"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import ctypes

    import torch.distributed as dist
    from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM, _lib
    from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW
    from from_ddpm_to_stable_diffusion_b200.parallel import set_shard
    from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep, train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(0)  # same weights and same Philox seeds on every rank; the streams are offset by global sample index
    model = Diffusion(CFG["channel_img"], CFG["channel_multy"], CFG["channel_base"], num_class=CFG["num_class"],
                      dropout=CFG["dropout"]).to(dev).train()
    trainer = TrainerDDPM(model, CFG["beta_1"], CFG["beta_T"], CFG["T"]).to(dev)
    opt = FusedClipAdamW(model, lr=CFG["lr"], weight_decay=CFG["weight_decay"], max_norm=CFG["grad_clip"])
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, CFG["channel_img"], CFG["img"], CFG["img"], generator=g).pin_memory()
    y_host = torch.randint(0, CFG["num_class"], (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    global_b = B * world
    overlap = bool(int(os.environ.get("TSD_DP_OVERLAP", "1")))
    stepper = GraphedTrainStep(trainer, opt, train_rand=0.05, overlap=overlap)
    issue_mode = "graph"
    try:
        stepper(x_dev, y_dev)  # captures the graphs (with the NCCL all-reduce inside when world > 1)
    except Exception as e:  # never lose the benchmark line to a capture problem: fall back, and say so in the config
        print(f"[bench] graph capture failed ({type(e).__name__}: {e}); falling back", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        try:
            overlap = False
            stepper = GraphedTrainStep(trainer, opt, train_rand=0.05, overlap=False)
            stepper(x_dev, y_dev)
            issue_mode = "graph, all-reduce between graphs (capture with NCCL failed)"
        except Exception as e2:
            print(f"[bench] second capture failed ({type(e2).__name__}: {e2}); eager stepping", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
            issue_mode = "eager (graph capture failed)"

            class _Eager:
                loss = torch.zeros((), device=dev)

                def __call__(self, x, y):
                    set_shard(trainer, rank * B)
                    self.loss = train_step(trainer, opt, x.to(dev, non_blocking=True), y.to(dev, non_blocking=True),
                                           train_rand=0.05).detach()
                    return self.loss

                @staticmethod
                def launches_per_step():
                    return 0

            stepper = _Eager()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    lib = _lib.lib()
    lib.tsd_launch_count.restype = ctypes.c_ulonglong
    for _ in range(max(args.warmup, 1)):
        stepper(x_dev, y_dev)  # the first call captures the graphs
    clocks = ClockSampler(local)
    clocks.start()
    n0 = lib.tsd_launch_count()
    ms = timed(lambda: stepper(x_dev, y_dev), args.steps)
    clk = clocks.stop()
    # replayed graph nodes are counted at capture time; eager launches by the library's own counter
    launches = stepper.launches_per_step() * args.steps + (lib.tsd_launch_count() - n0)
    ms_per_step = ms / args.steps
    value = global_b / (ms_per_step / 1e3)
    assert bool(torch.isfinite(stepper.loss)), "training loss is not finite"

    # end-to-end: pinned host inputs -> H2D every step, loss read back every step (02_train_direct.py:66-74)
    # The loss of step i is copied to pinned host memory right after the step and read by the host while step i+1
    # is already queued (one step of slack, as a training loop that logs the loss would do), so the device never idles
    # waiting for the host; every step's inputs are copied in and every step's loss is read out inside the timed region.
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "losses": []}

    def e2e_step():
        i = e2e_state["i"]
        loss = stepper(x_host, y_host)  # copies the pinned host batch into the graph's static buffers (H2D, async)
        loss_host[i & 1].copy_(loss.detach(), non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            e2e_state["losses"].append(float(loss_host[(i - 1) & 1]))
        e2e_state["i"] = i + 1

    e2e_ms = timed(e2e_step, args.steps) / args.steps  # the closing synchronize of timed() covers the last read
    e2e_state["losses"].append(float(loss_host[(e2e_state["i"] - 1) & 1]))
    assert len(e2e_state["losses"]) == args.steps and all(v == v for v in e2e_state["losses"])
    e2e_val = global_b / (e2e_ms / 1e3)

    # the same iteration issued eagerly from Python (one C-ABI call per kernel), for the launch-overhead comparison
    set_shard(trainer, rank * B)
    for _ in range(2):
        train_step(trainer, opt, x_dev, y_dev, train_rand=0.05)
    eager_ms = timed(lambda: train_step(trainer, opt, x_dev, y_dev, train_rand=0.05), 3) / 3

    # Per-kernel rooflines, measured live: one extra (untimed) eager step with a CUDA-event pair around every C-ABI call
    # on the launching stream, aggregated per entry point and shape.
    hbm, tf, how = peaks()
    _lib.PROFILE = True
    _lib.profile_report()
    train_step(trainer, opt, x_dev, y_dev, train_rand=0.0)
    agg = _lib.profile_report()
    _lib.PROFILE = False
    tot_ms = sum(v[1] for v in agg.values())

    def fam(name, pred=lambda k: True):
        n = sum(v[0] for (nm, k), v in agg.items() if nm == name and pred(k))
        t = sum(v[1] for (nm, k), v in agg.items() if nm == name and pred(k))
        return n, t

    # dominant kernel by time: self-attention at L = 4096, head_dim 16 (forward + backward launches)
    L, Cc, Hh = 4096, 128, 8
    nb, tb = fam("tsd_attn_bwd_ws", lambda k: k[1] == L)
    nf, tfw = fam("tsd_attn_fwd_ws", lambda k: k[1] == L)
    att_flops = 4.0 * B * L * L * Cc  # QK^T + PV per forward launch; backward = 2.5x (5 matmuls)
    att_exps = float(B) * Hh * L * L  # per pass; the one-pass backward evaluates them once
    ach = (nf * att_flops + nb * 2.5 * att_flops) / ((tfw + tb) * 1e-3) / 1e12
    roof = {"kernel": "attn_fwd_tc_kernel + attn_bwd_tc_kernel (tcgen05/TMEM; one-pass backward), L=4096, head_dim 16, %d launches/step" % (nf + nb),
            "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
            # DRAM bytes per launch (read + write) of the backward kernel at this shape, from ncu
            # (profiles/r01_attn_bwd_tc_dram_traffic_b256.csv); the forward kernel moves 1.088e9 (qkv once, out + lse once)
            "traffic": BWD_TC_TRAFFIC_B256 if B == 256 else None,
            "peak_source": how + " (bf16_tflops_sustained)", "share_of_step": (tfw + tb) / tot_ms,
            "avg_launch_ms": {"fwd": tfw / max(nf, 1), "bwd": tb / max(nb, 1)},
            "exp_bound": {"achieved_texp_s": (nf + nb) * att_exps / ((tfw + tb) * 1e-3) / 1e12, "mufu_peak_texp_s": 4.64,
                          "note": "head_dim 16 makes the kernel MUFU-exp bound, not tensor bound (tools/ex2_bench.cu)"}}
    extra = []
    n1, t1 = fam("tsd_conv3x3_fwd", lambda k: k[0] == 128 and k[1] == 0 and k[3] == 64 and k[6] == 128)
    if n1:
        fl = 2.0 * B * 64 * 64 * 128 * 9 * 128
        extra.append({"kernel": "gemm_tc_kernel<0,0,0,0> conv3x3 128->128 @64x64 (tcgen05 implicit GEMM, halo mode)", "bound": "tensor",
                      "achieved": fl * n1 / (t1 * 1e-3) / 1e12, "peak": tf, "unit": "TFLOP/s",
                      "frac": fl * n1 / (t1 * 1e-3) / 1e12 / tf, "launches": n1})
    n2, t2 = fam("tsd_gn_apply", lambda k: k[0] == 128 and k[1] == 0 and k[3] == 4096)
    if n2:
        by = 2.0 * B * 4096 * 128 * 2
        extra.append({"kernel": "gn_apply_kernel C=128 @64x64 (GroupNorm+SiLU(+dropout))", "bound": "hbm",
                      "achieved": by * n2 / (t2 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                      "frac": by * n2 / (t2 * 1e-3) / 1e9 / hbm, "launches": n2})
    n3, t3 = fam("tsd_adamw_clip_dev")
    if n3:
        by = 30945155 * 7 * 4.0
        extra.append({"kernel": "adamw_clip_kernel (clip + AdamW, 30.9 M params)", "bound": "hbm",
                      "achieved": by / (t3 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / (t3 * 1e-3) / 1e9 / hbm,
                      "launches": n3})
    step_tf = TRAIN_GF_PER_SAMPLE * B / (ms_per_step / 1e3) / 1e3
    del stepper
    torch.cuda.empty_cache()

    # BASELINE configs[3]: data-parallel training at a FIXED global batch of 2048 on 1 / 2 / 4 GPUs (8 GPUs: the main line
    # above is that configuration).  2048 / N images per rank are processed as micro-batches of `B` with gradient
    # accumulation; one all-reduce and one optimiser step per global batch.
    cfg3 = None
    if world in (1, 2, 4) and 2048 % (world * B) == 0 and args.cfg3_steps > 0:
        k = 2048 // (world * B)
        n3 = args.cfg3_steps if world > 1 else min(args.cfg3_steps, 2)
        try:  # a side line: never lose the benchmark line to it
            g3 = torch.Generator().manual_seed(4321 + rank)
            x3 = torch.randn(k * B, CFG["channel_img"], CFG["img"], CFG["img"], generator=g3).pin_memory()
            y3 = torch.randint(0, CFG["num_class"], (k * B,), generator=g3).pin_memory()
            st3 = GraphedTrainStep(trainer, opt, train_rand=0.05, micro_batches=k, overlap=overlap)
            st3(x3, y3)
            ms3 = timed(lambda: st3(x3, y3), n3) / n3
            assert bool(torch.isfinite(st3.loss)), "loss of the global-batch-2048 step is not finite"
            cfg3 = {"global_batch": 2048, "micro_batches_per_rank": k, "micro_batch": B, "ms_per_step": ms3,
                    "samples_per_s": 2048 / (ms3 / 1e3), "timed_steps": n3,
                    "note": "BASELINE configs[3]; inputs from pinned host memory every micro-batch"}
            del st3, x3, y3
        except Exception as e:
            if world > 1:
                raise  # a rank that drops out of a collective would hang the others: fail loudly instead
            cfg3 = {"failed": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.synchronize()
        torch.cuda.empty_cache()

    # sampling throughput (same model, eval mode): DDPM 64x64 images/s, CFG w=1.8, 2 forwards per step.  At N=1 the whole
    # reverse process (T=1000 steps) is executed and timed; at N>1 a prefix of `--sample-steps-multi` steps is timed and
    # scaled to T (every reverse step costs the same: same kernels, same shapes).
    sampling = None
    k = args.sample_steps if world == 1 else min(args.sample_steps, args.sample_steps_multi)
    if k > 0:
        model.eval()
        Bs = args.sample_batch
        sampler = SamplerDDPM(model, CFG["beta_1"], CFG["beta_T"], CFG["T"], w=CFG["w"]).to(dev)
        set_shard(sampler, rank * Bs)
        xT = torch.randn(Bs, CFG["channel_img"], CFG["img"], CFG["img"], device=dev)
        ys = torch.randint(1, CFG["num_class"] + 1, (Bs,), device=dev)
        k = min(k, CFG["T"])
        sampler(xT, ys, steps=range(CFG["T"] - 1, CFG["T"] - 1 - 4, -1))  # capture + warm-up
        n0 = lib.tsd_launch_count()
        out = []
        s_ms = timed(lambda: out.append(sampler(xT, ys, steps=range(CFG["T"] - 1, CFG["T"] - 1 - k, -1))), 1)
        assert bool(torch.isfinite(out[0]).all())
        per = s_ms / k
        sampling = {"metric": "ddpm_64x64_images_per_sec_T1000", "images_per_s": Bs * world / (per / 1e3 * CFG["T"]),
                    "ms_per_reverse_step": per, "batch_per_gpu": Bs, "T": CFG["T"], "w": CFG["w"],
                    "timed_reverse_steps": k, "full_reverse_process_timed": k == CFG["T"],
                    "wall_s_timed": s_ms / 1e3,
                    # executed work: the label-independent prefix (11.57 GF: head conv, first ResBlock, first attention
                    # block up to its self-attention) runs once for the conditional / unconditional pair
                    "gflop_per_image_step_executed": 2 * FWD_GF_PER_SAMPLE - 11.57,
                    "tflops": (2 * FWD_GF_PER_SAMPLE - 11.57) * Bs / (per / 1e3) / 1e3}
        # the fused sampling tail (final conv + CFG + posterior update + Philox noise, the north star's "update kernel")
        # against the HBM roofline: two eager reverse steps with per-call CUDA events
        if rank == 0:
            ug = getattr(sampler, "use_cuda_graph", True)
            sampler.use_cuda_graph = False
            _lib.PROFILE = True
            _lib.profile_report()
            sampler(xT, ys, steps=range(CFG["T"] - 1, CFG["T"] - 3, -1))
            agg_s = _lib.profile_report()
            _lib.PROFILE = False
            sampler.use_cuda_graph = ug
            nt_ = sum(v[0] for (nm, _k), v in agg_s.items() if nm == "tsd_tail_conv_sample")
            tt_ = sum(v[1] for (nm, _k), v in agg_s.items() if nm == "tsd_tail_conv_sample")
            if nt_:
                hw_ = CFG["img"] * CFG["img"]
                by = 2.0 * Bs * hw_ * 128 * 2 + 3.0 * Bs * CFG["channel_img"] * hw_ * 4
                extra.append({"kernel": "tail_conv_y_kernel<3,1> fused sampling tail (128->3 conv + CFG + posterior update + "
                              "Philox noise), %d image pairs" % Bs, "bound": "hbm", "achieved": by * nt_ / (tt_ * 1e-3) / 1e9,
                              "peak": hbm, "unit": "GB/s", "frac": by * nt_ / (tt_ * 1e-3) / 1e9 / hbm, "launches": nt_})
        del sampler, out
        torch.cuda.empty_cache()
        # BASELINE configs[4]: latent-space DDPM, 4x16x16 latents, batch 4096 (sharded over the ranks), num_class 10
        if args.latent_steps > 0:
            torch.manual_seed(1)
            lm = Diffusion(4, CFG["channel_multy"], CFG["channel_base"], num_class=10, dropout=0.0).to(dev).eval()
            ls = SamplerDDPM(lm, CFG["beta_1"], CFG["beta_T"], CFG["T"], w=CFG["w"]).to(dev)
            Bl = 4096 // world
            set_shard(ls, rank * Bl)
            zT = torch.randn(Bl, 4, 16, 16, device=dev)
            yl = torch.randint(1, 11, (Bl,), device=dev)
            kl = args.latent_steps
            ls(zT, yl, steps=range(CFG["T"] - 1, CFG["T"] - 1 - 4, -1))
            l_ms = timed(lambda: ls(zT, yl, steps=range(CFG["T"] - 1, CFG["T"] - 1 - kl, -1)), 1) / kl
            sampling["latent_4x16x16"] = {"images_per_s": 4096 / (l_ms / 1e3 * CFG["T"]), "ms_per_reverse_step": l_ms,
                                          "global_batch": 4096, "timed_reverse_steps": kl,
                                          "note": "BASELINE configs[4]; scaled from the timed prefix to T=1000"}
            del ls, lm
            torch.cuda.empty_cache()
        model.train()

    cpu = gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if _reference_modules() is not None:
            # context: the reference's own modules run eagerly on this GPU (the ATen/cuDNN/SDPA kernel set it would hit)
            gpu_eager = {"batch": 32, "unit": "samples/s", "what": "unmodified reference modules (oracle/_ref), eager "
                         "torch on this GPU, training step of batch 32 (fwd+bwd+clip+AdamW)"}
            for name, ac in (("fp32_tf32conv", None), ("bf16_autocast", torch.bfloat16)):
                try:
                    v, _ = reference_train_sample(dev, 32, 5, 2, autocast=ac)
                    gpu_eager[name] = v
                except Exception as e:  # context only: never fails the benchmark
                    gpu_eager[name] = f"failed: {type(e).__name__}: {e}"[:200]
                torch.cuda.empty_cache()
        threads = os.cpu_count() or 1
        v, dt, kind = cpu_train_sample(8, 3, 1, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": kind,
               "sample": "1 warm-up + 3 timed training steps of batch 8 (fwd+bwd+clip+AdamW), "
                         + ("unmodified reference modules (oracle/_ref)" if kind == "reference" else "oracle port")
                         + " on torch CPU fp32, all host threads"}

    if rank == 0:
        line = {
            "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"tiny UNet DDPM training step bf16 3x64x64, batch {B} per GPU (BASELINE configs[1])",
                       "global_batch": global_b, "parallelism": f"dp{world}", "dropout": CFG["dropout"],
                       "optimizer": "clip_grad_norm(1.0)+AdamW fused", "l2": "inputs larger than L2 (activations >> 126 MB)",
                       "issue": ("whole iteration replayed as a CUDA graph"
                                 + ("" if world == 1 else (", bucketed NCCL all-reduce captured inside the backward"
                                                           if overlap else ", NCCL all-reduce between two graphs")))
                       if issue_mode == "graph" else issue_mode},
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "sampling": sampling,
            "cfg3_global_batch_2048": cfg3,
            "eager_issue_ms_per_step": eager_ms,
            "roofline": roof,
            "roofline_other_kernels": extra,
            "step_necessary_tflops": step_tf,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="training batch per GPU")
    ap.add_argument("--sample-batch", type=int, default=256, help="images per GPU for the sampling figure")
    ap.add_argument("--sample-steps", type=int, default=1000, help="timed reverse steps at N=1 (0 = skip sampling figure)")
    ap.add_argument("--sample-steps-multi", type=int, default=200, help="timed reverse steps when N>1")
    ap.add_argument("--latent-steps", type=int, default=50, help="timed reverse steps of the latent config (0 = skip)")
    ap.add_argument("--cfg3-steps", type=int, default=3, help="timed steps of the fixed-global-batch-2048 line at N=1/2/4 (0 = skip; at most 2 at N=1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
