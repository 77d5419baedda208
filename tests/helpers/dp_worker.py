"""Two-rank NCCL check of the data-parallel path, launched by tests/test_dp_nccl_gpu.py through torchrun.

Every rank also computes the single-process answer on the full global batch, then its own shard through the
data-parallel code path (rank-offset random streams, loss normalised by the global batch, NCCL all-reduce of the flat
gradient) and compares:
  (1) eager train_step with the bucketed all-reduce overlapped with backward,
  (2) GraphedTrainStep with the all-reduce captured inside the graph, (3) with the all-reduce issued between graphs,
  (4) sampling: the shard of a global batch equals the same rows of the single-process result (no collective).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_unet as R  # noqa: E402  (deterministic weights only)
from from_ddpm_to_stable_diffusion_b200 import Diffusion, SamplerDDPM, TrainerDDPM  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.parallel import gather_batch, set_shard, shard_range  # noqa: E402
from from_ddpm_to_stable_diffusion_b200.training import GraphedTrainStep, train_step  # noqa: E402

MULTY = [1, 2, 2, 2]


class NoDrop:
    @staticmethod
    def rand():
        return 1.0


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def build(dev, dropout):
    torch.manual_seed(4321)
    sd = R.init_state_dict(0, 3, MULTY, 128, 3)
    m = Diffusion(3, MULTY, 128, num_class=3, dropout=dropout)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    tr = TrainerDDPM(m, 0.0015, 0.0195, 1000).to(dev)
    opt = FusedClipAdamW(m, lr=1e-4, weight_decay=1e-5, max_norm=1.0)
    return m, tr, opt


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = 8
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 3, (B,), generator=g).to(dev)
    lo, hi = shard_range(B, rank, world)

    # ---- single-process reference on the full batch (no process group involvement: plain python path)
    m0, tr0, opt0 = build(dev, 0.1)
    opt0.zero_grad()
    loss0 = tr0(x, y + 1).sum() / B ** 2
    loss0.backward()
    g0 = m0._engine._flat_grad.clone()
    opt0.step()
    p0 = opt0.flat_p.clone()

    def check(tag, loss_local, m, opt, grads_clipped):
        tot = loss_local.detach().clone()
        dist.all_reduce(tot)
        assert abs(tot.item() - loss0.item()) / abs(loss0.item()) < 1e-5, (tag, tot.item(), loss0.item())
        gg = m._engine._flat_grad
        ref = opt0.engine._flat_grad if grads_clipped else g0  # after step() the buffer holds the clipped gradient
        assert rel(gg, ref) < 1e-2, (tag, rel(gg, ref))  # dQ arrival-order sums flip a few bf16 roundings (~1e-3)
        assert rel(opt.flat_p, p0) < 2e-3, (tag, rel(opt.flat_p, p0))  # one +-lr AdamW step on re-associated gradients
        other = opt.flat_p.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(other, opt.flat_p), tag + ": ranks diverged"

    # (1) eager step, bucketed all-reduce overlapped with backward
    m1, tr1, opt1 = build(dev, 0.1)
    l1 = train_step(tr1, opt1, x[lo:hi], y[lo:hi], rng=NoDrop)
    check("eager", l1, m1, opt1, True)
    # (2) captured iteration, NCCL inside the graph   (3) all-reduce between two graphs
    for tag, overlap in (("graph+nccl", True), ("graph|nccl", False)):
        m2, tr2, opt2 = build(dev, 0.1)
        st = GraphedTrainStep(tr2, opt2, overlap=overlap, rng=NoDrop)
        l2 = st(x[lo:hi], y[lo:hi]).clone()
        check(tag, l2, m2, opt2, True)
    # (4) sampling shards, no collective in the loop.  The model must be the same on every rank: m1 went through the
    # data-parallel step above (check() asserted bit-identical weights across ranks), whereas every rank trained its own
    # m0 on the full batch and dQ's arrival-order sums make those copies differ in the last bits.
    m0 = m1
    m0.eval()
    xT = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    ys = torch.randint(1, 4, (B,), generator=g).to(dev)
    steps = range(999, 991, -1)
    s_full = SamplerDDPM(m0, 0.0015, 0.0195, 1000, w=1.8).to(dev)
    full = s_full(xT, ys, steps=steps)
    s_part = SamplerDDPM(m0, 0.0015, 0.0195, 1000, w=1.8).to(dev)
    set_shard(s_part, lo)
    part = s_part(xT[lo:hi], ys[lo:hi], steps=steps)
    gathered = gather_batch(part, B)
    assert torch.equal(gathered, full), (gathered - full).abs().max().item()
    dist.barrier()
    if rank == 0:
        print("DP_NCCL_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
