"""world_size-2 gloo test of the data-parallel host logic (CPU)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from from_ddpm_to_stable_diffusion_b200 import parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 7
    g = torch.Generator().manual_seed(3)
    full = torch.randn(total, 5, generator=g)
    b, e = P.shard_range(total, rank, world)
    local = full[b:e]
    # "per-rank gradient" of sum(x^2) * dp_loss_scale(global batch): the sum over ranks must equal the 1-process one
    grad_local = (2 * local * P.dp_loss_scale(total)).sum(0)
    P.all_reduce_sum_(grad_local)
    want = (2 * full * P.dp_loss_scale(total)).sum(0)
    ok1 = torch.allclose(grad_local, want, atol=1e-6)
    gathered = P.gather_batch(local * 2, total)
    ok2 = torch.equal(gathered, full * 2)
    q.put((rank, ok1, ok2, (b, e)))
    dist.destroy_process_group()


def test_dp_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == (0, 4) and res[1][3] == (4, 7)


def test_shard_range_covers_everything():
    from from_ddpm_to_stable_diffusion_b200.parallel import shard_range
    for total in (1, 7, 256, 2048):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
