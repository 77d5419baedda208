"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the B200 box, gloo in CPU tests).

The reference is single-process (02_train_direct.py:31); the hot path shards by batch only:
  * sampling: images are independent -> each rank samples its slice, no collective in the loop; the final
    images are gathered once;
  * training: data parallel -> one sum all-reduce of the flat fp32 gradient buffer per step, with the
    reference's loss normalisation (sum / B^2, 02_train_direct.py:70) taken over the GLOBAL batch.
"""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank, world):
    """Contiguous [begin, end) slice of `total` items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_reduce_sum_(flat, group=None):
    """In-place sum of a flat buffer over ranks (gradient exchange); no-op for a single process."""
    rank, world = world_info()
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def gather_batch(local, total, group=None):
    """Concatenate per-rank slices (made with shard_range) back into the full batch on every rank."""
    rank, world = world_info()
    if world == 1:
        return local
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(e - b for b, e in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: e - b] for o, (b, e) in zip(outs, sizes)], dim=0)


def set_shard(obj, sample0):
    """Tell a TrainerDDPM / SamplerDDPM (and the UNet engine behind it) the GLOBAL index of its first local sample:
    random numbers are keyed by global sample index, so N ranks draw what one rank would on the same global batch."""
    obj.rng.set_sample0(sample0)
    eng = getattr(getattr(obj, "model", None), "_engine", None)
    if eng is not None:
        eng.rng.set_sample0(sample0)


def dp_loss_scale(global_batch):
    """The reference normalises the summed loss by bs**2 (02_train_direct.py:70); under data parallelism every
    rank uses the global batch so that the all-reduced (summed) gradient equals the single-process one."""
    return 1.0 / float(global_batch) ** 2
