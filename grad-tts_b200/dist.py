"""Multi-GPU partitioning of the hot path: independent utterances / n-best samples are split across ranks.

Every op of the decoder is per-sample (GroupNorm per sample, softmax per (sample, head)), so the batch shards
with no data-path collective; the only communication is one all-gather of the finished mels
(SURVEY 8e).  One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced split: the first n_items % world_size ranks get one extra item.
    100 samples over 8 ranks -> 13,13,13,13,12,12,12,12 (BASELINE config 4)."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


def all_gather_batch(local, n_items, group=None):
    """Gather per-rank shards (dim 0, sizes from shard_counts) into the full batch on every rank.

    Uses one fixed-size all_gather (shards padded to the largest count), the collective of SURVEY K11."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(n_items, world)
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def sharded_call(fn, batch_tensors, n_items, gather=True, group=None):
    """Run `fn(*shards)` on this rank's slice of every tensor in `batch_tensors` (None entries pass through)
    and all-gather the result.  `fn` is e.g. `lambda z, mask, mu, spk: decoder(z, mask, mu, n, False, spk)`."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(n_items, world, rank)
    shards = [None if t is None else t[lo:hi].contiguous() for t in batch_tensors]
    local = fn(*shards) if hi > lo else None
    if local is None:
        raise RuntimeError("empty shard: more ranks than items")
    return all_gather_batch(local, n_items, group) if gather else local
