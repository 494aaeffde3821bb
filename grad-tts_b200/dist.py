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


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_batch(local, n_items, group=None, out=None):
    """Gather per-rank shards (dim 0, sizes from shard_counts) into the full batch on every rank.

    One collective.  When the batch divides evenly the shards land straight in the output tensor
    (`all_gather_into_tensor`, no staging copies); otherwise every shard is padded to the largest count, gathered into one
    staging tensor and compacted.  `out` (n_items, ...) may be supplied to avoid the allocation."""
    world, rank = _world(group)
    if world == 1:
        return local
    counts = shard_counts(n_items, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {counts[rank]} of {n_items}")
    tail = tuple(local.shape[1:])
    if out is None:
        out = torch.empty((n_items,) + tail, dtype=local.dtype, device=local.device)
    cmax = max(counts)
    if min(counts) == cmax:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    stage = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(stage, pad, group=group)
    lo = 0
    for r, c in enumerate(counts):
        out[lo:lo + c] = stage[r * cmax: r * cmax + c]
        lo += c
    return out


def sharded_call(fn, batch_tensors, n_items, gather=True, group=None):
    """Run `fn(*shards)` on this rank's slice of every tensor in `batch_tensors` (None entries pass through)
    and all-gather the result.  `fn` is e.g. `lambda z, mask, mu, spk: decoder(z, mask, mu, n, False, spk)`.

    The decision to refuse a batch smaller than the world size is taken identically on EVERY rank before any work or
    collective is issued (a rank-local raise would leave the other ranks hanging in the all-gather)."""
    world, rank = _world(group)
    if n_items < world:
        raise RuntimeError(f"cannot shard {n_items} item(s) over {world} ranks: every rank needs at least one "
                           "(run fewer ranks, or let one rank take the call)")
    lo, hi = shard_bounds(n_items, world, rank)
    shards = [None if t is None else t[lo:hi].contiguous() for t in batch_tensors]
    local = fn(*shards)
    return all_gather_batch(local, n_items, group) if gather else local
