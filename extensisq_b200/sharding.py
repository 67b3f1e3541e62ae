"""Ensemble sharding across the GPUs of one box.

Lanes are independent, so the data path has NO collective: rank r integrates
the contiguous block ``shard_bounds(N, r, world)`` of lanes.  The only
communication is the optional final gather of per-lane results.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_lanes, rank, world_size):
    """[lo, hi) of the contiguous lane block owned by `rank`; blocks differ in
    size by at most one lane."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_lanes), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_result(tensors, n_lanes, group=None, dst=None):
    """Gather per-lane tensors (dim 0 = this rank's lanes) from all ranks.

    `tensors` is a dict name -> tensor.  Returns the same dict with the full
    ``n_lanes`` leading dimension on every rank (``dst=None``) or on ``dst``
    only (others get None).  Works with NCCL (device tensors) and gloo (CPU
    tensors); shards may differ in size by one lane, so they are padded to the
    largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized():
        return tensors
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_lanes, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    out = {}
    for name, t in tensors.items():
        lo, hi = sizes[rank]
        assert t.shape[0] == hi - lo, (name, t.shape, lo, hi)
        pad = torch.zeros((maxn,) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        pad[:hi - lo] = t
        if dst is None:
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
        else:
            parts = ([torch.empty_like(pad) for _ in range(world)]
                     if rank == dst else None)
            dist.gather(pad, parts, dst=dst, group=group)
        if parts is None:
            out[name] = None
        else:
            out[name] = torch.cat([p[:h - l] for p, (l, h) in
                                   zip(parts, sizes)], dim=0)
    return out
