"""Slice-wise sharding of a volume across ranks (one process per GPU) and the final gather.

Slices are independent on this path (per-slice statistics), so there is no collective
on the data path: every rank denoises its contiguous range and the results are gathered once at the end.
"""
import torch
import torch.distributed as dist


def shard_range(n_slices, rank, world):
    """Contiguous [start, stop) of `rank`; remainders go to the first ranks."""
    base, rem = divmod(int(n_slices), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_slices(local, n_slices, rank=None, world=None):
    """All ranks receive the [n_slices, ...] result; ragged shards are padded to the largest shard for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    spans = [shard_range(n_slices, r, world) for r in range(world)]
    big = max(e - s for s, e in spans)
    padded = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * big: r * big + (e - s)] for r, (s, e) in enumerate(spans)])
