"""Cloud-level sharding of the operators over the GPUs of one box.

Every operator on this path works on independent clouds, so N ranks simply own contiguous ranges
of clouds: no data-path collective.  These helpers do the bookkeeping for both layouts."""
from typing import Tuple

import torch


def cloud_range(num_clouds: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous, balanced split: the first (num_clouds % world) ranks own one extra cloud"""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    base, extra = divmod(num_clouds, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_dense(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """DENSE layout [B, ...] -> this rank's clouds"""
    lo, hi = cloud_range(t.shape[0], world, rank)
    return t[lo:hi]


def shard_batched(pos: torch.Tensor, batch: torch.Tensor, world: int, rank: int, *features: torch.Tensor):
    """PARTIAL_DENSE / MESSAGE_PASSING layout (pos [N,3], sorted batch [N]) -> this rank's rows, with the
    batch vector renumbered from 0 and the global row offset (add it to local indices to get global ones)"""
    num_clouds = int(batch[-1].item()) + 1 if batch.numel() else 0
    lo, hi = cloud_range(num_clouds, world, rank)
    bounds = torch.searchsorted(batch, torch.tensor([lo, hi], dtype=batch.dtype, device=batch.device))
    r0, r1 = int(bounds[0]), int(bounds[1])
    out = [pos[r0:r1], batch[r0:r1] - lo, r0]
    out.extend(f[r0:r1] for f in features)
    return tuple(out)


def max_over_ranks(value: float, device=None) -> float:
    """the slowest rank's time (multi-GPU numbers are the max over ranks, never a wall clock)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
