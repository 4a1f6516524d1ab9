"""Sharding of a batch of clouds across data-parallel ranks (the only partitioning on the path).

The reference splits `total_bs // world_size` clouds per rank (main.py:72-79) through a DistributedSampler
(tools/builder.py:23-24); forward needs no collective.  Helpers here are device-agnostic so the N > 1 logic is
covered on CPU with the gloo backend (tests/test_dist_cpu.py).
"""

from __future__ import annotations

import torch


def shard_range(n_clouds: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of `n_clouds` for `rank`; the first n % world ranks take one extra cloud."""
    base, rem = divmod(n_clouds, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Device-time reduction used by bench.py: the slowest rank defines the step time."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_predictions(local: torch.Tensor) -> torch.Tensor:
    """utils/dist_utils.py:50-54 `gather_tensor`: concatenate per-rank predictions (validation only)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, dim=0)
