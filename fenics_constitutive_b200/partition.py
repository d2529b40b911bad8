"""Sharding of the quadrature-point axis: one rank per GPU, no data-path
collective (QPs are independent; in production the shard IS dolfinx's MPI mesh
partition, reference solver/_solver.py:64-68).  Only timing/reporting scalars
are reduced across ranks."""
from __future__ import annotations

import os


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of n QPs owned by `rank`; sizes differ by <= 1
    and shard starts are even (keeps 16-byte alignment of [n] float64 arrays)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    pairs = n // 2
    base, rem = divmod(pairs, world)
    lo = 2 * (rank * base + min(rank, rem))
    hi = 2 * ((rank + 1) * base + min(rank + 1, rem))
    if rank == world - 1:
        hi = n
    return lo, hi


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (
        int(os.environ.get("RANK", "0")),
        int(os.environ.get("LOCAL_RANK", "0")),
        int(os.environ.get("WORLD_SIZE", "1")),
    )


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of one scalar (timing).  No-op without a process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    """SUM all-reduce of one scalar (units processed, residual norms)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
