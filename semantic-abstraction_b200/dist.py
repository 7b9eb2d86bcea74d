"""Multi-GPU plumbing of the hot path: inference shards by independent units — images for the
relevancy extractor, voxel grids for the UNet — with NO data-path collective. One process per GPU
(torch.distributed, NCCL on GPUs / gloo in the CPU tests); the only reductions are the timing max and the unit
count used by bench.py."""
from __future__ import annotations

import os
from typing import List, Sequence, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_units(units: Sequence[T], rank: int, world: int) -> List[T]:
    """Contiguous block partition (rank r gets units [r*n/world, (r+1)*n/world)): every unit is processed exactly
    once, block sizes differ by at most one, order inside a rank is preserved (the relevancy assembly is order
    dependent *within* an image, never across images)."""
    n = len(units)
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    return list(units[lo:hi])


def max_over_ranks(value: float, device="cpu") -> float:
    """Wall/device time of a step = the slowest rank's."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
