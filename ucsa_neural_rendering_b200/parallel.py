"""Multi-GPU plumbing of SURVEY.md section 8(e): one process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  Host logic only -- no kernel is launched here.

* training  : rays shard (contiguous slice per rank), parameters replicate; ONE all-reduce(sum) over the flat
              gradient buffer per step, identical optimizer step on every rank, no broadcast afterwards;
* occupancy : every 16 steps all-reduce(max) of the density grid (ranks see different samples);
* rendering : views shard round-robin (view % world == rank), no collective on the data path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous slice [lo, hi) of `n_items` for `rank`; sizes differ by at most one, every item exactly once."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def views_of_rank(n_views: int, rank: int, world_size: int):
    """Round-robin view assignment for full-frame rendering (config 3)."""
    return list(range(rank, n_views, world_size))


def all_reduce_gradients(flat_grad: torch.Tensor):
    """Sum the flat gradient buffer over ranks (losses are pre-scaled by 1/world, so the sum is the global mean)."""
    _, w = world()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def sync_density_grid(density_grid: torch.Tensor):
    """Element-wise max over ranks of the occupancy grid; returns the (identical on all ranks) mean density."""
    _, w = world()
    if w > 1:
        dist.all_reduce(density_grid, op=dist.ReduceOp.MAX)
    return float(density_grid.clamp(min=0).mean())


def max_over_ranks(value: float, device=None) -> float:
    _, w = world()
    if w == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
