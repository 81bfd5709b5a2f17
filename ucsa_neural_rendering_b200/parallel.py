"""Multi-GPU plumbing of SURVEY.md section 8(e): one process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  Host logic only -- no kernel is launched here.

* training  : rays shard (contiguous slice per rank), parameters replicate.  On one NVLink domain the gradient exchange
              is fused into the optimizer kernel over symmetric (peer-mapped, multicast) memory: each rank sums and
              updates 1/world of the parameters and writes them to every rank (ucsa_adam_exchange, PeerExchange
              below).  Fallback: ONE NCCL all-reduce(sum) over the flat gradient buffer + identical Adam everywhere;
* occupancy : every 16 steps all-reduce(max) of the density grid (ranks see different samples);
* rendering : views shard round-robin (view % world == rank), no collective on the data path.
"""
from __future__ import annotations

import os

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


def owner_slice(n_params: int, rank: int, world_size: int):
    """Parameters [lo, hi) that `rank` sums, updates and broadcasts in the fused exchange: equal slices of whole
    float4 groups, every parameter exactly once."""
    quads = (n_params + 3) // 4
    per = (quads + world_size - 1) // world_size * 4
    lo = min(rank * per, n_params)
    return lo, min(lo + per, n_params)


class PeerExchange:
    """Symmetric-memory buffers of one flat parameter space (gradients, fp32 masters, fp16 working copy) and the peer /
    multicast addresses ucsa_adam_exchange needs.  torch.distributed._symmetric_memory is the plumbing (allocation,
    handle exchange, the cross-rank barrier kernel); the data path is our kernel."""

    def __init__(self, n_params: int, device):
        import torch.distributed._symmetric_memory as symm

        self.rank, self.world = world()
        self.n = n_params
        self.grad = symm.empty(n_params, dtype=torch.float32, device=device)
        self.param = symm.empty(n_params, dtype=torch.float32, device=device)
        self.param_h = symm.empty(n_params, dtype=torch.float16, device=device)
        self.grad.zero_()
        # per-rank overflow flag of the step (ucsa_grad_check writes it, every rank's ucsa_adam_exchange reads all)
        self.flag = symm.empty(4, dtype=torch.float32, device=device)
        self.flag.zero_()
        self._handles = [symm.rendezvous(t, dist.group.WORLD) for t in (self.grad, self.param, self.param_h, self.flag)]
        self.flag_ptrs, _ = self._addresses(self.flag, self._handles[3])
        self.grad_ptrs, self.mc_grad = self._addresses(self.grad, self._handles[0])
        self.param_ptrs, self.mc_param = self._addresses(self.param, self._handles[1])
        self.param_h_ptrs, self.mc_param_h = self._addresses(self.param_h, self._handles[2])
        # multimem through the NVSwitch pays off from 4 ranks on (ucsa_adam_exchange, 13.1 M parameters: 0.113 ms vs
        # 0.127 ms with peer loads at 4 ranks, 0.113 ms at 8); at 2 ranks plain peer loads / stores are faster
        # (0.078 ms vs 0.19 ms).  UCSA_PEER_MULTICAST=0/1 overrides.
        have_mc = bool(self.mc_grad and self.mc_param and self.mc_param_h)
        # fp32 masters of a slice stay on its owner (peers only need the fp16 working copy the kernels read): one third
        # of the store traffic of the exchange.  UCSA_PEER_BROADCAST_MASTERS=1 replicates the masters as well.
        self.broadcast_masters = os.environ.get("UCSA_PEER_BROADCAST_MASTERS", "0") == "1"
        want = os.environ.get("UCSA_PEER_MULTICAST", "auto")
        self.multicast = have_mc and (want == "1" or (want != "0" and self.world >= 4))
        self.begin, self.end = owner_slice(n_params, self.rank, self.world)

    def _addresses(self, tensor, handle):
        ptrs = [int(p) for p in handle.buffer_ptrs]
        delta = tensor.data_ptr() - ptrs[self.rank]
        if not 0 <= delta < handle.buffer_size:
            raise RuntimeError("symmetric tensor lies outside its rendezvoused buffer")
        mc = int(handle.multicast_ptr) if handle.multicast_ptr else 0
        return [p + delta for p in ptrs], (mc + delta if mc else 0)

    def gather_masters(self):
        """Make every rank's fp32 masters complete (collective; only needed when broadcast_masters is off, before a
        checkpoint / state_dict): each owner broadcasts its slice."""
        if self.broadcast_masters:
            return
        for r in range(self.world):
            lo, hi = owner_slice(self.n, r, self.world)
            if hi > lo:
                dist.broadcast(self.param[lo:hi], src=r)

    def barrier(self, channel: int):
        """Device-side barrier over all ranks on the current stream (CUDA-graph capturable)."""
        self._handles[0].barrier(channel=channel)


def peer_exchange_available() -> bool:
    """Symmetric memory needs an initialised NCCL group whose ranks share one NVLink / P2P domain."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return False
    if dist.get_backend() != "nccl":
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:  # noqa: BLE001
        return False
    return True
