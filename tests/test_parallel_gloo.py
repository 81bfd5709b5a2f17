"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding, the single flat-gradient all-reduce with
1/world loss scaling, occupancy-grid max-sync, view sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ucsa_neural_rendering_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rays, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a toy "model": gradient of a mean-squared loss over this rank's ray slice w.r.t. a replicated parameter
        g = torch.Generator().manual_seed(0)
        rays = torch.randn(n_rays, 8, generator=g)
        target = torch.randn(n_rays, generator=g)
        w_param = torch.zeros(8, requires_grad=True)
        lo, hi = parallel.shard_range(n_rays, rank, world)
        local = ((rays[lo:hi] @ w_param - target[lo:hi]) ** 2).sum() / n_rays  # global-mean normalisation
        local.backward()
        flat = w_param.grad.clone()
        parallel.all_reduce_gradients(flat)
        grid = torch.zeros(2, 4, 4, 4)
        grid.view(-1)[rank::world] = rank + 1.0
        mean_density = parallel.sync_density_grid(grid)
        slowest = parallel.max_over_ranks(1.0 + rank)
        torch.save({"flat": flat, "grid": grid, "mean": mean_density, "range": (lo, hi), "slowest": slowest,
                    "views": parallel.views_of_rank(11, rank, world)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_ray_sharded_gradients_match_single_process(tmp_path):
    world, n_rays = 2, 101  # odd on purpose: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(n_rays, 8, generator=g)
    target = torch.randn(n_rays, generator=g)
    w_param = torch.zeros(8, requires_grad=True)
    (((rays @ w_param - target) ** 2).mean()).backward()
    for r in res:
        torch.testing.assert_close(r["flat"], w_param.grad, rtol=1e-6, atol=1e-7)
    assert torch.equal(res[0]["flat"], res[1]["flat"]), "all ranks must apply the identical update"
    assert res[0]["range"] == (0, 51) and res[1]["range"] == (51, 101)
    assert torch.equal(res[0]["grid"], res[1]["grid"]) and float(res[0]["grid"].min()) >= 1.0
    assert res[0]["mean"] == res[1]["mean"] and res[0]["slowest"] == res[1]["slowest"] == 2.0
    assert sorted(res[0]["views"] + res[1]["views"]) == list(range(11))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 65536 + 3):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_owner_slices_cover_the_flat_parameter_space_once():
    """host logic of the fused exchange: equal float4-aligned slices, every parameter owned by exactly one rank"""
    for n in (13_089_248, 4096, 40, 4, 1004):
        for world in (1, 2, 3, 4, 8, 16):
            covered = 0
            for rank in range(world):
                lo, hi = parallel.owner_slice(n, rank, world)
                assert lo == covered and lo % 4 == 0 and lo <= hi <= n
                assert hi % 4 == 0 or hi == n
                covered = hi
            assert covered == n
