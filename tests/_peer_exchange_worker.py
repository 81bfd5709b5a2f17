"""Worker of tests/test_gpu_multi.py: launched by torchrun on >= 2 GPUs.  Trains a few steps with the fused
peer-memory exchange and with NCCL all-reduce + replicated Adam from identical starts, then asks the engine for its
single-step exchange check; rank 0 prints the results.  (Not a test module: the leading underscore keeps pytest from
collecting it.)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ucsa_neural_rendering_b200.engine import TrainEngine  # noqa: E402
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork  # noqa: E402


def make_batch(g, n, dev):
    o = (torch.rand(n, 3, generator=g) - 0.5) * 2
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    batch = (o, d, torch.ones(n), torch.rand(n, 3, generator=g).half(), torch.randint(0, 40, (n,), generator=g),
             torch.rand(n, generator=g) * 3)
    return [t.to(dev) for t in batch]


def run(exchange, multicast, masters, dev, rank, n, steps, inject_inf_at=None):
    os.environ["UCSA_PEER_MULTICAST"] = multicast
    os.environ["UCSA_PEER_BROADCAST_MASTERS"] = masters
    torch.manual_seed(0)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=40).to(dev).train()
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        net.encoder.params.copy_(((torch.rand(net.encoder.params.numel(), generator=g) * 2 - 1) * 0.3).to(dev))
    eng = TrainEngine(net, n, num_steps=128, upsample_steps=128, one_m_to_scene_uom=0.6, seed=11, exchange=exchange)
    g = torch.Generator().manual_seed(100 + rank)
    losses = []
    for it in range(steps):
        batch = make_batch(g, n, dev)
        if it == inject_inf_at and rank == 1:
            batch[3][0, 0] = float("inf")  # a poisoned ground-truth colour on ONE rank -> inf gradients there
        losses.append(float(eng.train_step(*batch)[0]))
    torch.cuda.synchronize()
    eng.gather_masters()
    params = torch.cat([m.params.detach().reshape(-1).clone() for m, _ in eng.groups])
    halves = torch.cat([m.half_params().reshape(-1).clone() for m, _ in eng.groups])
    return params, halves, losses, eng


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, steps = 512, 4
    ref_p, ref_h, ref_l, _ = run("nccl", "0", "0", dev, rank, n, steps)
    out = {}
    for tag, mc, masters in (("peer", "0", "0"), ("multicast", "1", "0"), ("peer+masters", "0", "1")):
        p, h, losses, eng = run("peer", mc, masters, dev, rank, n, steps)
        diff = (p - ref_p).abs()
        scale = float(ref_p.abs().max())
        check = eng.exchange_check()  # one more step, fused vs NCCL formulation from the same state
        out[tag] = dict(multicast=eng.peer.multicast, max_diff=float(diff.max()) / scale,
                        frac_within=float((diff <= 1e-3 * ref_p.abs() + 1e-6).float().mean()),
                        half_consistent=bool(torch.equal(h, p.half())),
                        loss_diff=max(abs(a - b) for a, b in zip(losses, ref_l)), check=check)
        del eng
    # overflow on one rank only: every rank must skip that step (GradScaler semantics across the job)
    for tag, exchange in (("inf_peer", "peer"), ("inf_nccl", "nccl")):
        p, h, losses, eng = run(exchange, "0", "0", dev, rank, n, 3, inject_inf_at=1)
        mine = h.clone()
        dist.broadcast(mine, src=0)
        same = torch.tensor([1.0 if torch.equal(mine, h) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out[tag] = dict(skipped=eng.skipped_steps, finite=bool(torch.isfinite(p).all()), identical=bool(same.item() == 1.0),
                        adam_steps=int(eng.step_dev) - eng.skipped_steps)
        del eng
    if rank == 0:
        print("PEER_EXCHANGE_RESULT", out, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
