"""Worker of tests/test_gpu_multi.py: launched by torchrun on >= 2 GPUs.  Trains a few steps with the fused
peer-memory exchange and with NCCL all-reduce + replicated Adam from identical starts; rank 0 prints the largest
parameter difference.  (Not a test module: the leading underscore keeps pytest from collecting it.)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ucsa_neural_rendering_b200.engine import TrainEngine  # noqa: E402
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork  # noqa: E402


def run(exchange, multicast, dev, rank, n, steps):
    os.environ["UCSA_PEER_MULTICAST"] = multicast
    torch.manual_seed(0)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=40).to(dev).train()
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        net.encoder.params.copy_(((torch.rand(net.encoder.params.numel(), generator=g) * 2 - 1) * 0.3).to(dev))
    eng = TrainEngine(net, n, num_steps=128, upsample_steps=128, one_m_to_scene_uom=0.6, seed=11, exchange=exchange)
    g = torch.Generator().manual_seed(100 + rank)
    losses = []
    for _ in range(steps):
        o = (torch.rand(n, 3, generator=g) - 0.5) * 2
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        batch = (o, d, torch.ones(n), torch.rand(n, 3, generator=g).half(), torch.randint(0, 40, (n,), generator=g),
                 torch.rand(n, generator=g) * 3)
        losses.append(float(eng.train_step(*[t.to(dev) for t in batch])[0]))
    torch.cuda.synchronize()
    params = torch.cat([m.params.detach().reshape(-1).clone() for m, _ in eng.groups])
    halves = torch.cat([m.half_params().reshape(-1).clone() for m, _ in eng.groups])
    return params, halves, losses, eng


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, steps = 512, 4
    ref_p, ref_h, ref_l, _ = run("nccl", "0", dev, rank, n, steps)
    out = {}
    for mc in ("0", "1"):
        p, h, losses, eng = run("peer", mc, dev, rank, n, steps)
        # replicas must be bit-identical across ranks: every copy of a parameter is written by its one owner
        mine = p.clone()
        dist.broadcast(mine, src=0)
        same = bool(torch.equal(mine, p))
        flags = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        scale = float(ref_p.abs().max())
        out[mc] = dict(multicast=eng.peer.multicast, identical=bool(flags.item() == 1.0),
                       max_diff=float((p - ref_p).abs().max()) / scale,
                       half_consistent=bool(torch.equal(h, p.half())),
                       loss_diff=max(abs(a - b) for a, b in zip(losses, ref_l)))
    if rank == 0:
        print("PEER_EXCHANGE_RESULT", out, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
