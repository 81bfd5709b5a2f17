"""Bring-up probe: torch symmetric memory (peer pointers, multicast, barrier inside a CUDA graph) on this box.
torchrun --nproc-per-node 2 scripts/probe_symm.py"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "mc", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None,
          "sigpad", hdl.signal_pad_size, flush=True)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", float(peer[0]), flush=True)
    hdl.barrier(channel=0)
    # barrier inside a CUDA graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        hdl.barrier(channel=1)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            t.add_(1.0)
            hdl.barrier(channel=1)
            peer.add_(0.0)
            hdl.barrier(channel=1)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        print(rank, "graph barrier ok, value", float(t[0]), flush=True)
    except Exception as e:  # noqa: BLE001
        print(rank, "graph barrier FAILED", repr(e), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
