#!/usr/bin/env python
"""Benchmark of the Semantic-NeRF hot path (BASELINE.json: train rays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): single-scene Semantic-NeRF training step, hash grid 16 levels x 2^19
entries, 4096-ray batch per GPU, 256 + 256 samples per ray (the reference's live path, renderer_semantics.py:127-128),
synthetic ScanNet-shaped scene with 640x480 views.  One step = render (perturb=True) + RGB/semantic/depth losses +
backward + Adam, exactly the work of training_step_nerf (joint_train_lightning_net.py:473-513).

`value`  : rays/s with every step's rays and ground truth already resident in HBM.
`e2e`    : the same step through the public API with HOST (pinned) inputs: H2D copy of the step's rays + ground
           truth and a D2H read of the loss inside the timed region.
`roofline`: dominant kernel (by device time inside the timed region, CUDA events on the launch stream) against
           the measured HBM peak; algorithmic bytes = 588 B/sample x samples per launch.
`cpu_baseline`: the CPU oracle port of the reference path (oracle/live_path.py) on a bounded sample, rank 0, N=1.
`--impl reference`: the same oracle port, K steps of a bounded sample each, on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
NUM_STEPS, UPSAMPLE_STEPS = 256, 256
N_CLASSES = 40
BOUND = 4.0
ENCODE_BYTES_PER_SAMPLE = 588  # 12 xyz + 512 gather/scatter + 64 features
CONFIG = {
    "workload": "semantic-nerf train step: hashgrid 16x2^19, 4096 rays/GPU, 256+256 samples/ray, "
                "synthetic ScanNet-shaped scene 640x480",
    "rays_per_gpu": RAYS_PER_GPU, "samples_per_ray": NUM_STEPS + UPSAMPLE_STEPS, "classes": N_CLASSES,
    "l2": "per-step working set ~0.9 GB >> 126 MB L2 (no explicit flush); the 25 MB fp16 table is L2-resident "
          "by design",
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take whatever was seen
            sm = [float(l.split(",")[1]) for _, l in self.rows[-3:] if l.count(",") >= 8] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU oracle arm
def oracle_step_rays_per_s(n_rays, steps, warmup, threads):
    """fwd + bwd of the reference path (oracle port) on `n_rays` rays x (256+256) samples; -> (rays/s, ms/step)."""
    import torch

    from oracle import live_path
    from ucsa_neural_rendering_b200.scene import SyntheticScene
    from ucsa_neural_rendering_b200.trainer import nerf_losses

    torch.set_num_threads(threads)
    heads = live_path.OracleHeads(bound=BOUND, num_semantic_classes=N_CLASSES, seed=1337, hash_amp=1e-4)
    opt = torch.optim.Adam([{"params": [heads.encoder]},
                            {"params": [heads.sigma_net, heads.color_net, heads.semantics_net], "weight_decay": 1e-6}],
                           lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scene = SyntheticScene(seed=0)
    g = torch.Generator().manual_seed(123)
    times = []
    for s in range(warmup + steps):
        pix = torch.randint(0, scene.W * scene.H, (n_rays,), generator=g)
        o, d, dn = scene.rays(s % scene.n_views, pix)
        rgb, depth, label = scene.ground_truth(o, d, dn)
        t0 = time.perf_counter()
        opt.zero_grad()
        out = live_path.run(heads, o[None], d[None], dn[None], num_steps=NUM_STEPS, upsample_steps=UPSAMPLE_STEPS,
                            perturb=True)
        loss, _ = nerf_losses(out, rgb[None], label[None], depth[None], scene.one_m_to_scene_uom)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return n_rays * len(times) / total, 1e3 * total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_rays = 64
    value, ms = oracle_step_rays_per_s(n_rays, args.steps, max(args.warmup, 1), cores)
    cfg = dict(CONFIG)
    cfg["reference_sample"] = f"{n_rays} rays x {NUM_STEPS + UPSAMPLE_STEPS} samples per step"
    line = {
        "impl": "reference", "metric": "semantic_nerf_train_rays_per_s", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 encode/MLP, fp32 composite",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/live_path.py fwd+bwd+Adam, {n_rays} rays x 512 samples per step, "
                                   f"{args.steps} steps"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ucsa_neural_rendering_b200 import _lib, ops
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork
    from ucsa_neural_rendering_b200.scene import SyntheticScene
    from ucsa_neural_rendering_b200.engine import TrainEngine
    from ucsa_neural_rendering_b200.trainer import nerf_losses

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rendering path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            # keep stdout to the one JSON line: NCCL writes its "NCCL version ..." banner to stdout
            os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the extension is missing

    scene = SyntheticScene(seed=0, device=dev)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=BOUND, cuda_ray=False, density_scale=1,
                              num_semantic_classes=N_CLASSES).to(dev).train()
    uom = scene.one_m_to_scene_uom
    engine = TrainEngine(net, RAYS_PER_GPU, num_steps=NUM_STEPS, upsample_steps=UPSAMPLE_STEPS,
                         one_m_to_scene_uom=uom, use_graph=not args.no_graph, exchange=args.exchange)

    total_steps = args.warmup + args.steps
    g = torch.Generator(device=dev).manual_seed(123 + rank)
    batches = []
    for s in range(total_steps):
        pix = torch.randint(0, scene.W * scene.H, (RAYS_PER_GPU,), device=dev, generator=g)
        o, d, dn = scene.rays(s % scene.n_views, pix)
        rgb, depth, label = scene.ground_truth(o, d, dn)
        batches.append(tuple(x[None].contiguous() for x in (o, d, dn, rgb.half(), label, depth)))
    host = [tuple(x.cpu().pin_memory() for x in b) for b in batches]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---------------------------------------------------------------- value: inputs resident in HBM
    # One step = one replay of the captured CUDA graph (engine.TrainEngine): render + losses + backward + Adam.
    _lib.stats.reset()
    engine.train_step(*batches[0])  # first call captures the graph: count the kernels of one step here
    torch.cuda.synchronize()
    for s in range(1, args.warmup):
        engine.train_step(*batches[s])
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for s in range(args.warmup, total_steps):
        engine.train_step(*batches[s])
    e1.record()
    barrier()
    t_wall1 = time.time()
    clock_info = clocks.stop(t_wall0, t_wall1)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * RAYS_PER_GPU * args.steps / (ms_total * 1e-3)

    # per-kernel device time: the same steps once more, launched eagerly with CUDA events around the kernels
    timed = {"ucsa_density_fwd", "ucsa_density_bwd", "ucsa_heads_fwd", "ucsa_heads_bwd", "ucsa_adam_step",
             "ucsa_adam_exchange", "ucsa_resample_merge"}
    eager = engine  # same engine, same buffers, launched kernel by kernel instead of replaying the graph
    was_graph, engine.use_graph = engine.use_graph, False
    eager.train_step(*batches[0])
    torch.cuda.synchronize()
    _lib.stats.reset()
    eager.train_step(*batches[1])
    torch.cuda.synchronize()
    per_step_launches = _lib.stats.launches  # kernels of ONE step; the graph replays exactly these
    by_name = dict(_lib.stats.by_name)
    launches = per_step_launches * args.steps
    _lib.stats.reset()
    _lib.stats.timed = set(timed)
    for s in range(args.warmup, total_steps):
        eager.train_step(*batches[s])
    torch.cuda.synchronize()
    kernel_ms = {k: _lib.stats.elapsed_ms(k) for k in timed}
    _lib.stats.timed = set()
    engine.use_graph = was_graph

    # ---------------------------------------------------------------- e2e: public API, host buffers
    opt = torch.optim.Adam([
        {"name": "encoding", "params": list(net.encoder.parameters())},
        {"name": "net", "params": list(net.sigma_net.parameters()) + list(net.color_net.parameters())
         + list(net.semantics_net.parameters()), "weight_decay": 1e-6}], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    for p_ in net.parameters():
        p_.grad = None

    def e2e_step(s):
        o, d, dn, rgb, label, depth = (x.to(dev, non_blocking=True) for x in host[s])
        opt.zero_grad(set_to_none=False)
        out = net.render(o, d, direction_norms=dn, staged=False, bg_color=None, perturb=True, seed=5000 + s,
                         ray_base=rank * RAYS_PER_GPU)
        loss, _ = nerf_losses(out, rgb, label, depth, uom, global_scale=1.0 / world)
        loss.backward()
        if world > 1:
            for p in net.parameters():
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
        opt.step()
        return float(loss.detach())  # D2H read of the step's result

    for s in range(args.warmup):
        e2e_step(s)
    barrier()
    e0.record()
    for s in range(args.warmup, total_steps):
        e2e_step(s)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * RAYS_PER_GPU * args.steps / (e2e_ms * 1e-3)

    # the engine fed from pinned host memory (H2D of the batch + D2H of the loss inside the timed region)
    barrier()
    e0.record()
    for s in range(args.warmup, total_steps):
        engine.load_batch(*host[s])
        float(engine.step()[0])
    e1.record()
    barrier()
    e2e_engine_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_engine_value = world * RAYS_PER_GPU * args.steps / (e2e_engine_ms * 1e-3)

    # ---------------------------------------------------------------- extra: full-frame rendering
    # one 640x480 view per rank (views shard, no collective), staged in chunks, no perturbation; reported beside the
    # headline, not part of it
    render_chunk = 65536
    view_pix = torch.arange(scene.W * scene.H, device=dev)
    net.eval()
    with torch.no_grad():
        vo, vd, vdn = scene.rays(rank % scene.n_views, view_pix)
        render_args = dict(direction_norms=vdn.view(1, -1, 1), staged=True, max_ray_batch=render_chunk, bg_color=None,
                           perturb=False, seed=99)
        for _ in range(2):
            out = net.render(vo[None], vd[None], **render_args)
        n_views = 3
        barrier()
        e0.record()
        for _ in range(n_views):
            out = net.render(vo[None], vd[None], **render_args)
            # the label map + u8 colours a caller writes out (pseudo-label epilogue, row f3)
            labels_u8, rgb_u8 = ops.label_epilogue(out["semantics"][0], out["image"][0])
        e1.record()
        barrier()
    net.train()
    render_ms = max_over_ranks(e0.elapsed_time(e1)) / n_views
    render_info = {"rays_per_s": world * scene.W * scene.H / (render_ms * 1e-3), "views_per_s": world / (render_ms * 1e-3),
                   "ms_per_view": render_ms, "rays_per_view": scene.W * scene.H, "chunk": render_chunk,
                   "samples_per_ray": NUM_STEPS + UPSAMPLE_STEPS, "api": "SemanticNeRFNetwork.render(staged=True)"}
    del labels_u8, rgb_u8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = measured_peaks()
    samples_per_launch = {
        # forward runs once per pass (coarse, fine), backward once over all 512 slots
        "ucsa_density_fwd": RAYS_PER_GPU * NUM_STEPS,
        "ucsa_density_bwd": RAYS_PER_GPU * (NUM_STEPS + UPSAMPLE_STEPS),
    }
    totals = {k: sum(v) for k, v in kernel_ms.items() if v}
    dominant = max(("ucsa_density_fwd", "ucsa_density_bwd"), key=lambda k: totals.get(k, 0.0))
    dur_ms = statistics.mean(kernel_ms[dominant])
    alg_bytes = ENCODE_BYTES_PER_SAMPLE * samples_per_launch[dominant]
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of the same kernel from the committed ncu capture (scripts/summarize_ncu.py traffic)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj["bytes_per_launch"].get(dominant.replace("ucsa_", "") + "_tc_kernel")
        traffic_src = tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    compulsory = 76 * samples_per_launch[dominant] + 13_074_912 * (2 if dominant == "ucsa_density_fwd" else 4)
    roofline = {
        "bound": "hbm", "kernel": dominant.replace("ucsa_", "") + "_tc_kernel", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dur_ms,
        "hbm_compulsory": {"bytes_per_launch": compulsory, "achieved": compulsory / (dur_ms * 1e-3) / 1e9,
                           "note": "76 B/sample (xyz, features, saved activations) + the table once (SURVEY 8d)"},
        # what actually binds the density kernels: the L2 request rate (scripts/gather_probe.cu, measured on B200:
        # 285 G 16-byte gathers/s, 192 G reductions/s whatever their width); operation counts per sample from
        # DESIGN.md section 5 (forward: 12 hashed levels x 4 corner pairs x 1.25; backward: ~68 after run-merging)
        "l2_request_roofline": {
            "ops_per_launch_est": (60 if dominant == "ucsa_density_fwd" else 68) * samples_per_launch[dominant],
            "peak_gops": 285.0 if dominant == "ucsa_density_fwd" else 192.0,
            "achieved_gops": (60 if dominant == "ucsa_density_fwd" else 68) * samples_per_launch[dominant]
            / (dur_ms * 1e-3) / 1e9,
            "frac": (60 if dominant == "ucsa_density_fwd" else 68) * samples_per_launch[dominant]
            / (dur_ms * 1e-3) / 1e9 / (285.0 if dominant == "ucsa_density_fwd" else 192.0),
            "note": "operation counts are analytic estimates; peaks measured by scripts/gather_probe.cu"},
        "note": "588 B/sample counts the 512 B of table gathers/scatters, which hit the L2-resident table rather "
                "than HBM (traffic = measured DRAM bytes of the launch); the kernel is bound by L2 reduction / L1 "
                "gather throughput, see DESIGN.md sections 5 and 9",
        "kernel_ms_per_step": {k.replace("ucsa_", ""): totals[k] / args.steps for k in totals},
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = 64
        v, ms = oracle_step_rays_per_s(n_cpu, 3, 1, cores)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": f"oracle/live_path.py fwd+bwd+Adam on {n_cpu} rays x 512 samples, 3 steps after 1 warm-up "
                         f"({ms:.0f} ms/step)"}

    line = {
        "metric": "semantic_nerf_train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 encode/MLP (fp32 accumulate), fp32 sampling/composite", "data": "synthetic",
        "config": dict(CONFIG, parallelism=f"ray-sharded dp{world}"),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps,
                "api": "SemanticNeRFNetwork.render + nerf_losses + torch.optim.Adam (the drop-in API a Lightning "
                       "loop calls), pinned host inputs",
                "engine": {"value": e2e_engine_value, "ms_per_step": e2e_engine_ms / args.steps,
                           "api": "TrainEngine.load_batch(pinned host) + step() + loss readback"}},
        "gpu_launches": launches, "gpu_launches_per_step": by_name,
        "step_impl": "cuda-graph replay of %d kernels" % per_step_launches if not args.no_graph else "eager kernel chain",
        "render": render_info,
        "gradient_exchange": {"none": "single GPU", "nccl": "NCCL all-reduce + replicated Adam",
                              "peer": "ucsa_adam_exchange over symmetric memory (%s)" % (
                                  "multimem.ld_reduce / multimem.st via NVSwitch" if engine.peer is not None
                                  and engine.peer.multicast else "peer loads / stores")}[engine.exchange],
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else a library writes there (NCCL's version
    banner, warnings) is sent to stderr by re-pointing file descriptor 1 for the rest of the run."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default=None,
                    help="multi-GPU gradient exchange (default: peer memory when available)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
