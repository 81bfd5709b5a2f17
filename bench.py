#!/usr/bin/env python
"""Benchmark of the Semantic-NeRF hot path (BASELINE.json: train rays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): single-scene Semantic-NeRF training step, hash grid 16 levels x 2^19
entries, 4096-ray batch per GPU, 256 + 256 samples per ray (the reference's live path, renderer_semantics.py:127-128),
synthetic ScanNet-shaped scene with 640x480 views.  One step = render (perturb=True) + RGB/semantic/depth losses +
backward + Adam, exactly the work of training_step_nerf (joint_train_lightning_net.py:473-513).

`value`  : rays/s with every step's rays and ground truth already resident in HBM.
`e2e`    : the same step through the public drop-in API (SemanticNeRFNetwork.render + nerf_losses + backward +
           FusedAdam.step in the caller's own loop) with HOST (pinned) inputs: H2D copy of the step's rays + ground
           truth and a D2H read of the loss inside the timed region; `e2e.torch_adam` keeps the caller's optimizer.
`roofline`: dominant kernel (by device time inside the timed region, CUDA events on the launch stream) against
           the measured HBM peak; algorithmic bytes = 588 B/sample x samples per launch.
`cpu_baseline`: the CPU oracle port of the reference path (oracle/live_path.py), one full 4096-ray step, rank 0, N=1.
`--impl reference`: the same oracle port on all host cores, K timed steps of the SAME 4096-ray batch size (processed
           in 1024-ray slices with gradient accumulation to bound memory; warm-up steps use 256 rays).
`config1` : BASELINE.json configs[0], the 4096 x 128 x 40 dense composite (forward and backward) against the HBM
           roofline, with the reference's run() on the host cores beside it.
Extra keys: `render` (config 3, one 640x480 view per rank), `trained` (the headline step after 300 optimisation steps,
           when the w > 1e-4 mask really prunes), `strong_scaling_2p16` (config 4's 2^16-ray global batch split over
           the ranks), `exchange_check` (N > 1: fused peer exchange vs the NCCL formulation, replicas identical).
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
DTYPE = "fp16 encode/MLP (fp32 accumulate), fp32 sampling/composite"
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
        note = None
        if not sm:  # region shorter than the sampling period: take the samples nearest to it
            near = []
            for ts, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    near.append((abs(ts - 0.5 * (t0 + t1)), float(parts[1])))
                    smax = float(parts[2])
                except (ValueError, IndexError):
                    continue
            near.sort()
            sm = [c for _, c in near[:3]]
            note = "no sample inside the timed region; nearest samples used"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax, "reasons": ["no nvidia-smi sample parsed"], "samples": 0}
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------ CPU oracle arm
ORACLE_SLICE = 1024  # rays per forward/backward slice of the CPU arm (bounds memory; gradients accumulate)
ORACLE_WARMUP_RAYS = 256


def oracle_step_rays_per_s(n_rays, steps, warmup, threads):
    """training_step_nerf of the reference on the CPU oracle port: render (256+256 samples) + losses + backward + Adam
    on `n_rays` rays per step, processed in ORACLE_SLICE-ray slices whose gradients accumulate (same arithmetic as one
    pass: the losses are means over the whole batch).  Warm-up steps use ORACLE_WARMUP_RAYS rays (untimed).
    -> (rays/s, ms/step)."""
    import torch

    from oracle import live_path
    from oracle.losses import nerf_losses
    from ucsa_neural_rendering_b200.scene import SyntheticScene

    torch.set_num_threads(threads)
    heads = live_path.OracleHeads(bound=BOUND, num_semantic_classes=N_CLASSES, seed=1337, hash_amp=1e-4)
    opt = torch.optim.Adam([{"params": [heads.encoder]},
                            {"params": [heads.sigma_net, heads.color_net, heads.semantics_net], "weight_decay": 1e-6}],
                           lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scene = SyntheticScene(seed=0)
    g = torch.Generator().manual_seed(123)
    times, fixed = [], []
    for s in range(warmup + steps):
        n = n_rays if s >= warmup else min(n_rays, ORACLE_WARMUP_RAYS)
        pix = torch.randint(0, scene.W * scene.H, (n,), generator=g)
        o, d, dn = scene.rays(s % scene.n_views, pix)
        rgb, depth, label = scene.ground_truth(o, d, dn)
        t0 = time.perf_counter()
        opt.zero_grad()
        t_fix = time.perf_counter() - t0
        for lo in range(0, n, ORACLE_SLICE):
            hi = min(lo + ORACLE_SLICE, n)
            out = live_path.run(heads, o[None, lo:hi], d[None, lo:hi], dn[None, lo:hi], num_steps=NUM_STEPS,
                                upsample_steps=UPSAMPLE_STEPS, perturb=True)
            loss, _ = nerf_losses(out, rgb[None, lo:hi], label[None, lo:hi], depth[None, lo:hi],
                                  scene.one_m_to_scene_uom, global_scale=(hi - lo) / n)
            loss.backward()
        t1 = time.perf_counter()
        opt.step()
        t_fix += time.perf_counter() - t1
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
            fixed.append(t_fix)
    total = sum(times)
    # share of the step that does not depend on the ray count: zero-filling and Adam over the 13.1 M parameters
    # (the dense 52 MB gradient accumulation inside backward is per slice and stays in the variable part)
    oracle_step_rays_per_s.fixed_share = sum(fixed) / total
    return n_rays * len(times) / total, 1e3 * total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_rays = RAYS_PER_GPU  # the same batch as the CUDA arm
    value, ms = oracle_step_rays_per_s(n_rays, args.steps, args.warmup, cores)
    sample = (f"oracle/live_path.py + oracle/losses.py fwd+bwd + torch Adam, {n_rays} rays x "
              f"{NUM_STEPS + UPSAMPLE_STEPS} samples per timed step (slices of {ORACLE_SLICE} rays, gradients "
              f"accumulated), {args.steps} steps; {args.warmup} warm-up steps of {ORACLE_WARMUP_RAYS} rays; "
              f"ray-count-independent share of a step (zero_grad + Adam over 13.1 M parameters): "
              f"{100 * oracle_step_rays_per_s.fixed_share:.1f} %")
    line = {
        "impl": "reference", "metric": "semantic_nerf_train_rays_per_s", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": dict(CONFIG, parallelism=f"ray-sharded dp{args.gpus}"),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ CUDA arm
def time_config1(dev, peak):
    """BASELINE.json configs[0] on the GPU: dense composite of 4096 rays x 128 samples x 40 classes, forward and
    backward (csrc/composite_dense.cu), CUDA events per launch.  The forward working set (95 MB) would fit the 126 MB
    L2, so four independent input sets are rotated: a launch's inputs were last touched ~300 MB of traffic ago."""
    import torch

    from ucsa_neural_rendering_b200 import ops

    n, t, c, sets, reps = 4096, 128, 40, 4, 5
    g = torch.Generator(device=dev).manual_seed(1234)
    f32 = dict(dtype=torch.float32, device=dev)
    data = []
    for _ in range(sets):
        sigma = 50.0 * torch.rand(n, t, generator=g, **f32) ** 4
        rgb = torch.rand(n, t, 3, generator=g, **f32)
        prob = torch.softmax(torch.randn(n, t, c, generator=g, **f32), dim=-1)
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, **f32), dim=-1)
        nears, fars = ops.near_far_from_aabb(torch.zeros(n, 3, **f32), d, torch.tensor([-BOUND] * 3 + [BOUND] * 3, **f32))
        z = nears[:, None] + (fars - nears)[:, None] * torch.linspace(0, 1, t, **f32)[None]
        data.append(dict(sigma=sigma, rgb=rgb, prob=prob, z=z.contiguous(), dn=torch.ones(n, **f32),
                         w=torch.empty(n, t, **f32), depth=torch.empty(n, **f32), image=torch.empty(n, 3, **f32),
                         sem=torch.empty(n, c, **f32), gd=torch.randn(n, generator=g, **f32),
                         gi=torch.randn(n, 3, generator=g, **f32), gs=torch.randn(n, c, generator=g, **f32),
                         d_sigma=torch.empty(n, t, **f32), d_rgb=torch.empty(n, t, 3, **f32),
                         d_prob=torch.empty(n, t, c, **f32)))

    def fwd(x):
        ops.composite_dense_fwd(x["sigma"], x["z"], x["rgb"], x["prob"], x["dn"], 1.0, x["w"], x["depth"], x["image"],
                                x["sem"])

    def bwd(x):
        ops.composite_dense_bwd(x["sigma"], x["z"], x["rgb"], x["w"], x["dn"], x["gd"], x["gi"], x["gs"], 1.0,
                                x["d_sigma"], x["d_rgb"], x["d_prob"])

    for x in data:  # warm-up (also fills the weights the backward reads)
        fwd(x)
        bwd(x)
    torch.cuda.synchronize()
    # back-to-back launches between two events on the launch stream (per-launch event pairs would add ~3 us of launch
    # latency to a 20 us kernel); consecutive launches work on different input sets
    def timed(fn):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            for x in data:
                fn(x)
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / (reps * sets)

    ms_f = min(timed(fwd) for _ in range(3))
    ms_b = min(timed(bwd) for _ in range(3))
    # algorithmic bytes, SURVEY.md section 8(d): forward 180 B/sample + 180 B/ray; backward reads 180 + writes 176 per
    # sample and reads 352 per ray
    bytes_f = n * t * 180 + n * 180
    bytes_b = n * t * (180 + 176) + n * 352
    # what this backward really has to move: d_prob = w * g_semantics does not read p (semantic weights are detached,
    # renderer_semantics.py:270), so per sample it reads sigma, z, rgb, w (24 B) and writes 176 B
    bytes_b_needed = n * t * (24 + 176) + n * 184
    return {
        "workload": "dense composite 4096 rays x 128 samples x 40 classes (BASELINE.json configs[0])",
        "l2": "4 rotating input sets of 95 MB (forward) / 188 MB (backward) each: no launch finds its inputs in L2",
        "fwd": {"ms": ms_f, "algorithmic_bytes": bytes_f, "achieved_gbs": bytes_f / ms_f / 1e6,
                "frac_of_hbm_peak": bytes_f / ms_f / 1e6 / peak},
        "bwd": {"ms": ms_b, "algorithmic_bytes": bytes_b, "achieved_gbs": bytes_b / ms_b / 1e6,
                "frac_of_hbm_peak": bytes_b / ms_b / 1e6 / peak, "bytes_the_kernel_needs": bytes_b_needed,
                "frac_of_hbm_peak_needed_bytes": bytes_b_needed / ms_b / 1e6 / peak,
                "note": "SURVEY 8(d) counts a 160 B/sample read of p that the backward does not need; the second figure "
                        "uses the bytes the kernel must move (its 92 MB of writes may still sit in L2 when it ends)"},
        "rays_per_s_fwd": n / (ms_f * 1e-3), "rays_per_s_fwd_bwd": n / ((ms_f + ms_b) * 1e-3), "peak_gbs": peak,
        "launches": 2 * 3 * sets * reps,
        "timing": "CUDA events around %d back-to-back launches, best of 3 passes" % (sets * reps),
    }


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
    clocks = ClockSampler(local_rank)  # started early: nvidia-smi needs a moment before its first row
    clocks.start()
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

    def make_batches(count, n_rays):
        out = []
        for s in range(count):
            pix = torch.randint(0, scene.W * scene.H, (n_rays,), device=dev, generator=g)
            o, d, dn = scene.rays(s % scene.n_views, pix)
            rgb, depth, label = scene.ground_truth(o, d, dn)
            out.append(tuple(x[None].contiguous() for x in (o, d, dn, rgb.half(), label, depth)))
        return out

    batches = make_batches(total_steps, RAYS_PER_GPU)
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

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_engine_steps(eng, data, first, last):
        """device time (max over ranks) of eng.train_step over data[first:last], barrier + synchronize on both sides"""
        barrier()
        e0.record()
        for s in range(first, last):
            eng.train_step(*data[s % len(data)])
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------------------------------------------------------- value: inputs resident in HBM
    # One step = one replay of the captured CUDA graph (engine.TrainEngine): render + losses + backward + Adam.
    _lib.stats.reset()
    engine.train_step(*batches[0])  # first call captures the graph
    torch.cuda.synchronize()
    for s in range(1, args.warmup):
        engine.train_step(*batches[s])
    t_wall0 = time.time()
    ms_total = time_engine_steps(engine, batches, args.warmup, total_steps)
    t_wall1 = time.time()
    value = world * RAYS_PER_GPU * args.steps / (ms_total * 1e-3)

    # per-kernel device time: the same steps once more, launched eagerly with CUDA events around the kernels
    timed = {"ucsa_density_fwd", "ucsa_density_bwd", "ucsa_heads_fwd", "ucsa_heads_bwd", "ucsa_adam_step",
             "ucsa_adam_exchange", "ucsa_resample_merge", "ucsa_grad_check", "ucsa_weights_compact", "ucsa_weights_bwd"}
    was_graph, engine.use_graph = engine.use_graph, False
    engine.train_step(*batches[0])
    torch.cuda.synchronize()
    _lib.stats.reset()
    engine.train_step(*batches[1])
    torch.cuda.synchronize()
    per_step_launches = _lib.stats.launches  # kernels of ONE step; the graph replays exactly these
    by_name = dict(_lib.stats.by_name)
    launches = per_step_launches * args.steps
    _lib.stats.reset()
    _lib.stats.timed = set(timed)
    for s in range(args.warmup, total_steps):
        engine.train_step(*batches[s])
    torch.cuda.synchronize()
    kernel_ms = {k: _lib.stats.elapsed_ms(k) for k in timed}
    _lib.stats.timed = set()
    engine.use_graph = was_graph
    k_over_s_init = float(engine.ws.ray_off[-1]) / (RAYS_PER_GPU * (NUM_STEPS + UPSAMPLE_STEPS))

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

    # ---------------------------------------------------------------- multi-GPU: evidence for the fused exchange
    exchange_check = engine.exchange_check() if world > 1 else None
    exchange_mode = engine.exchange
    multicast = bool(engine.peer is not None and engine.peer.multicast)

    # ---------------------------------------------------------------- extra: the same step on a trained model
    # at random initialisation nearly every sample passes w > 1e-4 (K/S ~ 0.93: the heads see the worst case); after a
    # few hundred steps the mask prunes.  300 optimisation steps over the batches above, then K timed steps.
    trained = None
    if not args.no_extras:
        n_train = args.train_steps
        gt = torch.Generator(device=dev).manual_seed(777 + rank)
        for s in range(n_train):  # fresh pixels of a fresh view every step, generated on the device
            pix = torch.randint(0, scene.W * scene.H, (RAYS_PER_GPU,), device=dev, generator=gt)
            o, d, dn = scene.rays(int(torch.randint(0, scene.n_views, (1,), generator=gt, device=dev)), pix)
            rgb, depth, label = scene.ground_truth(o, d, dn)
            engine.train_step(o, d, dn, rgb.half(), label, depth)
        ms_tr = time_engine_steps(engine, batches, args.warmup, total_steps)
        k_over_s = float(engine.ws.ray_off[-1]) / (RAYS_PER_GPU * (NUM_STEPS + UPSAMPLE_STEPS))
        loss4 = [float(v) for v in engine.loss]
        trained = {"after_steps": n_train + 2 * total_steps + 2, "ms_per_step": ms_tr / args.steps,
                   "rays_per_s": world * RAYS_PER_GPU * args.steps / (ms_tr * 1e-3), "k_over_s": k_over_s,
                   "k_over_s_at_init": k_over_s_init, "loss": loss4,
                   "skipped_steps": engine.skipped_steps}
    engine.gather_masters()

    # ---------------------------------------------------------------- e2e: public API, host buffers
    def param_groups():
        return [{"name": "encoding", "params": list(net.encoder.parameters())},
                {"name": "net", "params": list(net.sigma_net.parameters()) + list(net.color_net.parameters())
                 + list(net.semantics_net.parameters()), "weight_decay": 1e-6}]

    # The drop-in configuration INTEGRATION.md describes: the network module and the optimizer class come from this
    # package (two changed lines in the caller), everything else is the caller's loop: render(), the three losses,
    # backward(), optimizer.step().  FusedAdam = torch.optim.Adam's numbers behind torch.optim's interface.
    from ucsa_neural_rendering_b200.optim import FusedAdam

    opt = FusedAdam(param_groups(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15, network=net)
    for p_ in net.parameters():
        p_.grad = None

    loss_host = torch.zeros(total_steps, dtype=torch.float32).pin_memory()

    def e2e_step(s, sync):
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
        if sync:
            return float(loss.detach())  # blocking D2H read of the step's result
        loss_host[s].copy_(loss.detach(), non_blocking=True)  # D2H read of the step's result, pinned destination
        return None

    def time_e2e(sync):
        for s in range(args.warmup):
            e2e_step(s, sync)
        barrier()
        e0.record()
        for s in range(args.warmup, total_steps):
            e2e_step(s, sync)
        e1.record()
        barrier()  # (synchronizes: every loss has landed in host memory before the clock is read)
        return max_over_ranks(e0.elapsed_time(e1))

    # the step's loss goes to pinned host memory with an asynchronous copy each step (what a Lightning loop does: the
    # logged loss is only read on the host at the logging interval); "blocking" reads it with .item() every step
    e2e_ms = time_e2e(sync=False)
    e2e_value = world * RAYS_PER_GPU * args.steps / (e2e_ms * 1e-3)
    assert bool(torch.isfinite(loss_host[args.warmup:]).all()) and float(loss_host[args.warmup:].abs().sum()) > 0
    e2e_blocking_ms = time_e2e(sync=True)
    # the same loop with the caller's optimizer left untouched (torch.optim.Adam, ~10 foreach passes over the table)
    for p_ in net.parameters():
        p_.grad = None
    opt = torch.optim.Adam(param_groups(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    e2e_torch_adam_ms = time_e2e(sync=False)
    opt = None
    for p_ in net.parameters():
        p_.grad = None

    # ---------------------------------------------------------------- extra: full-frame rendering
    # one 640x480 view per rank (views shard, no collective), staged, no perturbation, with the caller's DEFAULT
    # max_ray_batch = 4096 (joint_train_lightning_net.py:237 / renderer_semantics.py:306): the network re-chunks to
    # 65536 rays internally; "strict_4096" forces the reference's chunking.
    render_info = None
    if not args.no_extras:
        view_pix = torch.arange(scene.W * scene.H, device=dev)
        net.eval()
        with torch.no_grad():
            vo, vd, vdn = scene.rays(rank % scene.n_views, view_pix)
            render_args = dict(direction_norms=vdn.view(1, -1, 1), staged=True, bg_color=None, perturb=False, seed=99)

            def time_render(n_views):
                for _ in range(2):
                    net.render(vo[None], vd[None], **render_args)
                barrier()
                e0.record()
                for _ in range(n_views):
                    out = net.render(vo[None], vd[None], **render_args)
                    # the label map + u8 colours a caller writes out (pseudo-label epilogue, row f3)
                    ops.label_epilogue(out["semantics"][0], out["image"][0])
                e1.record()
                barrier()
                return max_over_ranks(e0.elapsed_time(e1)) / n_views

            render_ms = time_render(3)
            chunk_default = net.stage_chunk
            net.stage_chunk = None
            render_ms_strict = time_render(2)
            net.stage_chunk = chunk_default
        net.train()
        rays_view = scene.W * scene.H
        render_info = {"rays_per_s": world * rays_view / (render_ms * 1e-3), "views_per_s": world / (render_ms * 1e-3),
                       "ms_per_view": render_ms, "rays_per_view": rays_view, "max_ray_batch": 4096,
                       "internal_chunk": chunk_default, "samples_per_ray": NUM_STEPS + UPSAMPLE_STEPS,
                       "api": "SemanticNeRFNetwork.render(staged=True) with the default max_ray_batch",
                       "strict_4096": {"rays_per_s": world * rays_view / (render_ms_strict * 1e-3),
                                       "ms_per_view": render_ms_strict}}

    # ---------------------------------------------------------------- extra: GPU stand-in for the tcnn path
    # tiny-cuda-nn cannot be installed here, so BASELINE.json's "10x the reference GPU path" has no measurable
    # denominator.  Stand-in, clearly labelled: the same step with the network heads in eager PyTorch (index gathers +
    # fp16 cuBLAS matmuls + autograd, baseline/eager_torch_heads.py) through the reference-shaped run(); rank 0 only.
    stand_in = None
    if not args.no_extras and rank == 0:
        try:
            from baseline.eager_torch_heads import EagerTorchNetwork

            ref_net = EagerTorchNetwork(bound=BOUND, num_semantic_classes=N_CLASSES).to(dev).train()
            ref_opt = torch.optim.Adam([{"params": [ref_net.table]},
                                        {"params": [p for n_, p in ref_net.named_parameters() if n_ != "table"],
                                         "weight_decay": 1e-6}], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
            scaler = torch.amp.GradScaler("cuda")

            def ref_step(s):
                o, d, dn, rgb, label, depth = batches[s % len(batches)]
                ref_opt.zero_grad(set_to_none=True)
                out = ref_net.render(o, d, direction_norms=dn, staged=False, bg_color=None, perturb=True, seed=7000 + s)
                loss, _ = nerf_losses(out, rgb, label, depth, uom)
                scaler.scale(loss).backward()
                scaler.step(ref_opt)
                scaler.update()

            for s in range(2):
                ref_step(s)
            torch.cuda.synchronize()
            k_ref = 5
            e0.record()
            for s in range(k_ref):
                ref_step(2 + s)
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / k_ref
            stand_in = {"what": "STAND-IN, not tiny-cuda-nn: reference-shaped run() with eager-PyTorch heads (hash-grid "
                                "as index gathers, fp16 cuBLAS MLPs, autograd) + GradScaler + torch Adam on this GPU",
                        "ms_per_step": ms_ref, "rays_per_s": RAYS_PER_GPU / (ms_ref * 1e-3), "steps": k_ref,
                        "rays_per_step": RAYS_PER_GPU}
            del ref_net, ref_opt
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001 - an extra must never take the headline down
            stand_in = {"error": repr(exc)[:300]}

    # ---------------------------------------------------------------- extra: the occupancy-grid path (rows a16-a20)
    # dormant in the reference (cuda_ray=False is hard-coded by its only caller); measured here so that it has
    # numbers: grid refresh (3 launches + 1 host read), one training render (march + heads + ragged composite,
    # forward + backward) of 4096 rays, and the inference wavefront over one 640x480 view.  Parameters of the trained
    # network above; rank 0 only.
    occupancy = None
    if not args.no_extras and rank == 0:
        try:
            occ_net = SemanticNeRFNetwork(encoding="hashgrid", bound=BOUND, cuda_ray=True, density_scale=1,
                                          num_semantic_classes=N_CLASSES).to(dev)
            with torch.no_grad():
                for dst, src in ((occ_net.encoder, net.encoder), (occ_net.sigma_net, net.sigma_net),
                                 (occ_net.color_net, net.color_net), (occ_net.semantics_net, net.semantics_net)):
                    dst.params.copy_(src.params)
            occ_net.train()

            def timed_ms(fn, reps):
                fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            ms_refresh = timed_ms(occ_net.update_extra_state, 5)
            o, d, dn = batches[0][:3]

            def occ_train():
                for p_ in occ_net.parameters():
                    p_.grad = None
                out = occ_net.render(o, d, direction_norms=dn, staged=False, perturb=True, dt_gamma=1 / 128,
                                     force_all_rays=True)
                (out["image"].sum() + out["depth"].sum() + out["semantics"].sum()).backward()

            ms_train = timed_ms(occ_train, 5)
            samples = int(occ_net.step_counter[(occ_net.local_step - 1) % 16, 0])
            occ_net.fused_packed = False  # the module-by-module heads under autograd, for comparison
            ms_train_modules = timed_ms(occ_train, 5)
            occ_net.fused_packed = True
            occ_net.eval()
            with torch.no_grad():
                vo, vd, vdn = scene.rays(0, torch.arange(scene.W * scene.H, device=dev))
                ms_view = timed_ms(lambda: occ_net.render(vo[None], vd[None], direction_norms=vdn.view(1, -1, 1),
                                                          staged=True, perturb=False, dt_gamma=1 / 128), 2)
            occupancy = {"grid_refresh_ms": ms_refresh, "cells": int(occ_net.density_grid.numel()),
                         "mean_density": float(occ_net.mean_density),
                         "occupied_fraction": float((occ_net.density_grid > min(0.01, occ_net.mean_density)).float().mean()),
                         "train_render_fwd_bwd_ms": ms_train, "train_rays": RAYS_PER_GPU, "train_samples": samples,
                         "train_rays_per_s": RAYS_PER_GPU / (ms_train * 1e-3),
                         "train_render_fwd_bwd_ms_module_level_heads": ms_train_modules,
                         "infer_view_ms": ms_view, "infer_rays_per_s": scene.W * scene.H / (ms_view * 1e-3),
                         "note": "training render = march + fused heads node (density kernel + tensor-core heads "
                                 "kernels each way) + ragged composite, forward + backward, no optimizer; "
                                 "module_level_heads = the same with the nine module-level launches under autograd"}
            del occ_net
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001 - an extra must never take the headline down
            occupancy = {"error": repr(exc)[:300]}

    # ---------------------------------------------------------------- extra: config 4's 2^16-ray global batch, strong scaling
    strong, sweep = None, None
    if not args.no_extras:
        del engine
        torch.cuda.empty_cache()

        def strong_scaling(global_rays):
            """one engine with global_rays / world rays per GPU: a few timed steps"""
            n_strong = global_rays // world
            eng2 = TrainEngine(net, n_strong, num_steps=NUM_STEPS, upsample_steps=UPSAMPLE_STEPS,
                               one_m_to_scene_uom=uom, use_graph=not args.no_graph, exchange=args.exchange)
            b2 = make_batches(3, n_strong)
            for s in range(3):
                eng2.train_step(*b2[s])
            k2 = max(3, args.steps // 4)
            ms2 = time_engine_steps(eng2, b2, 0, k2)
            eng2.gather_masters()
            del eng2
            torch.cuda.empty_cache()
            return {"global_batch": n_strong * world, "rays_per_gpu": n_strong, "steps": k2, "ms_per_step": ms2 / k2,
                    "rays_per_s": n_strong * world * k2 / (ms2 * 1e-3), "scaling": "strong"}

        try:
            strong = strong_scaling(65536)  # BASELINE.json configs[3]: ray batch 2^16 sharded over the GPUs
        except Exception as exc:  # noqa: BLE001 - an extra must never take the headline down
            strong = {"error": repr(exc)[:300]}
        if world >= 2:  # BASELINE.json configs[4]: 2^18-ray batches, ray-sharded at 2 / 4 / 8 GPUs
            try:
                sweep = strong_scaling(262144)
            except Exception as exc:  # noqa: BLE001
                sweep = {"error": repr(exc)[:300]}

    clock_info = clocks.stop(t_wall0, t_wall1)
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
    kname = dominant.replace("ucsa_", "") + "_tc_kernel"
    traffic, traffic_src, counters = None, None, None
    try:  # per-launch ncu counters of the same kernel from the committed capture (scripts/summarize_ncu.py traffic)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj["bytes_per_launch"].get(kname)
        traffic_src = tj["source"]
        counters = tj.get("l2_requests_per_launch", {}).get(kname)
    except (OSError, ValueError, KeyError):
        pass
    compulsory = 76 * samples_per_launch[dominant] + 13_074_912 * (2 if dominant == "ucsa_density_fwd" else 4)
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dur_ms,
        "hbm_compulsory": {"bytes_per_launch": compulsory, "achieved": compulsory / (dur_ms * 1e-3) / 1e9,
                           "note": "76 B/sample (xyz, features, saved activations) + the table once (SURVEY 8d)"},
        "note": "588 B/sample counts the 512 B of table gathers/scatters, which hit the L2-resident table rather "
                "than HBM (traffic = measured DRAM bytes of the launch); the kernel is bound by L2 reduction / L1 "
                "gather throughput, see DESIGN.md sections 5 and 9",
        "kernel_ms_per_step": {k.replace("ucsa_", ""): totals[k] / args.steps for k in totals},
    }
    if counters:
        # what actually binds the density kernels: the L2 request rate (scripts/gather_probe.cu, measured on B200:
        # 285 G 16-byte gathers/s, 192 G reductions/s whatever their width).  Request counts are ncu counters of the
        # committed capture (lts__t_requests_srcunit_tex_op_red / _op_read), not estimates.
        op = "red" if dominant == "ucsa_density_bwd" else "read"
        peak_g = 192.0 if op == "red" else 285.0
        reqs = counters.get("lts_requests_" + op)
        if reqs:
            roofline["l2_request_roofline"] = {
                "op": op, "requests_per_launch": reqs, "peak_greq_s": peak_g,
                "achieved_greq_s": reqs / (dur_ms * 1e-3) / 1e9, "frac": reqs / (dur_ms * 1e-3) / 1e9 / peak_g,
                "source": traffic_src, "counters": counters}

    config1 = time_config1(dev, peak) if not args.no_extras else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, ms = oracle_step_rays_per_s(RAYS_PER_GPU, 1, 1, cores)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": f"oracle/live_path.py fwd+bwd+Adam, ONE step of {RAYS_PER_GPU} rays x 512 samples (the GPU "
                         f"arm's batch; slices of {ORACLE_SLICE} rays) after one {ORACLE_WARMUP_RAYS}-ray warm-up step "
                         f"({ms:.0f} ms/step, of which {100 * oracle_step_rays_per_s.fixed_share:.1f} % do not depend "
                         f"on the ray count: zero_grad + Adam over 13.1 M parameters)"}
        if config1 is not None:
            from oracle import config1 as oracle_config1

            c1 = oracle_config1.time_cpu(threads=cores, repeats=3, warmup=1)
            config1["cpu_reference"] = dict(c1, kind="port", what="oracle/config1.py: live_path.run(num_steps=128, "
                                            "upsample_steps=0) with synthetic heads, best of 3 (renderer_semantics.py:"
                                            "123-299 on the host cores)")
            config1["cpu_reference"]["rays_per_s_fwd"] = 4096 / (c1["fwd_ms"] * 1e-3)

    line = {
        "metric": "semantic_nerf_train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": dict(CONFIG, parallelism=f"ray-sharded dp{world}"),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps,
                "api": "SemanticNeRFNetwork.render + nerf_losses + backward + FusedAdam.step (the drop-in modules of "
                       "INTEGRATION.md inside the caller's own loop), pinned host inputs, loss copied to pinned host "
                       "memory every step (non-blocking)",
                "torch_adam": {"value": world * RAYS_PER_GPU * args.steps / (e2e_torch_adam_ms * 1e-3),
                               "ms_per_step": e2e_torch_adam_ms / args.steps,
                               "note": "same loop with the caller's torch.optim.Adam left in place"},
                "blocking_readback": {"value": world * RAYS_PER_GPU * args.steps / (e2e_blocking_ms * 1e-3),
                                      "ms_per_step": e2e_blocking_ms / args.steps,
                                      "note": "same, with float(loss) (a host synchronisation) every step"},
                "engine": {"value": e2e_engine_value, "ms_per_step": e2e_engine_ms / args.steps,
                           "api": "TrainEngine.load_batch(pinned host) + step() + loss readback"}},
        "gpu_launches": launches, "gpu_launches_per_step": by_name,
        "step_impl": "cuda-graph replay of %d kernels" % per_step_launches if not args.no_graph else "eager kernel chain",
        "render": render_info, "trained": trained, "strong_scaling_2p16": strong, "strong_scaling_2p18": sweep,
        "config1": config1,
        "stand_in": stand_in, "occupancy_path": occupancy,
        "gradient_exchange": {"none": "single GPU", "nccl": "NCCL all-reduce + replicated Adam",
                              "peer": "ucsa_adam_exchange over symmetric memory (%s)" % (
                                  "multimem.ld_reduce / multimem.st via NVSwitch" if multicast
                                  else "peer loads / stores")}[exchange_mode],
        "exchange_check": exchange_check,
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
    ap.add_argument("--train-steps", type=int, default=1500,
                    help="optimisation steps before the 'trained' measurement (an extra)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra measurements (render, trained model, strong scaling, config 1)")
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
