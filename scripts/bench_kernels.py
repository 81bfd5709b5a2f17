#!/usr/bin/env python
"""Per-kernel timings at BASELINE config-2 sizes (4096 rays x 256+256 samples) on realistic, ray-ordered samples.

    python scripts/bench_kernels.py [--iters 10]

Prints one line per kernel: ms per launch (CUDA events), samples/s and GB/s against SURVEY.md 8(d) bytes."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from ucsa_neural_rendering_b200 import ops  # noqa: E402
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork  # noqa: E402
from ucsa_neural_rendering_b200.scene import SyntheticScene  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()  # 256 MB write: evicts L2 between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda")
    n, tc, tf = 4096, 256, 256
    t = tc + tf
    scene = SyntheticScene(seed=0, device=dev)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=40).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(1)
    pix = torch.randint(0, scene.W * scene.H, (n,), device=dev, generator=g)
    o, d, dn = scene.rays(0, pix)
    aabb = net.aabb_train
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    grid = net.encoder.grid
    table_h, w_sig = net.encoder.half_params(), net.sigma_net.half_params()
    w_col, w_sem = net.color_net.half_params(), net.semantics_net.half_params()
    f32 = dict(dtype=torch.float32, device=dev)
    f16 = dict(dtype=torch.float16, device=dev)
    nears, fars = ops.near_far_from_aabb(o, d, aabb)
    z_cat = torch.empty(n, t, **f32)
    lin = torch.linspace(0, 1, tc, device=dev)
    ops.sample_coarse(nears, fars, lin, z_cat, tc, perturb=True, seed=3)
    sigma = torch.empty(n, t, **f32)
    h = torch.empty(n, t, 16, **f16)
    enc = torch.empty(n, t, 32, **f16)
    hid = torch.empty(n, t, 64, **f16)
    common = dict(rays_o=o, rays_d=d, aabb=aabb, z_cat=z_cat, sigma=sigma, h=h, enc=enc, hid=hid, tiled=True)
    order = torch.empty(n, t, dtype=torch.int32, device=dev)
    rows = []

    def rec(name, ms, samples, bytes_per_sample):
        rows.append((name, ms, samples / ms * 1e-6, samples * bytes_per_sample / ms * 1e-6))

    ms = timeit(lambda: ops.density_fwd(grid, table_h, w_sig, 4.0, k0=0, k1=tc, **common), args.iters, flush)
    rec("density_fwd (coarse pass, 1.05M samples)", ms, n * tc, 588)
    ops.resample_merge(sigma, z_cat, order, tc, tf, 1.0, seed=3)
    ms = timeit(lambda: ops.resample_merge(sigma, z_cat, order, tc, tf, 1.0, seed=3), args.iters, flush)
    rec("resample_merge (4096 rays)", ms, n * t, 12)
    ms = timeit(lambda: ops.density_fwd(grid, table_h, w_sig, 4.0, k0=tc, k1=t, **common), args.iters, flush)
    rec("density_fwd (fine pass, 1.05M samples)", ms, n * tf, 588)

    w_sorted = torch.empty(n, t, **f32)
    depth = torch.empty(n, **f32)
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    use = torch.empty(n, t, dtype=torch.uint8, device=dev)
    off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.weights_fwd(z_cat, sigma, order, dn.view(-1), 1.0, w_sorted, depth, cnt, use), args.iters, flush)
    rec("weights_fwd", ms, n * t, 16)
    ops.scan_counts(cnt, off)
    k = int(off[-1])
    k_max = n * t
    sel = torch.empty(k_max, dtype=torch.int32, device=dev)
    w_sel = torch.empty(k_max, **f32)
    z_sel = torch.empty(k_max, **f32)
    ms = timeit(lambda: ops.compact_masked(w_sorted, z_cat, order, off, sel, w_sel, z_sel), args.iters, flush)
    rec("compact_masked", ms, n * t, 16)
    rgb = torch.empty(k_max, 3, **f32)
    logits = torch.empty(k_max, 48, **f16)
    hc1, hc2, hs = (torch.empty(ops.tile_rows(k_max), 64, **f16) for _ in range(3))
    image = torch.zeros(n, 3, **f32)
    sem = torch.zeros(n, 40, **f32)
    ms = timeit(lambda: ops.heads_fwd(sel, off, n, t, k_max, d, h, w_col, w_sem, 40, rgb, logits, hc1, hc2, hs,
                                      w_sel=w_sel, image=image, semantics=sem), args.iters, flush)
    rec(f"heads_fwd + composite (K={k} rows)", ms, k, 32 + 12 + 96 + 3 * 128)
    ms = timeit(lambda: ops.composite_fwd(off, w_sel, rgb, logits, n, 40, image, sem), args.iters, flush)
    rec("composite_fwd (stand-alone)", ms, k, 4 + 12 + 96)
    gi, gd, gs = torch.randn(n, 3, **f32), torch.randn(n, **f32), torch.randn(n, 40, **f32)
    d_rgb = torch.empty(k_max, 3, **f32)
    d_log = torch.empty(k_max, 48, **f32)
    d_w = torch.empty(k_max, **f32)
    ms = timeit(lambda: ops.composite_bwd(off, sel, w_sel, z_sel, rgb, logits, gi, gd, gs, dn.view(-1), n, 40, d_rgb,
                                          d_log, d_w), args.iters, flush)
    rec("composite_bwd (stand-alone)", ms, k, 4 + 4 + 12 + 96 + 12 + 192 + 4)
    dh = torch.empty(n, t, 16, **f16)
    g_col = torch.zeros(ops.COLOR_PARAMS, **f32)
    g_sem = torch.zeros(ops.SEM_PARAMS, **f32)
    ms = timeit(lambda: ops.heads_bwd(sel, off, n, t, k_max, d, h, w_col, w_sem, 40, rgb, hc1, hc2, hs, w_sel,
                                      z_sel, gi, gd, gs, dn.view(-1), 128.0, dh, d_w, g_col, g_sem), args.iters, flush)
    rec(f"heads_bwd + composite bwd (K={k} rows)", ms, k, 32 + 3 * 128 + 12 + 96 + 32)
    d_sigma = torch.empty(n, t, **f32)
    ms = timeit(lambda: ops.weights_bwd(z_cat, sigma, order, w_sorted, off, d_w, 1.0, d_sigma), args.iters, flush)
    rec("weights_bwd", ms, n * t, 20)
    g_tab = torch.zeros(net.encoder.params.numel(), **f32)
    g_sig = torch.zeros(ops.SIGMA_PARAMS, **f32)
    ms = timeit(lambda: ops.density_bwd(grid, w_sig, 4.0, rays_o=o, rays_d=d, aabb=aabb, z_cat=z_cat, k0=0, k1=t, h=h,
                                        enc=enc, hid=hid, d_sigma=d_sigma, dh=dh, use_geo=use, loss_scale=128.0,
                                        grad_table=g_tab, grad_w_sigma=g_sig, tiled=True,
                                        replicas=net.encoder.grad_replicas()), args.iters, flush)
    rec("density_bwd (2.1M samples)", ms, n * t, 588)
    # the scatter alone, one thread per (sample, level)
    x01 = torch.rand(n * t, 3, **f32)
    with torch.no_grad():
        zz = z_cat.view(n, t, 1)
        p = (o.view(n, 1, 3) + d.view(n, 1, 3) * zz).clamp(-4, 4)
        x01 = ((p + 4) / 8).reshape(-1, 3).contiguous()
    d_enc = torch.randn(n * t, 32, device=dev).half()
    ms = timeit(lambda: ops.hashgrid_bwd(x01, grid, d_enc, 1.0, g_tab), args.iters, flush)
    rec("hashgrid_bwd alone (thread = sample x level)", ms, n * t, 588)
    enc2 = torch.empty(n * t, 32, **f16)
    ms = timeit(lambda: ops.hashgrid_fwd(x01, table_h, grid, enc2), args.iters, flush)
    rec("hashgrid_fwd alone (thread = sample x level)", ms, n * t, 588)
    ms = timeit(lambda: g_tab.zero_(), args.iters, flush)
    rec("zero grad table (52 MB)", ms, g_tab.numel(), 4)

    print(f"K/S = {k / (n * t):.3f}")
    for name, ms, msps, gbs in rows:
        print(f"{name:50s} {ms:8.3f} ms  {msps:9.1f} Msamples/s  {gbs:9.1f} GB/s")


if __name__ == "__main__":
    main()
