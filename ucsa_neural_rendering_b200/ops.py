"""Tensor-level wrappers over the C ABI (include/ucsa_nerf.h).

Each function checks device / dtype / contiguity, passes raw device pointers plus the current CUDA stream, and
raises ``UcsaError`` on a non-zero status.  Nothing here computes anything on the host.
"""
from __future__ import annotations

import ctypes
import math

import torch

from ._lib import GridDesc, UcsaError, check, lib

MAX_CLASSES = 48
SIGMA_PARAMS = 3072
COLOR_PARAMS = 7168
SEM_PARAMS = 16 * 64 + MAX_CLASSES * 64


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise UcsaError(f"{name} must be a CUDA tensor (the rendering path has no CPU fallback)")
    if not t.is_contiguous():
        raise UcsaError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise UcsaError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def per_level_scale(bound: float) -> float:
    # network_tcnn_semantics.py:34
    return float(2.0 ** (math.log2(2048 * bound / 16) / (16 - 1)))


def make_grid_desc(bound: float, base_resolution: int = 16, log2_hashmap_size: int = 19) -> GridDesc:
    g = GridDesc()
    check(lib().ucsa_grid_desc_init(per_level_scale(bound), base_resolution, log2_hashmap_size, ctypes.byref(g)),
          "grid_desc_init")
    return g


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    n = rays_o.shape[0]
    nears = torch.empty(n, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty_like(nears)
    check(lib().ucsa_near_far_from_aabb(_ptr(rays_o, torch.float32, "rays_o"), _ptr(rays_d, torch.float32, "rays_d"),
                                        _ptr(aabb, torch.float32, "aabb"), n, float(min_near), _ptr(nears),
                                        _ptr(fars), _stream()), "near_far_from_aabb")
    return nears, fars


def near_far_from_aabb_into(rays_o, rays_d, aabb, nears, fars, min_near=0.2):
    check(lib().ucsa_near_far_from_aabb(_ptr(rays_o, torch.float32, "rays_o"), _ptr(rays_d, torch.float32, "rays_d"),
                                        _ptr(aabb, torch.float32, "aabb"), rays_o.shape[0], float(min_near),
                                        _ptr(nears, torch.float32), _ptr(fars, torch.float32), _stream()),
          "near_far_from_aabb")


def sample_coarse(nears, fars, lin, z_cat, tc, *, perturb, t_rand=None, seed=0, ray_base=0, step_dev=None):
    n, t = z_cat.shape
    check(lib().ucsa_sample_coarse(_ptr(nears, torch.float32), _ptr(fars, torch.float32), _ptr(lin, torch.float32),
                                   _ptr(t_rand, torch.float32, "t_rand"), seed, _ptr(step_dev, torch.int32), ray_base,
                                   int(bool(perturb)), n, tc, t, _ptr(z_cat, torch.float32), _stream()),
          "sample_coarse")


def tile_rows(rows):
    """rows of a tile-layout activation buffer (UCSA_TILE_ROWS): whole 128-row tiles"""
    return (rows + 127) // 128 * 128


def _check_tiled(k_max, width, **bufs):
    for name, b in bufs.items():
        if b is not None and b.numel() < tile_rows(k_max) * width:
            raise ValueError(f"{name}: tile-layout buffer needs tile_rows({k_max}) * {width} elements, has {b.numel()}")


def density_fwd(grid, table_h, w_sigma_h, bound, *, xyz=None, rays_o=None, rays_d=None, aabb=None, z_cat=None,
                k0=0, k1=1, sigma, h, enc=None, hid=None, simt=False, tiled=False):
    """tiled: enc / hid are opaque tile-layout buffers of tile_rows(rows) rows (tensor-core kernels only; in ray
    mode see density_tiled); the matching density_bwd call must say the same."""
    if xyz is not None:
        n, t = xyz.shape[0], 1
    else:
        n, t = z_cat.shape
    args = (_ptr(xyz, torch.float32, "xyz"), _ptr(rays_o, torch.float32, "rays_o"),
            _ptr(rays_d, torch.float32, "rays_d"), _ptr(aabb, torch.float32, "aabb"),
            _ptr(z_cat, torch.float32, "z_cat"), n, t, k0, k1, float(bound),
            _ptr(table_h, torch.float16, "table_h"), ctypes.byref(grid),
            _ptr(w_sigma_h, torch.float16, "w_sigma_h"), _ptr(sigma, torch.float32, "sigma"),
            _ptr(h, torch.float16, "h"), _ptr(enc, torch.float16, "enc"), _ptr(hid, torch.float16, "hid"))
    if simt:
        if tiled:
            raise ValueError("the CUDA-core density kernels use row-major enc / hid")
        check(lib().ucsa_density_fwd_simt(*args, _stream()), "density_fwd_simt")
        return
    if tiled:
        _check_tiled(n * t, 32, enc=enc)
        _check_tiled(n * t, 64, hid=hid)
    check(lib().ucsa_density_fwd(*args, int(tiled), _stream()), "density_fwd")


def density_tiled(tc, tf):
    """whether the ray-mode density passes over Tc coarse + Tf fine slots may use tile-layout enc / hid"""
    return tc % 128 == 0 and tf % 128 == 0


def density_bwd(grid, w_sigma_h, bound, *, xyz=None, rays_o=None, rays_d=None, aabb=None, z_cat=None, k0=0, k1=1,
                h, enc, hid, d_sigma, dh, use_geo, loss_scale, grad_table, grad_w_sigma, simt=False, replicas=None,
                tiled=False, dh2=None):
    """replicas: zero-filled fp32 [R, 2*dense_entries] (see grad_replicas()); folded into grad_table before returning"""
    if xyz is not None:
        n, t = xyz.shape[0], 1
    else:
        n, t = z_cat.shape
    head = (_ptr(xyz, torch.float32, "xyz"), _ptr(rays_o, torch.float32), _ptr(rays_d, torch.float32),
            _ptr(aabb, torch.float32), _ptr(z_cat, torch.float32), n, t, k0, k1, float(bound), ctypes.byref(grid),
            _ptr(w_sigma_h, torch.float16), _ptr(h, torch.float16), _ptr(enc, torch.float16), _ptr(hid, torch.float16))
    mid = (_ptr(d_sigma, torch.float32, "d_sigma"), _ptr(dh, torch.float16, "dh"), _ptr(use_geo, torch.uint8, "use_geo"),
           float(loss_scale), _ptr(grad_table, torch.float32, "grad_table"))
    tail = (_ptr(grad_w_sigma, torch.float32, "grad_w_sigma"), _stream())
    if simt:
        if tiled or dh2 is not None:
            raise ValueError("the CUDA-core density kernels use row-major enc / hid and a single dh")
        check(lib().ucsa_density_bwd_simt(*head, *mid, *tail), "density_bwd_simt")
        return
    mid = (mid[0], mid[1], _ptr(dh2, torch.float16, "dh2")) + mid[2:]
    if tiled:
        _check_tiled(n * t, 32, enc=enc)
        _check_tiled(n * t, 64, hid=hid)
    n_rep = 0 if replicas is None else replicas.shape[0]
    check(lib().ucsa_density_bwd(*head, int(tiled), *mid, _ptr(replicas, torch.float32, "replicas"), n_rep, *tail),
          "density_bwd")
    if n_rep:
        check(lib().ucsa_reduce_grad_replicas(_ptr(replicas, torch.float32), n_rep, ctypes.byref(grid),
                                              _ptr(grad_table, torch.float32), _stream()), "reduce_grad_replicas")


def dense_entries(grid) -> int:
    for lvl in range(16):
        if grid.hashed[lvl]:
            return int(grid.offset[lvl])
    return int(grid.total_entries)


def grad_replicas(grid, device, n_replicas=16):
    """Zero-filled private copies of the dense levels' gradient (kept zero by ucsa_reduce_grad_replicas)."""
    return torch.zeros(n_replicas, 2 * dense_entries(grid), dtype=torch.float32, device=device)


def resample_merge(sigma, z_cat, order, tc, tf, density_scale, *, u=None, seed=0, ray_base=0, step_dev=None):
    n = z_cat.shape[0]
    check(lib().ucsa_resample_merge(_ptr(sigma, torch.float32), _ptr(z_cat, torch.float32), _ptr(u, torch.float32, "u"),
                                    seed, _ptr(step_dev, torch.int32), ray_base, n, tc, tf, float(density_scale),
                                    _ptr(order, torch.int32, "order"), _stream()), "resample_merge")


def weights_fwd(z_cat, sigma, order, direction_norms, density_scale, w_sorted, depth, ray_count, use_geo):
    n, t = z_cat.shape
    check(lib().ucsa_weights_fwd(_ptr(z_cat, torch.float32), _ptr(sigma, torch.float32), _ptr(order, torch.int32),
                                 _ptr(direction_norms, torch.float32), n, t, float(density_scale),
                                 _ptr(w_sorted, torch.float32), _ptr(depth, torch.float32), _ptr(ray_count, torch.int32),
                                 _ptr(use_geo, torch.uint8), _stream()), "weights_fwd")


def weights_scratch(n_rays, device):
    """Zero-filled scratch block of ucsa_weights_compact (the kernel leaves it zeroed)."""
    return torch.zeros(2 + 2 * ((n_rays + 3) // 4), dtype=torch.int32, device=device)


def weights_compact(z_cat, sigma, order, direction_norms, density_scale, w_sorted, depth, ray_off, use_geo, sel, w_sel,
                    z_sel, scratch):
    """weights_fwd + scan_counts + compact_masked in one launch (decoupled look-back scan over the rays)."""
    n, t = z_cat.shape
    if scratch.numel() < 2 + 2 * ((n + 3) // 4):
        raise ValueError("weights_compact: scratch block too small for this ray count")
    check(lib().ucsa_weights_compact(_ptr(z_cat, torch.float32), _ptr(sigma, torch.float32), _ptr(order, torch.int32),
                                     _ptr(direction_norms, torch.float32), n, t, float(density_scale),
                                     _ptr(w_sorted, torch.float32), _ptr(depth, torch.float32),
                                     _ptr(ray_off, torch.int32), _ptr(use_geo, torch.uint8), _ptr(sel, torch.int32),
                                     _ptr(w_sel, torch.float32), _ptr(z_sel, torch.float32),
                                     _ptr(scratch, torch.int32), _stream()), "weights_compact")


def scan_counts(ray_count, ray_off):
    check(lib().ucsa_scan_counts(_ptr(ray_count, torch.int32), ray_count.shape[0], _ptr(ray_off, torch.int32),
                                 _stream()), "scan_counts")


def compact_masked(w_sorted, z_cat, order, ray_off, sel, w_sel, z_sel):
    n, t = z_cat.shape
    check(lib().ucsa_compact_masked(_ptr(w_sorted, torch.float32), _ptr(z_cat, torch.float32), _ptr(order, torch.int32),
                                    _ptr(ray_off, torch.int32), n, t, _ptr(sel, torch.int32), _ptr(w_sel, torch.float32),
                                    _ptr(z_sel, torch.float32), _stream()), "compact_masked")


def heads_fwd(sel, ray_off, n_rays, t, k_max, rays_d, h, w_color_h, w_sem_h, n_classes, rgb, logits=None,
              hc1=None, hc2=None, hs=None, w_sel=None, image=None, semantics=None):
    """colour + semantic heads; with w_sel / image / semantics also the compositing (image, semantics zero-filled).
    hc1 / hc2 / hs: opaque tile-layout buffers of tile_rows(k_max) * 64 halves (all three or none)."""
    if (hc1 is None) != (hc2 is None) or (hc1 is None) != (hs is None):
        raise ValueError("heads_fwd: pass hc1, hc2 and hs together")
    _check_tiled(k_max, 64, hc1=hc1, hc2=hc2, hs=hs)
    check(lib().ucsa_heads_fwd(_ptr(sel, torch.int32), _ptr(ray_off, torch.int32), n_rays, t, k_max,
                               _ptr(rays_d, torch.float32), _ptr(h, torch.float16), _ptr(w_color_h, torch.float16),
                               _ptr(w_sem_h, torch.float16), n_classes, _ptr(w_sel, torch.float32),
                               _ptr(rgb, torch.float32), _ptr(logits, torch.float16), _ptr(hc1, torch.float16),
                               _ptr(hc2, torch.float16), _ptr(hs, torch.float16), _ptr(image, torch.float32),
                               _ptr(semantics, torch.float32), _stream()), "heads_fwd")


def heads_bwd(sel, ray_off, n_rays, t, k_max, rays_d, h, w_color_h, w_sem_h, n_classes, rgb, hc1, hc2, hs,
              w_sel, z_sel, g_image, g_depth, g_semantics, direction_norms, loss_scale, dh, d_w_sel, grad_w_color,
              grad_w_sem, dh_sem=None):
    """dh_sem: optional second buffer for the semantic head's share of dL/dgeo_feat (the two kernels then run
    concurrently; pass it to density_bwd as dh2)"""
    _check_tiled(k_max, 64, hc1=hc1, hc2=hc2, hs=hs)
    check(lib().ucsa_heads_bwd(_ptr(sel, torch.int32), _ptr(ray_off, torch.int32), n_rays, t, k_max,
                               _ptr(rays_d, torch.float32), _ptr(h, torch.float16), _ptr(w_color_h, torch.float16),
                               _ptr(w_sem_h, torch.float16), n_classes, _ptr(rgb, torch.float32),
                               _ptr(hc1, torch.float16), _ptr(hc2, torch.float16),
                               _ptr(hs, torch.float16), _ptr(w_sel, torch.float32), _ptr(z_sel, torch.float32),
                               _ptr(g_image, torch.float32, "g_image"), _ptr(g_depth, torch.float32, "g_depth"),
                               _ptr(g_semantics, torch.float32, "g_semantics"), _ptr(direction_norms, torch.float32),
                               float(loss_scale), _ptr(dh, torch.float16), _ptr(dh_sem, torch.float16, "dh_sem"),
                               _ptr(d_w_sel, torch.float32), _ptr(grad_w_color, torch.float32),
                               _ptr(grad_w_sem, torch.float32), _stream()),
          "heads_bwd")


def heads_fwd_simt(sel, ray_off, n_rays, t, k_max, rays_d, h, w_color_h, w_sem_h, n_classes, rgb, logits,
                   hc1=None, hc2=None, hs=None):
    """CUDA-core cross-check of the heads (no fused compositing)"""
    check(lib().ucsa_heads_fwd_simt(_ptr(sel, torch.int32), _ptr(ray_off, torch.int32), n_rays, t, k_max,
                                    _ptr(rays_d, torch.float32), _ptr(h, torch.float16), _ptr(w_color_h, torch.float16),
                                    _ptr(w_sem_h, torch.float16), n_classes, _ptr(rgb, torch.float32),
                                    _ptr(logits, torch.float16), _ptr(hc1, torch.float16), _ptr(hc2, torch.float16),
                                    _ptr(hs, torch.float16), _stream()), "heads_fwd_simt")


def heads_bwd_simt(sel, ray_off, n_rays, t, k_max, rays_d, h, w_color_h, w_sem_h, n_classes, rgb, hc1, hc2, hs, d_rgb,
                   d_logits, loss_scale, dh, grad_w_color, grad_w_sem):
    check(lib().ucsa_heads_bwd_simt(_ptr(sel, torch.int32), _ptr(ray_off, torch.int32), n_rays, t, k_max,
                                    _ptr(rays_d, torch.float32), _ptr(h, torch.float16), _ptr(w_color_h, torch.float16),
                                    _ptr(w_sem_h, torch.float16), n_classes, _ptr(rgb, torch.float32),
                                    _ptr(hc1, torch.float16), _ptr(hc2, torch.float16), _ptr(hs, torch.float16),
                                    _ptr(d_rgb, torch.float32), _ptr(d_logits, torch.float32), float(loss_scale),
                                    _ptr(dh, torch.float16), _ptr(grad_w_color, torch.float32),
                                    _ptr(grad_w_sem, torch.float32), _stream()), "heads_bwd_simt")


def composite_fwd(ray_off, w_sel, rgb, logits, n_rays, n_classes, image, semantics):
    check(lib().ucsa_composite_fwd(_ptr(ray_off, torch.int32), _ptr(w_sel, torch.float32), _ptr(rgb, torch.float32),
                                   _ptr(logits, torch.float16), n_rays, n_classes, _ptr(image, torch.float32),
                                   _ptr(semantics, torch.float32), _stream()), "composite_fwd")


def composite_bwd(ray_off, sel, w_sel, z_sel, rgb, logits, g_image, g_depth, g_semantics, direction_norms, n_rays,
                  n_classes, d_rgb, d_logits, d_w_sel):
    check(lib().ucsa_composite_bwd(_ptr(ray_off, torch.int32), _ptr(sel, torch.int32), _ptr(w_sel, torch.float32),
                                   _ptr(z_sel, torch.float32), _ptr(rgb, torch.float32), _ptr(logits, torch.float16),
                                   _ptr(g_image, torch.float32, "g_image"), _ptr(g_depth, torch.float32, "g_depth"),
                                   _ptr(g_semantics, torch.float32, "g_semantics"),
                                   _ptr(direction_norms, torch.float32), n_rays, n_classes,
                                   _ptr(d_rgb, torch.float32), _ptr(d_logits, torch.float32),
                                   _ptr(d_w_sel, torch.float32), _stream()), "composite_bwd")


def weights_bwd(z_cat, sigma, order, w_sorted, ray_off, d_w_sel, density_scale, d_sigma):
    n, t = z_cat.shape
    check(lib().ucsa_weights_bwd(_ptr(z_cat, torch.float32), _ptr(sigma, torch.float32), _ptr(order, torch.int32),
                                 _ptr(w_sorted, torch.float32), _ptr(ray_off, torch.int32), _ptr(d_w_sel, torch.float32),
                                 n, t, float(density_scale), _ptr(d_sigma, torch.float32), _stream()), "weights_bwd")


def composite_dense_fwd(sigma, z, rgb, prob, direction_norms, density_scale, weights, depth, image, semantics):
    n, t = sigma.shape
    c = prob.shape[-1]
    check(lib().ucsa_composite_dense_fwd(_ptr(sigma, torch.float32, "sigma"), _ptr(z, torch.float32, "z"),
                                         _ptr(rgb, torch.float32, "rgb"), _ptr(prob, torch.float32, "prob"),
                                         _ptr(direction_norms, torch.float32, "direction_norms"), n, t, c,
                                         float(density_scale), _ptr(weights, torch.float32), _ptr(depth, torch.float32),
                                         _ptr(image, torch.float32), _ptr(semantics, torch.float32), _stream()),
          "composite_dense_fwd")


def composite_dense_bwd(sigma, z, rgb, weights, direction_norms, g_depth, g_image, g_semantics, density_scale,
                        d_sigma, d_rgb, d_prob):
    n, t = sigma.shape
    c = g_semantics.shape[-1]
    check(lib().ucsa_composite_dense_bwd(_ptr(sigma, torch.float32), _ptr(z, torch.float32), _ptr(rgb, torch.float32),
                                         _ptr(weights, torch.float32), _ptr(direction_norms, torch.float32),
                                         _ptr(g_depth, torch.float32, "g_depth"), _ptr(g_image, torch.float32, "g_image"),
                                         _ptr(g_semantics, torch.float32, "g_semantics"), n, t, c,
                                         float(density_scale), _ptr(d_sigma, torch.float32), _ptr(d_rgb, torch.float32),
                                         _ptr(d_prob, torch.float32), _stream()), "composite_dense_bwd")


def hashgrid_fwd(x01, table_h, grid, enc):
    check(lib().ucsa_hashgrid_fwd(_ptr(x01, torch.float32, "x01"), x01.shape[0], _ptr(table_h, torch.float16),
                                  ctypes.byref(grid), _ptr(enc, torch.float16), _stream()), "hashgrid_fwd")


def hashgrid_bwd(x01, grid, d_enc, inv_loss_scale, grad_table):
    check(lib().ucsa_hashgrid_bwd(_ptr(x01, torch.float32), x01.shape[0], ctypes.byref(grid),
                                  _ptr(d_enc, torch.float16, "d_enc"), float(inv_loss_scale),
                                  _ptr(grad_table, torch.float32), _stream()), "hashgrid_bwd")


def hashgrid_indices(x01, grid):
    idx = torch.empty(x01.shape[0], 16, 8, dtype=torch.int32, device=x01.device)
    check(lib().ucsa_hashgrid_indices(_ptr(x01, torch.float32), x01.shape[0], ctypes.byref(grid), idx.data_ptr(),
                                      _stream()), "hashgrid_indices")
    return idx


def sh4_fwd(d01, out):
    check(lib().ucsa_sh4_fwd(_ptr(d01, torch.float32, "d01"), d01.shape[0], _ptr(out, torch.float16), _stream()),
          "sh4_fwd")


def _dims_array(dims):
    return (ctypes.c_uint32 * len(dims))(*dims)


def mlp_fwd(x_h, w_h, dims, y_h, acts_h=None, simt=False):
    arr = _dims_array(dims)
    fn = lib().ucsa_mlp_fwd_simt if simt else lib().ucsa_mlp_fwd
    check(fn(_ptr(x_h, torch.float16, "x"), x_h.shape[0], _ptr(w_h, torch.float16),
                             ctypes.cast(arr, ctypes.c_void_p), len(dims) - 1, _ptr(y_h, torch.float16),
                             _ptr(acts_h, torch.float16), _stream()), "mlp_fwd")


def mlp_bwd(x_h, w_h, dims, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, simt=False):
    arr = _dims_array(dims)
    fn = lib().ucsa_mlp_bwd_simt if simt else lib().ucsa_mlp_bwd
    check(fn(_ptr(x_h, torch.float16), x_h.shape[0], _ptr(w_h, torch.float16),
                             ctypes.cast(arr, ctypes.c_void_p), len(dims) - 1, _ptr(acts_h, torch.float16),
                             _ptr(dy_h, torch.float16, "dy"), float(inv_loss_scale), _ptr(dx_h, torch.float16),
                             _ptr(grad_w, torch.float32), _stream()), "mlp_bwd")


def cast_f32_to_f16(src, dst):
    check(lib().ucsa_cast_f32_to_f16(_ptr(src, torch.float32, "src"), src.numel(), _ptr(dst, torch.float16, "dst"),
                                     _stream()), "cast_f32_to_f16")


def adam_step(param, grad, exp_avg, exp_avg_sq, param_h, *, lr, beta1, beta2, eps, weight_decay, grad_scale_inv,
              found_inf, step, step_dev=None, skipped_dev=None, wd_begin=0, grad_scale_dev=None):
    """torch.optim.Adam (L2 weight decay) with GradScaler semantics: gradients are multiplied by grad_scale_inv; when
    found_inf[0] != 0 nothing is written.  With step_dev, Adam's step count is step_dev[0] - skipped_dev[0].
    weight_decay applies to parameters with index >= wd_begin."""
    check(lib().ucsa_adam_step(_ptr(param, torch.float32), _ptr(grad, torch.float32), _ptr(exp_avg, torch.float32),
                               _ptr(exp_avg_sq, torch.float32), _ptr(param_h, torch.float16), param.numel(), float(lr),
                               float(beta1), float(beta2), float(eps), float(weight_decay), int(wd_begin),
                               float(grad_scale_inv), _ptr(grad_scale_dev, torch.float32, "grad_scale"),
                               _ptr(found_inf, torch.float32), int(step),
                               _ptr(step_dev, torch.int32), _ptr(skipped_dev, torch.int32), _stream()), "adam_step")


def grad_check(grad, found_inf, scratch2, skipped_dev=None):
    """found_inf[0] = 1.0 if grad holds an inf / NaN else 0.0 (GradScaler's check); bumps skipped_dev[0] when set.
    scratch2: two zero-initialised int32 words owned by the caller."""
    check(lib().ucsa_grad_check(_ptr(grad, torch.float32, "grad"), grad.numel(), _ptr(found_inf, torch.float32),
                                _ptr(skipped_dev, torch.int32), _ptr(scratch2, torch.int32, "scratch2"), _stream()),
          "grad_check")


def adam_exchange(peer, exp_avg, exp_avg_sq, *, wd_begin, lr, beta1, beta2, eps, weight_decay, step, step_dev=None,
                  found_inf_ptrs=None, skipped_dev=None, broadcast_masters=True):
    """Fused reduce-scatter + Adam + all-gather over peer memory; `peer` is a parallel.PeerExchange.  The caller puts a
    cross-rank barrier before (gradients complete) and after (parameters visible).  found_inf_ptrs: the ranks'
    overflow flags (any set -> every rank skips); broadcast_masters=False keeps fp32 masters on the owner."""
    arr = ctypes.c_uint64 * peer.world
    mc = peer.multicast
    check(lib().ucsa_adam_exchange(arr(*peer.grad_ptrs), arr(*peer.param_ptrs), arr(*peer.param_h_ptrs),
                                   peer.mc_grad if mc else None, peer.mc_param if mc else None,
                                   peer.mc_param_h if mc else None, peer.world, peer.rank, peer.begin, peer.end,
                                   int(wd_begin), _ptr(exp_avg, torch.float32), _ptr(exp_avg_sq, torch.float32),
                                   float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                   _ptr(step_dev, torch.int32),
                                   arr(*found_inf_ptrs) if found_inf_ptrs is not None else None,
                                   _ptr(skipped_dev, torch.int32), int(bool(broadcast_masters)), _stream()),
          "adam_exchange")


LOSS_SCRATCH_BYTES = 2048  # UCSA_LOSS_SCRATCH_BYTES


def loss_scratch(device):
    """Zero-filled scratch block of ucsa_nerf_loss; allocate one per engine / workspace (never shared across streams)."""
    return torch.zeros(LOSS_SCRATCH_BYTES // 4, dtype=torch.float32, device=device)


def nerf_loss(image, depth, semantics, gt_rgb, labels, gt_depth, uom, w_sem, w_depth, global_scale, loss4, g_image,
              g_depth, g_semantics, scratch):
    n, c = semantics.shape
    half = gt_rgb.dtype == torch.float16
    check(lib().ucsa_nerf_loss(_ptr(image, torch.float32), _ptr(depth, torch.float32), _ptr(semantics, torch.float32),
                               _ptr(gt_rgb, torch.float16) if half else None,
                               None if half else _ptr(gt_rgb, torch.float32), _ptr(labels, torch.int64, "labels"),
                               _ptr(gt_depth, torch.float32, "gt_depth"), n, c, float(uom), float(w_sem), float(w_depth),
                               float(global_scale), _ptr(loss4, torch.float32), _ptr(g_image, torch.float32),
                               _ptr(g_depth, torch.float32), _ptr(g_semantics, torch.float32),
                               _ptr(scratch, torch.float32, "scratch"), _stream()), "nerf_loss")


# ---------------------------------------------------------------------------------------------- occupancy-grid path
MARCH_STAGE_MAX_RAYS = 1 << 16  # staging = 4 KB per ray (256 MB at this size); larger batches march twice


def march_rays_train(rays_o, rays_d, grid, bitfield, mean_density, bound, dt_gamma, nears, fars, max_points, counter,
                     perturb, staged=True):
    """staged: record the parameter of every occupied step in the counting pass and write the samples with a
    sample-parallel pass (each ray is marched once instead of twice); same samples either way"""
    n = rays_o.shape[0]
    dev = rays_o.device
    c, h = grid.shape[0], grid.shape[1]
    xyzs = torch.zeros(max_points, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(max_points, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(max_points, 2, dtype=torch.float32, device=dev)
    rays = torch.empty(n, 3, dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * n + 2 * ((n + 1023) // 1024) + 2, dtype=torch.int32, device=dev)  # UCSA_MARCH_SCRATCH_INTS
    t_stage = torch.empty(n * 1024, dtype=torch.float32, device=dev) if staged and 0 < n <= MARCH_STAGE_MAX_RAYS else None
    check(lib().ucsa_march_rays_train(_ptr(rays_o, torch.float32), _ptr(rays_d, torch.float32),
                                      _ptr(grid, torch.float32, "density_grid"), _ptr(bitfield, torch.int32, "bitfield"),
                                      float(mean_density), float(bound), float(dt_gamma), n, c, h, max_points,
                                      _ptr(nears, torch.float32), _ptr(fars, torch.float32), _ptr(xyzs), _ptr(dirs),
                                      _ptr(deltas), _ptr(rays), _ptr(counter, torch.int32, "counter"), int(perturb),
                                      _ptr(scratch), _ptr(t_stage), _stream()), "march_rays_train")
    return xyzs, dirs, deltas, rays


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, grid, bitfield, mean_density,
               nears, fars, max_points, perturb):
    dev = rays_o.device
    c, h = grid.shape[0], grid.shape[1]
    xyzs = torch.zeros(max_points, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(max_points, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(max_points, 2, dtype=torch.float32, device=dev)
    check(lib().ucsa_march_rays(n_alive, n_step, _ptr(rays_alive, torch.int32), _ptr(rays_t, torch.float32),
                                _ptr(rays_o, torch.float32), _ptr(rays_d, torch.float32), float(bound), float(dt_gamma),
                                c, h, _ptr(grid, torch.float32), _ptr(bitfield, torch.int32), float(mean_density),
                                _ptr(nears, torch.float32), _ptr(fars, torch.float32), _ptr(xyzs), _ptr(dirs),
                                _ptr(deltas), int(perturb), _stream()), "march_rays")
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, local_semantics, deltas, rays, n_classes, weights_sum, depth, image,
                                 semantics, logits=None):
    """logits: fp16 [M, ld] of the semantic head instead of probabilities (soft-max inside the kernel)"""
    check(lib().ucsa_composite_rays_train_forward(_ptr(sigmas, torch.float32), _ptr(rgbs, torch.float32),
                                                  _ptr(local_semantics, torch.float32), _ptr(logits, torch.float16),
                                                  0 if logits is None else logits.shape[1],
                                                  _ptr(deltas, torch.float32),
                                                  _ptr(rays, torch.int32), sigmas.shape[0], rays.shape[0], n_classes,
                                                  _ptr(weights_sum, torch.float32), _ptr(depth, torch.float32),
                                                  _ptr(image, torch.float32), _ptr(semantics, torch.float32),
                                                  _stream()), "composite_rays_train_forward")


def composite_rays_train_backward(grad_ws, grad_image, grad_sem, sigmas, rgbs, deltas, rays, weights_sum, image,
                                  n_classes, grad_sigmas, grad_rgbs, grad_local_sem):
    check(lib().ucsa_composite_rays_train_backward(_ptr(grad_ws, torch.float32), _ptr(grad_image, torch.float32),
                                                   _ptr(grad_sem, torch.float32), _ptr(sigmas, torch.float32),
                                                   _ptr(rgbs, torch.float32), _ptr(deltas, torch.float32),
                                                   _ptr(rays, torch.int32), _ptr(weights_sum, torch.float32),
                                                   _ptr(image, torch.float32), sigmas.shape[0], rays.shape[0], n_classes,
                                                   _ptr(grad_sigmas, torch.float32), _ptr(grad_rgbs, torch.float32),
                                                   _ptr(grad_local_sem, torch.float32), _stream()),
          "composite_rays_train_backward")


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, local_semantics, deltas, n_classes, weights_sum,
                   depth, image, semantics, logits=None):
    """logits: fp16 [M, ld] of the semantic head instead of probabilities (soft-max inside the kernel)"""
    check(lib().ucsa_composite_rays(n_alive, n_step, _ptr(rays_alive, torch.int32), _ptr(rays_t, torch.float32),
                                    _ptr(sigmas, torch.float32), _ptr(rgbs, torch.float32),
                                    _ptr(local_semantics, torch.float32), _ptr(logits, torch.float16, "logits"),
                                    0 if logits is None else logits.shape[-1], _ptr(deltas, torch.float32), n_classes,
                                    _ptr(weights_sum, torch.float32), _ptr(depth, torch.float32),
                                    _ptr(image, torch.float32), _ptr(semantics, torch.float32), _stream()),
          "composite_rays")


def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, scratch=None):
    """scratch: optional int32 [ceil(n_alive / 1024)] for the many-CTA form (allocated here when the list is long)"""
    if scratch is None and n_alive > 4096:
        scratch = torch.empty((n_alive + 1023) // 1024, dtype=torch.int32, device=rays_alive.device)
    check(lib().ucsa_compact_rays(n_alive, _ptr(rays_alive, torch.int32), _ptr(rays_alive_old, torch.int32),
                                  _ptr(rays_t, torch.float32), _ptr(rays_t_old, torch.float32),
                                  _ptr(alive_counter, torch.int32), _ptr(scratch, torch.int32), _stream()),
          "compact_rays")


def grid_update(density_grid, fresh, decay):
    check(lib().ucsa_grid_update(_ptr(density_grid, torch.float32), _ptr(fresh, torch.float32), density_grid.numel(),
                                 float(decay), _stream()), "grid_update")


def grid_packbits(density_grid, mean_density, bitfield):
    check(lib().ucsa_grid_packbits(_ptr(density_grid, torch.float32), density_grid.numel(), float(mean_density),
                                   _ptr(bitfield, torch.int32), _stream()), "grid_packbits")


def grid_density(grid, table_h, w_sigma_h, bound, cascades, grid_h, seed, sigma_cells):
    """sigma at one jittered point per occupancy-grid cell ([cascades, H, H, H], marching order)"""
    check(lib().ucsa_grid_density(_ptr(table_h, torch.float16, "table_h"), ctypes.byref(grid),
                                  _ptr(w_sigma_h, torch.float16, "w_sigma_h"), float(bound), int(cascades), int(grid_h),
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(sigma_cells, torch.float32, "sigma_cells"),
                                  _stream()), "grid_density")


def grid_update_pack(density_grid, fresh, decay, fresh_scale, sum_scratch, mean_dev, bitfield):
    check(lib().ucsa_grid_update_pack(_ptr(density_grid, torch.float32), _ptr(fresh, torch.float32),
                                      density_grid.numel(), float(decay), float(fresh_scale),
                                      _ptr(sum_scratch, torch.float64, "sum_scratch"),
                                      _ptr(mean_dev, torch.float32, "mean_dev"), _ptr(bitfield, torch.int32),
                                      _stream()), "grid_update_pack")


# ---------------------------------------------------------------------------------------------- front / back ends
def generate_rays(pose, intrinsics, height, width, inds=None, n=None):
    """Pinhole rays of one view (ngp_utils.py:28-70, joint_train_lightning_net.py:109-151).  pose: device [4,4] f32
    cam2world; intrinsics (fx, fy, cx, cy); inds: device int64 pixel indices (row-major) or None for pixels 0..n-1
    (default: the whole image).  -> rays_o [N,3], rays_d [N,3], direction_norms [N]"""
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    if inds is None:
        n = height * width if n is None else n
    else:
        n = inds.numel()
    dev = pose.device
    rays_o = torch.empty(n, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    norms = torch.empty(n, dtype=torch.float32, device=dev)
    check(lib().ucsa_generate_rays(_ptr(pose, torch.float32, "pose"), fx, fy, cx, cy, width, height,
                                   _ptr(inds, torch.int64, "inds"), n, _ptr(rays_o), _ptr(rays_d), _ptr(norms),
                                   _stream()), "generate_rays")
    return rays_o, rays_d, norms


def gather_gt(image_h, inds, labels=None, depth=None):
    """Ground truth of the sampled pixels (joint_train_lightning_net.py:180-187).  image_h: device fp16 [C,H,W];
    labels: int64 [H,W] or None; depth: f32 [H,W] or None.  -> gt_rgb fp16 [N,C], labels [N] | None, depth [N] | None"""
    c = image_h.shape[0]
    hw = image_h[0].numel()
    n = inds.numel()
    dev = image_h.device
    gt_rgb = torch.empty(n, c, dtype=torch.float16, device=dev)
    gt_labels = torch.empty(n, dtype=torch.int64, device=dev) if labels is not None else None
    gt_depth = torch.empty(n, dtype=torch.float32, device=dev) if depth is not None else None
    check(lib().ucsa_gather_gt(_ptr(image_h, torch.float16, "image"), _ptr(labels, torch.int64, "labels"),
                               _ptr(depth, torch.float32, "depth"), hw, c, _ptr(inds, torch.int64, "inds"), n,
                               _ptr(gt_rgb), _ptr(gt_labels), _ptr(gt_depth), _stream()), "gather_gt")
    return gt_rgb, gt_labels, gt_depth


def label_epilogue_into(semantics, image, labels, rgb, bgr=False):
    """ucsa_label_epilogue into caller-owned u8 buffers (labels [N] | None, rgb [N,3] | None)"""
    n, c = semantics.shape
    check(lib().ucsa_label_epilogue(_ptr(image, torch.float32, "image"), _ptr(semantics, torch.float32, "semantics"),
                                    n, c, int(bgr), _ptr(labels, torch.uint8), _ptr(rgb, torch.uint8), _stream()),
          "label_epilogue")


def label_epilogue(semantics, image=None, bgr=False, want_labels=True):
    """Pseudo-label epilogue of rendered pixels (joint_train_lightning_net.py:246-250,755-768).  semantics [N,C] f32,
    image [N,3] f32 or None.  -> label_u8 [N] (argmax + 1) | None, rgb_u8 [N,3] | None"""
    n, c = semantics.shape
    dev = semantics.device
    labels = torch.empty(n, dtype=torch.uint8, device=dev) if want_labels else None
    rgb = torch.empty(n, 3, dtype=torch.uint8, device=dev) if image is not None else None
    check(lib().ucsa_label_epilogue(_ptr(image, torch.float32, "image"), _ptr(semantics, torch.float32, "semantics"),
                                    n, c, int(bgr), _ptr(labels), _ptr(rgb), _stream()), "label_epilogue")
    return labels, rgb
