"""ctypes wrapper of oracle/raymarch_ref.c (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build_c_oracle())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pcg32_stream(initstate, initseq, n):
    u = np.zeros(n, dtype=np.uint32)
    f = np.zeros(n, dtype=np.float32)
    lib().ref_pcg32_stream(ctypes.c_uint64(initstate), ctypes.c_uint64(initseq), n, _p(u), _p(f))
    return u, f


def near_far(rays_o, rays_d, aabb, min_near=0.2):
    o, d, a = _f(rays_o), _f(rays_d), _f(aabb)
    n = o.shape[0]
    nears, fars = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib().ref_near_far(_p(o), _p(d), _p(a), n, ctypes.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def march_rays_train(rays_o, rays_d, grid, mean_density, bound, dt_gamma, nears, fars, max_points, perturb=0):
    o, d, g, ne, fa = _f(rays_o), _f(rays_d), _f(grid), _f(nears), _f(fars)
    n = o.shape[0]
    c, h = g.shape[0], g.shape[1]
    xyzs = np.zeros((max_points, 3), np.float32)
    dirs = np.zeros((max_points, 3), np.float32)
    deltas = np.zeros((max_points, 2), np.float32)
    rays = np.zeros((n, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().ref_march_rays_train(_p(o), _p(d), _p(g), ctypes.c_float(mean_density), ctypes.c_float(bound),
                               ctypes.c_float(dt_gamma), n, c, h, max_points, _p(ne), _p(fa), _p(xyzs), _p(dirs),
                               _p(deltas), _p(rays), _p(counter), int(perturb))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, grid, mean_density, nears, fars,
               perturb=0):
    o, d, g, ne, fa, rt = _f(rays_o), _f(rays_d), _f(grid), _f(nears), _f(fars), _f(rays_t)
    ra = np.ascontiguousarray(rays_alive, dtype=np.int32)
    c, h = g.shape[0], g.shape[1]
    m = n_alive * n_step
    xyzs, dirs, deltas = np.zeros((m, 3), np.float32), np.zeros((m, 3), np.float32), np.zeros((m, 2), np.float32)
    lib().ref_march_rays(n_alive, n_step, _p(ra), _p(rt), _p(o), _p(d), ctypes.c_float(bound), ctypes.c_float(dt_gamma),
                         c, h, _p(g), ctypes.c_float(mean_density), _p(ne), _p(fa), _p(xyzs), _p(dirs), _p(deltas),
                         int(perturb))
    return xyzs, dirs, deltas


def composite_train_fwd(sigmas, rgbs, deltas, rays):
    s, c, dl = _f(sigmas), _f(rgbs), _f(deltas)
    r = np.ascontiguousarray(rays, dtype=np.int32)
    m, n = s.shape[0], r.shape[0]
    ws, depth, image = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    lib().ref_composite_train_fwd(_p(s), _p(c), _p(dl), _p(r), m, n, _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_train_bwd(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image):
    gw, gi, s, c, dl, ws, im = _f(grad_ws), _f(grad_image), _f(sigmas), _f(rgbs), _f(deltas), _f(weights_sum), _f(image)
    r = np.ascontiguousarray(rays, dtype=np.int32)
    m, n = s.shape[0], r.shape[0]
    gs, gr = np.zeros(m, np.float32), np.zeros((m, 3), np.float32)
    lib().ref_composite_train_bwd(_p(gw), _p(gi), _p(s), _p(c), _p(dl), _p(r), _p(ws), _p(im), m, n, _p(gs), _p(gr))
    return gs, gr


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    ra = np.ascontiguousarray(rays_alive, dtype=np.int32)
    rt, s, c, dl = _f(rays_t).copy(), _f(sigmas), _f(rgbs), _f(deltas)
    ws, dp, im = _f(weights_sum).copy(), _f(depth).copy(), _f(image).copy()
    lib().ref_composite_rays(n_alive, n_step, _p(ra), _p(rt), _p(s), _p(c), _p(dl), _p(ws), _p(dp), _p(im))
    return rt, ws, dp, im


def compact_rays(n_alive, rays_alive_old, rays_t_old):
    ra_old = np.ascontiguousarray(rays_alive_old, dtype=np.int32)
    rt_old = _f(rays_t_old)
    ra, rt = np.zeros_like(ra_old), np.zeros_like(rt_old)
    cnt = np.zeros(1, np.int32)
    lib().ref_compact_rays(n_alive, _p(ra), _p(ra_old), _p(rt), _p(rt_old), _p(cnt))
    return ra, rt, int(cnt[0])
