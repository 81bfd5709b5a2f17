"""The fused LIVE-path pipeline as two kernel chains over an explicit workspace.

``forward_chain`` is SemanticNeRFRenderer.run() (renderer_semantics.py:123-299) with the network heads inlined;
``backward_chain`` is its gradient.  Neither allocates, synchronises or touches the host, so a chain can be replayed
from a CUDA graph.  ``SemanticNeRFNetwork.run`` drives them through one autograd node (fresh workspace per call);
``engine.TrainEngine`` drives them directly on a static workspace.

Kernel sequence, forward:  near_far -> sample_coarse -> density (coarse) -> resample_merge -> density (fine)
                           -> weights -> scan -> compact -> heads + compositing
               backward:   heads + compositing backward -> weights backward -> density backward (hash scatter)
"""
from __future__ import annotations

import weakref

import torch

from . import ops


class RenderWorkspace:
    """Every buffer of one run() call for N rays, Tc + Tf samples, C classes."""

    def __init__(self, n, tc, tf, c, device, need_grad):
        t = tc + tf
        f32 = dict(dtype=torch.float32, device=device)
        f16 = dict(dtype=torch.float16, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.n, self.tc, self.tf, self.t, self.c, self.need_grad = n, tc, tf, t, c, need_grad
        self.tiled = ops.density_tiled(tc, tf)  # enc / hid in tile layout (bulk copies) when the passes align
        self.k_max = n * t  # worst case; the kernels read the true K from ray_off[n] on the device (no host sync)
        self.lin = torch.linspace(0.0, 1.0, tc, device=device)
        self.nears = torch.empty(n, **f32)
        self.fars = torch.empty(n, **f32)
        self.z_cat = torch.empty(n, t, **f32)
        self.sigma = torch.empty(n, t, **f32)
        self.h = torch.empty(n, t, 16, **f16)
        self.order = torch.empty(n, t, **i32) if tf > 0 else None
        self.w_sorted = torch.empty(n, t, **f32)
        self.scan_scratch = ops.weights_scratch(n, device)
        self.use_geo = torch.empty(n, t, dtype=torch.uint8, device=device)
        self.ray_off = torch.empty(n + 1, **i32)
        self.sel = torch.empty(self.k_max, **i32)
        self.w_sel = torch.empty(self.k_max, **f32)
        self.z_sel = torch.empty(self.k_max, **f32)
        self.rgb = torch.empty(self.k_max, 3, **f32)
        self.depth = torch.empty(n, **f32)
        self.image = torch.empty(n, 3, **f32)
        self.semantics = torch.empty(n, c, **f32)
        if need_grad:
            self.enc = torch.empty(n, t, 32, **f16)
            self.hid = torch.empty(n, t, 64, **f16)
            rows = ops.tile_rows(self.k_max)  # tile-layout buffers (ucsa_nerf.h: ucsa_heads_fwd)
            self.hc1 = torch.empty(rows, 64, **f16)
            self.hc2 = torch.empty(rows, 64, **f16)
            self.hs = torch.empty(rows, 64, **f16)
            self.d_w_sel = torch.empty(self.k_max, **f32)
            self.dh = torch.empty(n, t, 16, **f16)
            self.d_sigma = torch.empty(n, t, **f32)
        else:
            self.enc = self.hid = self.hc1 = self.hc2 = self.hs = None
        self._token = None  # weak reference to the autograd node that still needs this workspace (cached_workspace)

    def fresh_outputs(self):
        """New depth / image / semantics tensors (the ones a caller keeps); everything else is scratch."""
        f32 = dict(dtype=torch.float32, device=self.z_cat.device)
        self.depth = torch.empty(self.n, **f32)
        self.image = torch.empty(self.n, 3, **f32)
        self.semantics = torch.empty(self.n, self.c, **f32)


# Workspaces of the autograd path, per network and shape: render() is called with the same (rays, samples) shape every
# step, and a workspace is ~25 allocations / ~1 GB at 4096 rays.  A workspace whose backward is still pending (its
# autograd node is alive) is never handed out again; a second one is allocated instead.
_WS_CACHE = weakref.WeakKeyDictionary()
_WS_CACHE_MAX = 4


class WorkspaceToken:
    """Lifetime marker held by the autograd node that owns a cached workspace."""


def cached_workspace(net, n, tc, tf, c, device, need_grad):
    """-> (workspace, token or None).  The caller keeps `token` alive for as long as it needs the scratch contents
    (i.e. until its backward has run); outputs are always fresh tensors."""
    per_net = _WS_CACHE.setdefault(net, {})
    key = (n, tc, tf, c, str(device), bool(need_grad))
    ws = per_net.get(key)
    if ws is None or (ws._token is not None and ws._token() is not None):
        fresh = RenderWorkspace(n, tc, tf, c, device, need_grad)
        if ws is None:
            while len(per_net) >= _WS_CACHE_MAX:
                per_net.pop(next(iter(per_net)))
            per_net[key] = fresh
        ws = fresh
    else:
        ws.fresh_outputs()
    token = None
    if need_grad:
        token = WorkspaceToken()
        ws._token = weakref.ref(token)
    return ws, token


def forward_chain(net, ws, rays_o, rays_d, dnorm, aabb, *, perturb, t_rand=None, u=None, seed=0, ray_base=0,
                  step_dev=None):
    n, tc, tf, t, c = ws.n, ws.tc, ws.tf, ws.t, ws.c
    grid = net.encoder.grid
    table_h = net.encoder.half_params()
    w_sig = net.sigma_net.half_params()
    ops.near_far_from_aabb_into(rays_o, rays_d, aabb, ws.nears, ws.fars)
    ops.sample_coarse(ws.nears, ws.fars, ws.lin, ws.z_cat, tc, perturb=perturb, t_rand=t_rand, seed=seed,
                      ray_base=ray_base, step_dev=step_dev)
    common = dict(rays_o=rays_o, rays_d=rays_d, aabb=aabb, z_cat=ws.z_cat, sigma=ws.sigma, h=ws.h, enc=ws.enc,
                  hid=ws.hid, tiled=ws.tiled and ws.enc is not None)
    ops.density_fwd(grid, table_h, w_sig, net.bound, k0=0, k1=tc, **common)
    if tf > 0:
        ops.resample_merge(ws.sigma, ws.z_cat, ws.order, tc, tf, net.density_scale, u=u, seed=seed, ray_base=ray_base,
                           step_dev=step_dev)
        ops.density_fwd(grid, table_h, w_sig, net.bound, k0=tc, k1=t, **common)
    ops.weights_compact(ws.z_cat, ws.sigma, ws.order, dnorm, net.density_scale, ws.w_sorted, ws.depth, ws.ray_off,
                        ws.use_geo, ws.sel, ws.w_sel, ws.z_sel, ws.scan_scratch)
    ws.image.zero_()
    ws.semantics.zero_()
    ops.heads_fwd(ws.sel, ws.ray_off, n, t, ws.k_max, rays_d, ws.h, net.color_net.half_params(),
                  net.semantics_net.half_params(), c, ws.rgb, None, ws.hc1, ws.hc2, ws.hs, w_sel=ws.w_sel,
                  image=ws.image, semantics=ws.semantics)  # heads + compositing in one kernel


def backward_chain(net, ws, rays_o, rays_d, dnorm, aabb, g_image, g_depth, g_sem, grad_table, grad_sigma, grad_color,
                   grad_sem):
    """Accumulates into the four gradient buffers (fp32; the caller zero-fills them)."""
    n, t, c = ws.n, ws.t, ws.c
    scale = float(net.loss_scale)
    ops.heads_bwd(ws.sel, ws.ray_off, n, t, ws.k_max, rays_d, ws.h, net.color_net.half_params(),
                  net.semantics_net.half_params(), c, ws.rgb, ws.hc1, ws.hc2, ws.hs, ws.w_sel, ws.z_sel,
                  g_image, g_depth, g_sem, dnorm, scale, ws.dh, ws.d_w_sel, grad_color, grad_sem)
    ops.weights_bwd(ws.z_cat, ws.sigma, ws.order, ws.w_sorted, ws.ray_off, ws.d_w_sel, net.density_scale, ws.d_sigma)
    ops.density_bwd(net.encoder.grid, net.sigma_net.half_params(), net.bound, rays_o=rays_o, rays_d=rays_d, aabb=aabb,
                    z_cat=ws.z_cat, k0=0, k1=t, h=ws.h, enc=ws.enc, hid=ws.hid, d_sigma=ws.d_sigma, dh=ws.dh,
                    use_geo=ws.use_geo, loss_scale=scale, grad_table=grad_table, grad_w_sigma=grad_sigma,
                    replicas=net.encoder.grad_replicas(), tiled=ws.tiled)
