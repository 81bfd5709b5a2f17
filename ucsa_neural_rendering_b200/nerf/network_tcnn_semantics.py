"""SemanticNeRFNetwork -- drop-in for nr4seg/nerf/network_tcnn_semantics.py without tiny-cuda-nn.

Same constructor, attributes (``encoder``, ``sigma_net``, ``encoder_dir``, ``color_net``, ``semantics_net``,
each an ``nn.Module`` with ``.parameters()`` for the two Adam groups of joint_train_lightning_net.py:897-919)
and methods (``forward``, ``density``, ``color``, ``semantics``, ``run``, ``render``) as the reference
(network_tcnn_semantics.py:10-207).  ``run`` is overridden with the fused pipeline: one autograd node whose
forward and backward are chains of libucsa_nerf.so kernels with no eager tensor math in between.

Parameters are fp32 masters (like the tcnn torch bindings keep them); kernels read an fp16 working copy that
is refreshed whenever the master's version counter moves (i.e. after every optimizer step).
"""
from __future__ import annotations

import os

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops, pipeline
from .activation import trunc_exp  # noqa: F401  (re-exported like the reference module)
from .renderer_semantics import SemanticNeRFRenderer


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class _HalfCache(nn.Module):
    """fp32 master ``params`` + lazily refreshed fp16 copy."""

    def __init__(self):
        super().__init__()
        self._half = None
        self._half_key = None

    def half_params(self) -> torch.Tensor:
        p = self.params
        key = (p.data_ptr(), p._version, p.device)
        if self._half is None or self._half_key != key or self._half.device != p.device:
            if self._half is None or self._half.shape != p.shape or self._half.device != p.device:
                self._half = torch.empty(p.shape, dtype=torch.float16, device=p.device)
            ops.cast_f32_to_f16(p.detach(), self._half)
            self._half_key = key
        return self._half


class _HashGridFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x01, params, module):
        x01 = x01.detach().float().contiguous()
        enc = torch.empty(x01.shape[0], 32, dtype=torch.float16, device=x01.device)
        ops.hashgrid_fwd(x01, module.half_params(), module.grid, enc)
        ctx.save_for_backward(x01)
        ctx.module = module
        return enc

    @staticmethod
    def backward(ctx, d_enc):
        (x01,) = ctx.saved_tensors
        m = ctx.module
        scale = m.loss_scale
        grad = torch.zeros_like(m.params)
        ops.hashgrid_bwd(x01, m.grid, (d_enc.float() * scale).half().contiguous(), 1.0 / scale, grad)
        return None, grad, None


class HashGridEncoding(_HalfCache):
    """tcnn.Encoding(HashGrid) of network_tcnn_semantics.py:36-46: [S,3] in [0,1] -> [S,32] fp16."""

    def __init__(self, bound, seed=1337, loss_scale=128.0):
        super().__init__()
        self.n_input_dims = 3
        self.n_output_dims = 32
        self.loss_scale = loss_scale
        self.grid = ops.make_grid_desc(bound)
        n = int(self.grid.total_entries) * 2
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * 1e-4)  # tcnn: U(-1e-4, 1e-4)

    def forward(self, x):
        return _HashGridFn.apply(x, self.params, self)

    def grad_replicas(self, n_replicas=None):
        """Scratch for the backward scatter of the dense levels (ops.density_bwd); allocated once, kept zero."""
        if n_replicas is None:
            # 0 since the backward scatter merges same-cell runs per warp (grid.cuh: scatter_level_runs): the few coarse
            # cells next to the camera no longer serialise the L2 atomic units.  Kept as a knob for other ray sets.
            n_replicas = int(os.environ.get("UCSA_GRAD_REPLICAS", "0"))
        if n_replicas == 0:
            return None
        rep = getattr(self, "_grad_replicas", None)
        if rep is None or rep.device != self.params.device or rep.shape[0] != n_replicas:
            rep = ops.grad_replicas(self.grid, self.params.device, n_replicas)
            self._grad_replicas = rep
        return rep


class SHEncoding(nn.Module):
    """tcnn.Encoding(SphericalHarmonics, degree 4) of network_tcnn_semantics.py:64-70: [K,3] in [0,1] -> [K,16] fp16."""

    def __init__(self):
        super().__init__()
        self.n_input_dims = 3
        self.n_output_dims = 16

    def forward(self, d01):
        d01 = d01.detach().float().contiguous()
        out = torch.empty(d01.shape[0], 16, dtype=torch.float16, device=d01.device)
        ops.sh4_fwd(d01, out)
        return out


class _MlpFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, params, module):
        dims = module.dims
        n = x.shape[0]
        xh = x.detach().half()
        if xh.shape[1] < dims[0]:  # tcnn pads the input with ones: the extra column acts as a bias
            xh = torch.cat([xh, torch.ones(n, dims[0] - xh.shape[1], dtype=torch.float16, device=x.device)], 1)
        xh = xh.contiguous()
        y = torch.empty(n, dims[-1], dtype=torch.float16, device=x.device)
        acts = torch.empty(n, sum(dims[1:-1]), dtype=torch.float16, device=x.device)
        ops.mlp_fwd(xh, module.half_params(), dims, y, acts)
        ctx.save_for_backward(xh, acts)
        ctx.module = module
        ctx.n_in = x.shape[1]
        ctx.x_dtype = x.dtype
        return y[:, :module.n_output_dims]

    @staticmethod
    def backward(ctx, dy):
        xh, acts = ctx.saved_tensors
        m = ctx.module
        dims = m.dims
        n = xh.shape[0]
        scale = m.loss_scale
        dyh = torch.zeros(n, dims[-1], dtype=torch.float16, device=xh.device)
        dyh[:, :m.n_output_dims] = (dy.float() * scale).half()
        dx = torch.empty(n, dims[0], dtype=torch.float16, device=xh.device)
        grad = torch.zeros_like(m.params)
        ops.mlp_bwd(xh, m.half_params(), dims, acts, dyh, 1.0 / scale, dx, grad)
        return (dx[:, :ctx.n_in].float() / scale).to(ctx.x_dtype), grad, None


class FusedMLP(_HalfCache):
    """tcnn.Network(FullyFusedMLP, ReLU, no output activation): fp16 in/out, fp32 accumulate, no bias."""

    def __init__(self, n_input_dims, n_output_dims, n_hidden_layers, n_neurons=64, seed=0, out_pad=None,
                 loss_scale=128.0):
        super().__init__()
        assert n_neurons == 64, "FullyFusedMLP is built for 64-wide hidden layers"
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.loss_scale = loss_scale
        self.dims = [_pad16(n_input_dims)] + [n_neurons] * n_hidden_layers + [out_pad or _pad16(n_output_dims)]
        g = torch.Generator().manual_seed(seed)
        parts = []
        for fi, fo in zip(self.dims[:-1], self.dims[1:]):
            s = math.sqrt(6.0 / (fi + fo))  # tcnn default: Xavier-uniform
            parts.append((torch.rand(fo * fi, generator=g) * 2 - 1) * s)
        self.params = nn.Parameter(torch.cat(parts))

    def forward(self, x):
        return _MlpFn.apply(x, self.params, self)


class _DensityFn(torch.autograd.Function):
    """density(x) for free-standing points: encode + sigma MLP + trunc_exp in one kernel each way."""

    @staticmethod
    def forward(ctx, xyz, enc_params, sigma_params, net, grad_enabled=True):
        xyz = xyz.detach().float().contiguous()
        s = xyz.shape[0]
        dev = xyz.device
        # needs_input_grad mirrors requires_grad only (True under no_grad as well); the caller passes its grad mode
        need = grad_enabled and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        sigma = torch.empty(s, dtype=torch.float32, device=dev)
        h = torch.empty(s, 16, dtype=torch.float16, device=dev)
        rows = ops.tile_rows(s)  # enc / hid: tile-layout buffers, whole 128-row tiles
        enc = torch.empty(rows, 32, dtype=torch.float16, device=dev) if need else None
        hid = torch.empty(rows, 64, dtype=torch.float16, device=dev) if need else None
        ops.density_fwd(net.encoder.grid, net.encoder.half_params(), net.sigma_net.half_params(), net.bound,
                        xyz=xyz, sigma=sigma, h=h, enc=enc, hid=hid, tiled=need)
        ctx.net = net
        if need:
            ctx.save_for_backward(xyz, h, enc, hid)
        geo = h[:, 1:]
        return sigma, geo

    @staticmethod
    def backward(ctx, d_sigma, d_geo):
        net = ctx.net
        xyz, h, enc, hid = ctx.saved_tensors
        s = xyz.shape[0]
        dev = xyz.device
        scale = net.loss_scale
        dh = torch.zeros(s, 16, dtype=torch.float16, device=dev)
        use_geo = None
        if d_geo is not None:
            dh[:, 1:] = (d_geo.float() * scale).half()
            use_geo = torch.ones(s, dtype=torch.uint8, device=dev)
        g_table = torch.zeros_like(net.encoder.params)
        g_sigma = torch.zeros_like(net.sigma_net.params)
        ops.density_bwd(net.encoder.grid, net.sigma_net.half_params(), net.bound, xyz=xyz, h=h, enc=enc, hid=hid,
                        d_sigma=None if d_sigma is None else d_sigma.float().contiguous(), dh=dh, use_geo=use_geo,
                        loss_scale=scale, grad_table=g_table, grad_w_sigma=g_sigma, tiled=True)
        return None, g_table, g_sigma, None, None


class _FusedRender(torch.autograd.Function):
    """The whole of SemanticNeRFRenderer.run() (renderer_semantics.py:123-299) with the network heads inlined: one
    autograd node around pipeline.forward_chain / backward_chain."""

    @staticmethod
    def forward(ctx, enc_params, sigma_params, color_params, sem_params, rays_o, rays_d, dnorm, net, cfg):
        # ctx.needs_input_grad only mirrors requires_grad of the inputs (it is True under torch.no_grad() as well) and
        # grad mode is off inside forward, so the caller's grad mode comes in through cfg: without it, inference would
        # allocate and write every saved activation (480 B per sample).
        need = cfg["grad_enabled"] and any(ctx.needs_input_grad[:4])
        ws, token = pipeline.cached_workspace(net, rays_o.shape[0], cfg["num_steps"], cfg["upsample_steps"],
                                              net.num_semantic_classes, rays_o.device, need)
        pipeline.forward_chain(net, ws, rays_o, rays_d, dnorm, cfg["aabb"], perturb=cfg["perturb"],
                               t_rand=cfg["t_rand"], u=cfg["u"], seed=cfg["seed"], ray_base=cfg["ray_base"])
        if need:
            ctx.net, ctx.ws, ctx.aabb, ctx.token = net, ws, cfg["aabb"], token  # token: the workspace stays ours
            ctx.save_for_backward(rays_o, rays_d, dnorm)
        return ws.depth, ws.image, ws.semantics

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_depth, g_image, g_sem):
        net, ws = ctx.net, ctx.ws
        rays_o, rays_d, dnorm = ctx.saved_tensors
        if ws is None:
            raise RuntimeError("the fused render node was already differentiated (its workspace has been released)")
        n_table = net.encoder.params.numel()
        sizes = (n_table, ops.SIGMA_PARAMS, ops.COLOR_PARAMS, ops.SEM_PARAMS)
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=rays_o.device)  # one allocation, one memset
        g_table, g_sig, g_col, g_semw = torch.split(flat, sizes)
        pipeline.backward_chain(net, ws, rays_o, rays_d, dnorm, ctx.aabb, g_image.float().contiguous(),
                                g_depth.float().contiguous(), g_sem.float().contiguous(), g_table, g_sig, g_col, g_semw)
        ctx.ws = ctx.token = None  # the workspace may be handed out again
        return g_table, g_sig, g_col, g_semw, None, None, None, None, None


def _packed_heads_forward(net, x, d, need):
    """density + both heads on a packed stream of points x, d [M,3] (f32, contiguous): the fused density kernel, then
    the two tensor-core heads kernels on ALL points (identity row list, one direction per point).  `need`: keep the
    activations the backward kernels read.  -> dict of sigma [M] f32, rgb [M,3] f32, logits [M,48] f16 + saved buffers."""
    m = x.shape[0]
    dev = x.device
    rows = ops.tile_rows(m)
    f16 = dict(dtype=torch.float16, device=dev)
    st = {"sigma": torch.empty(m, dtype=torch.float32, device=dev), "rgb": torch.empty(m, 3, dtype=torch.float32, device=dev),
          "h": torch.empty(m, 16, **f16), "logits": torch.empty(m, ops.MAX_CLASSES, **f16),
          "enc": torch.empty(rows, 32, **f16) if need else None, "hid": torch.empty(rows, 64, **f16) if need else None,
          "hc1": torch.empty(rows, 64, **f16) if need else None, "hc2": torch.empty(rows, 64, **f16) if need else None,
          "hs": torch.empty(rows, 64, **f16) if need else None,
          "sel": torch.arange(m, dtype=torch.int32, device=dev)}
    # ray_off[n_rays] = K is all the heads kernels read of it: [0, m] serves the forward pass (one "ray" owning all
    # rows, no compositing), the slot at index m the backward pass (m "rays" of one row each)
    off = torch.zeros(m + 1, dtype=torch.int32, device=dev)
    off[1].fill_(m)
    off[m].fill_(m)
    st["off"] = off
    ops.density_fwd(net.encoder.grid, net.encoder.half_params(), net.sigma_net.half_params(), net.bound, xyz=x,
                    sigma=st["sigma"], h=st["h"], enc=st["enc"], hid=st["hid"], tiled=need)
    ops.heads_fwd(st["sel"], off, 1, 1, m, d, st["h"], net.color_net.half_params(), net.semantics_net.half_params(),
                  net.num_semantic_classes, st["rgb"], st["logits"], hc1=st["hc1"], hc2=st["hc2"], hs=st["hs"])
    return st


_PACKED_SAVED = ("h", "enc", "hid", "rgb", "hc1", "hc2", "hs", "sel", "off")


def _packed_heads_backward(net, x, d, saved, d_sigma, d_rgb, d_prob):
    """Backward of _packed_heads_forward from per-point gradients d_sigma [M], d_rgb [M,3], d_prob [M,C] (f32; d_prob is
    the gradient of the soft-max OUTPUT).  The heads kernels are the ones of the LIVE path; they take their per-row
    gradients through the fused compositing backward, so every point is presented as a ray of its own with weight 1:
    dL/drgb_row = 1 * g_image[row], dL/dlogits_row = soft-max backward of 1 * g_semantics[row].
    -> the four parameter gradients (views of one flat buffer)."""
    h, enc, hid, rgb, hc1, hc2, hs, sel, off = saved
    m = x.shape[0]
    dev = x.device
    c = net.num_semantic_classes
    f32 = dict(dtype=torch.float32, device=dev)
    scale = net.loss_scale
    sizes = (net.encoder.params.numel(), ops.SIGMA_PARAMS, ops.COLOR_PARAMS, ops.SEM_PARAMS)
    flat = torch.zeros(sum(sizes), **f32)
    g_table, g_sig, g_col, g_semw = torch.split(flat, sizes)
    ones, zeros = torch.ones(m, **f32), torch.zeros(m, **f32)
    dh = torch.zeros(m, 16, dtype=torch.float16, device=dev)
    d_w = torch.empty(m, **f32)  # dL/dw of the unit weights: not used
    ops.heads_bwd(sel, off, m, 1, m, d, h, net.color_net.half_params(), net.semantics_net.half_params(), c, rgb,
                  hc1, hc2, hs, ones, zeros, d_rgb, zeros, d_prob, ones, scale, dh, d_w, g_col, g_semw)
    ops.density_bwd(net.encoder.grid, net.sigma_net.half_params(), net.bound, xyz=x, h=h, enc=enc, hid=hid,
                    d_sigma=d_sigma, dh=dh, use_geo=torch.ones(m, dtype=torch.uint8, device=dev), loss_scale=scale,
                    grad_table=g_table, grad_w_sigma=g_sig, tiled=True)
    return g_table, g_sig, g_col, g_semw


class _PackedHeadsFn(torch.autograd.Function):
    """forward(x, d) of the network (network_tcnn_semantics.py:102-128) for a packed stream of points as ONE autograd
    node: the fused density kernel and the two tensor-core heads kernels each way instead of nine module-level
    launches with eager glue between them."""

    @staticmethod
    def forward(ctx, x, d, enc_params, sigma_params, color_params, sem_params, net, grad_enabled):
        x = x.detach().float().contiguous()
        d = d.detach().float().contiguous()
        need = grad_enabled and any(ctx.needs_input_grad[2:6])
        st = _packed_heads_forward(net, x, d, need)
        prob = F.softmax(st["logits"][:, :net.num_semantic_classes].float(), dim=-1)
        if need:
            ctx.net = net
            ctx.save_for_backward(x, d, *(st[k] for k in _PACKED_SAVED))
        return st["sigma"], st["rgb"], prob

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_sigma, d_rgb, d_prob):
        x, d, *saved = ctx.saved_tensors
        grads = _packed_heads_backward(ctx.net, x, d, saved, d_sigma.float().contiguous(), d_rgb.float().contiguous(),
                                       d_prob.float().contiguous())
        return (None, None, *grads, None, None)


class _PackedRenderFn(torch.autograd.Function):
    """The training render of the occupancy-grid path (run_cuda) behind the march as ONE autograd node: heads on the
    packed samples (_packed_heads_forward) and the ragged compositing of raymarching.cu:318-487 + the semantic channels,
    which takes the fp16 logits directly (soft-max inside the kernel: no [M,C] fp32 probability tensor, no eager
    soft-max).  Backward: ragged compositing backward -> per-sample gradients -> heads -> density.  Depth receives no
    gradient, like in the reference (raymarching.py:209)."""

    @staticmethod
    def forward(ctx, enc_params, sigma_params, color_params, sem_params, xyzs, dirs, deltas, rays, net, grad_enabled):
        x = xyzs.detach().float().contiguous()
        d = dirs.detach().float().contiguous()
        deltas = deltas.detach().float().contiguous()
        rays = rays.contiguous()
        n = rays.shape[0]
        dev = x.device
        c = net.num_semantic_classes
        need = grad_enabled and any(ctx.needs_input_grad[:4])
        st = _packed_heads_forward(net, x, d, need)
        sig = st["sigma"] if net.density_scale == 1 else st["sigma"] * net.density_scale
        f32 = dict(dtype=torch.float32, device=dev)
        ws, depth = torch.empty(n, **f32), torch.empty(n, **f32)
        image, semantics = torch.empty(n, 3, **f32), torch.empty(n, c, **f32)
        ops.composite_rays_train_forward(sig, st["rgb"], None, deltas, rays, c, ws, depth, image, semantics,
                                         logits=st["logits"])
        if need:
            ctx.net = net
            ctx.save_for_backward(x, d, deltas, rays, sig, ws, image, *(st[k] for k in _PACKED_SAVED))
        ctx.mark_non_differentiable(depth)
        return ws, depth, image, semantics

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_ws, g_depth, g_image, g_sem):
        net = ctx.net
        x, d, deltas, rays, sig, ws, image, *saved = ctx.saved_tensors
        m = x.shape[0]
        c = net.num_semantic_classes
        f32 = dict(dtype=torch.float32, device=x.device)
        # zero-filled: rows past the last sample (the stream is padded to whole tiles) and dropped rays stay 0
        d_sigma, d_rgb, d_prob = torch.zeros(m, **f32), torch.zeros(m, 3, **f32), torch.zeros(m, c, **f32)
        ops.composite_rays_train_backward(g_ws.float().contiguous(), g_image.float().contiguous(),
                                          g_sem.float().contiguous(), sig, saved[3], deltas, rays, ws, image, c, d_sigma,
                                          d_rgb, d_prob)
        if net.density_scale != 1:
            d_sigma = d_sigma * net.density_scale
        grads = _packed_heads_backward(net, x, d, saved, d_sigma, d_rgb, d_prob)
        return (*grads, None, None, None, None, None, None)


class SemanticNeRFNetwork(SemanticNeRFRenderer):

    def __init__(self,
                 encoding="HashGrid",
                 encoding_dir="SphericalHarmonics",
                 num_layers=2,
                 hidden_dim=64,
                 geo_feat_dim=15,
                 num_layers_color=3,
                 hidden_dim_color=64,
                 num_layers_semantics=2,
                 hidden_dim_semantics=64,
                 bound=1,
                 num_semantic_classes=41,
                 **kwargs):
        super().__init__(bound, **kwargs, num_semantic_classes=num_semantic_classes)
        if (num_layers, hidden_dim, geo_feat_dim, num_layers_color, hidden_dim_color, num_layers_semantics,
                hidden_dim_semantics) != (2, 64, 15, 3, 64, 2, 64):
            raise NotImplementedError(
                "the fused kernels are built for the reference architecture: sigma 32-64-16, colour 32-64-64-3, "
                "semantics 16-64-C (network_tcnn_semantics.py:12-24 defaults)")
        if not 1 <= num_semantic_classes <= ops.MAX_CLASSES:
            raise NotImplementedError(f"num_semantic_classes must be in [1, {ops.MAX_CLASSES}]")

        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.loss_scale = 128.0  # fp16 backward scale inside the kernels (tcnn's default for fp16)

        self.encoder = HashGridEncoding(bound, seed=1337, loss_scale=self.loss_scale)
        self.sigma_net = FusedMLP(32, 1 + geo_feat_dim, num_layers - 1, hidden_dim, seed=1338)

        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir = SHEncoding()
        self.in_dim_color = self.encoder_dir.n_output_dims + self.geo_feat_dim
        self.color_net = FusedMLP(self.in_dim_color, 3, num_layers_color - 1, hidden_dim_color, seed=1339)

        self.num_layers_semantics = num_layers_semantics
        self.hidden_dim_semantics = hidden_dim_semantics
        self.num_semantic_classes = num_semantic_classes
        self.in_dim_semantics = self.geo_feat_dim
        self.semantics_net = FusedMLP(self.in_dim_semantics, num_semantic_classes, num_layers_semantics - 1,
                                      hidden_dim_semantics, seed=1340, out_pad=ops.MAX_CLASSES)
        # render(staged=True) without gradients re-chunks to equal parts of at most this many rays whatever
        # max_ray_batch says (the caller's default of 4096 means 75 launches of every kernel per 640x480 frame): random
        # numbers are keyed by the global ray index, so the result does not depend on the chunking.  Inference saves no
        # activations, so a part of 2^18 rays x 512 samples needs ~9 GB of scratch (a 640x480 view = two parts of
        # 153600 rays, 5.6 GB; 46.3 ms per view against 47.3 ms with 65536-ray parts).
        self.stage_chunk = 262144
        self.fused_packed = True  # run_cuda's training render: heads as one fused autograd node (_PackedHeadsFn)

    # ------------------------------------------------------------------ module-level API of the reference
    def forward(self, x, d):
        # x: [N, 3] in [-bound, bound]; d: [N, 3] normalised
        dens = self.density(x)
        sigma, geo_feat = dens["sigma"], dens["geo_feat"]
        d = (d + 1) / 2
        d = self.encoder_dir(d)
        h = torch.cat([d, geo_feat], dim=-1)
        h = self.color_net(h)
        color = torch.sigmoid(h)
        semantics = self.semantics_net(geo_feat)
        semantics = F.softmax(semantics.float(), dim=-1)
        return sigma, color, semantics

    def density(self, x):
        sigma, geo_feat = _DensityFn.apply(x, self.encoder.params, self.sigma_net.params, self,
                                           torch.is_grad_enabled())
        return {"sigma": sigma, "geo_feat": geo_feat}

    @torch.no_grad()
    def forward_packed(self, x, d, want_logits=False):
        """forward(x, d) for a packed stream of points without autograd (the inference wavefront of run_cuda): the fused
        density kernel, then the two tensor-core heads kernels on ALL points (identity row list, one direction per
        point), then the soft-max -- 3 kernels + 1 torch op instead of the 9 launches of the module-by-module forward.
        -> sigma [M] f32, colour [M,3] f32, class probabilities [M,C] f32 (want_logits: the fp16 logits [M,48])"""
        x = x.detach().float().contiguous()
        d = d.detach().float().contiguous()
        m = x.shape[0]
        dev = x.device
        c = self.num_semantic_classes
        sigma = torch.empty(m, dtype=torch.float32, device=dev)
        rgb = torch.empty(m, 3, dtype=torch.float32, device=dev)
        if m == 0:
            empty = torch.empty(0, ops.MAX_CLASSES, dtype=torch.float16, device=dev)
            return sigma, rgb, (empty if want_logits else empty[:, :c].float())
        st = getattr(self, "_packed_state", None)
        if st is None or st["sel"].device != dev or st["sel"].numel() < m:
            cap = max(m, 1 << 16)
            st = {"sel": torch.arange(cap, dtype=torch.int32, device=dev),
                  "off": torch.zeros(2, dtype=torch.int32, device=dev)}
            self._packed_state = st
        st["off"][1] = m  # one "ray" that owns all rows: the heads kernels read K from the device
        h = torch.empty(m, 16, dtype=torch.float16, device=dev)
        logits = torch.empty(m, ops.MAX_CLASSES, dtype=torch.float16, device=dev)
        ops.density_fwd(self.encoder.grid, self.encoder.half_params(), self.sigma_net.half_params(), self.bound, xyz=x,
                        sigma=sigma, h=h)
        ops.heads_fwd(st["sel"], st["off"], 1, 1, m, d, h, self.color_net.half_params(),
                      self.semantics_net.half_params(), c, rgb, logits)
        return sigma, rgb, (logits if want_logits else F.softmax(logits[:, :c].float(), dim=-1))

    def forward_packed_train(self, x, d):
        """forward(x, d) for a packed stream of points WITH autograd (the training render of run_cuda) as one fused node
        (_PackedHeadsFn): -> sigma [M] f32, colour [M,3] f32, class probabilities [M,C] f32.  `fused_packed = False`
        on the network restores the module-by-module forward."""
        if x.shape[0] == 0 or not self.fused_packed:
            sigma, color, sem = self(x, d)
            return sigma, color.float(), sem
        return _PackedHeadsFn.apply(x, d, self.encoder.params, self.sigma_net.params, self.color_net.params,
                                    self.semantics_net.params, self, torch.is_grad_enabled())

    def render_packed_train(self, xyzs, dirs, deltas, rays):
        """Heads + ragged compositing of a marched training batch (run_cuda) as one fused autograd node
        (_PackedRenderFn) -> weights_sum [N], depth [N], image [N,3], semantics [N,C]; None when the fused form does
        not apply (`fused_packed = False`, no samples): the caller then composes the pieces itself."""
        if xyzs.shape[0] == 0 or not self.fused_packed:
            return None
        return _PackedRenderFn.apply(self.encoder.params, self.sigma_net.params, self.color_net.params,
                                     self.semantics_net.params, xyzs, dirs, deltas, rays, self,
                                     torch.is_grad_enabled())

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        # masked evaluation, network_tcnn_semantics.py:147-178
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=torch.float32, device=x.device)
            if not mask.any():
                return rgbs
            d = d[mask]
            geo_feat = geo_feat[mask]
        d = (d + 1) / 2
        d = self.encoder_dir(d)
        h = torch.cat([d, geo_feat.to(d.dtype)], dim=-1)
        h = self.color_net(h)
        h = torch.sigmoid(h)
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
        else:
            rgbs = h
        return rgbs

    def semantics(self, x, d, mask=None, geo_feat=None, **kwargs):
        # masked evaluation + softmax, network_tcnn_semantics.py:180-207
        if mask is not None:
            semantics = torch.zeros(mask.shape[0], self.num_semantic_classes, dtype=torch.float32, device=x.device)
            if not mask.any():
                return semantics
            geo_feat = geo_feat[mask]
        h = self.semantics_net(geo_feat)
        if mask is not None:
            semantics[mask] = F.softmax(h.to(semantics.dtype), dim=-1)
        else:
            semantics = F.softmax(h.float(), dim=-1)
        return semantics

    # ------------------------------------------------------------------ occupancy grid (cuda_ray=True)
    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """Refresh the occupancy grid from the current density field (call every ~16 training steps): three launches
        -- ucsa_grid_density (one jittered point per cell through the fused encode + sigma-MLP kernel, points generated
        on the fly), ucsa_grid_update_pack (EMA-max + mean, then the bitfield) -- and ONE host read (mean_density is a
        Python float in the reference's interface).  The generic eager form lives in SemanticNeRFRenderer."""
        if not self.cuda_ray:
            return
        grid = self.density_grid
        dev = grid.device
        h = grid.shape[1]
        st = getattr(self, "_grid_state", None)
        if st is None or st["fresh"].device != dev or st["fresh"].numel() != grid.numel():
            st = {"fresh": torch.empty(grid.numel(), dtype=torch.float32, device=dev),
                  "sum": torch.zeros(1, dtype=torch.float64, device=dev),
                  "mean": torch.zeros(1, dtype=torch.float32, device=dev)}
            self._grid_state = st
        if getattr(self, "density_bitfield", None) is None or self.density_bitfield.device != dev:
            self.density_bitfield = torch.zeros(grid.numel() // 32, dtype=torch.int32, device=dev)
        ops.grid_density(self.encoder.grid, self.encoder.half_params(), self.sigma_net.half_params(), self.bound,
                         self.cascade, h, self._next_seed(), st["fresh"])
        ops.grid_update_pack(grid, st["fresh"], decay, float(self.density_scale), st["sum"], st["mean"],
                             self.density_bitfield)
        self.mean_density = float(st["mean"])  # the one host read
        self.iter_density += 1
        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0

    # ------------------------------------------------------------------ fused rendering
    def run(self, rays_o, rays_d, direction_norms, num_steps=256, upsample_steps=256, bg_color=None,
            perturb=False, epoch=None, **kwargs):
        if kwargs.get("generic", False):
            kwargs.pop("generic")
            return super().run(rays_o, rays_d, direction_norms, num_steps=num_steps,
                               upsample_steps=upsample_steps, bg_color=bg_color, perturb=perturb, epoch=epoch,
                               **kwargs)
        prefix = rays_o.shape[:-1]
        o = rays_o.contiguous().view(-1, 3).float().contiguous()
        d = rays_d.contiguous().view(-1, 3).float().contiguous()
        dn = direction_norms.contiguous().view(-1).float().contiguous()
        t_rand, u = kwargs.get("t_rand"), kwargs.get("u")
        cfg = dict(
            num_steps=int(num_steps), upsample_steps=int(upsample_steps), perturb=bool(perturb),
            t_rand=None if t_rand is None else t_rand.float().contiguous(),
            u=None if u is None else u.float().contiguous(),
            seed=kwargs.get("seed", None) or self._next_seed(), ray_base=int(kwargs.get("ray_base", 0)),
            aabb=self.aabb_train if self.training else self.aabb_infer, grad_enabled=torch.is_grad_enabled())
        depth, image, semantics = _FusedRender.apply(self.encoder.params, self.sigma_net.params,
                                                     self.color_net.params, self.semantics_net.params, o, d, dn,
                                                     self, cfg)
        return {
            "depth": depth.view(*prefix),
            "image": image.view(*prefix, 3),
            "semantics": semantics.view(*prefix, self.num_semantic_classes),
        }
