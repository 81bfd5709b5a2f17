"""SemanticNeRFRenderer -- host-side mirror of nr4seg/nerf/renderer_semantics.py on libucsa_nerf.so.

Same constructor, buffers, ``run`` / ``render`` signatures, return dict and error behaviour as the reference
(renderer_semantics.py:61-358).  ``run`` here is the *generic* form: it calls the subclass' ``density`` /
``color`` / ``semantics`` exactly where the reference does (so any subclass written against the reference
keeps working) and executes the renderer's own arithmetic -- AABB test, sample placement, importance
resampling + merge, weights / masks and the three composites, forward and backward -- in CUDA kernels.
``SemanticNeRFNetwork`` overrides it with the fully fused pipeline.

Two documented liberties, both opt-in through ``**kwargs`` the reference signature already accepts:
``t_rand=[N,num_steps]`` and ``u=[N,upsample_steps]`` inject the uniform numbers the reference draws with
``torch.rand`` (renderer_semantics.py:166 and :28; the latter on the *CPU* generator even for inference).
Without them a counter-based generator keyed by (module seed, call counter, ray index, sample index) is used.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .raymarching import raymarching


def sample_pdf(bins, weights, n_samples, det=False):
    """Compatibility export of renderer_semantics.py:10-46 (inverse-CDF sampling) in torch ops.

    ``run`` does not call it -- resampling is fused into ``ucsa_resample_merge`` -- it is kept because the
    reference module exports it."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, device=weights.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples]).to(weights.device)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp_min(0)
    hi = hi.clamp_max(cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    return b_lo + (u - c_lo) / span * (b_hi - b_lo)


class _CompositeDense(torch.autograd.Function):
    """weights + masks + depth / image / semantics composites (renderer_semantics.py:238-285)."""

    @staticmethod
    def forward(ctx, sigma, z, rgb, prob, direction_norms, density_scale):
        sigma = sigma.detach().float().contiguous()
        z = z.detach().float().contiguous()
        rgb = rgb.detach().float().contiguous()
        prob = prob.detach().float().contiguous()
        dn = direction_norms.detach().float().contiguous()
        n, t = sigma.shape
        c = prob.shape[-1]
        dev = sigma.device
        weights = torch.empty(n, t, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        semantics = torch.empty(n, c, dtype=torch.float32, device=dev)
        ops.composite_dense_fwd(sigma, z, rgb, prob, dn, density_scale, weights, depth, image, semantics)
        ctx.save_for_backward(sigma, z, rgb, weights, dn)
        ctx.density_scale = density_scale
        ctx.n_classes = c
        return depth, image, semantics

    @staticmethod
    def backward(ctx, g_depth, g_image, g_semantics):
        sigma, z, rgb, weights, dn = ctx.saved_tensors
        n, t = sigma.shape
        dev = sigma.device
        d_sigma = torch.empty_like(sigma)
        d_rgb = torch.empty_like(rgb)
        d_prob = torch.empty(n, t, ctx.n_classes, dtype=torch.float32, device=dev)
        ops.composite_dense_bwd(sigma, z, rgb, weights, dn, g_depth.float().contiguous(),
                                g_image.float().contiguous(), g_semantics.float().contiguous(),
                                ctx.density_scale, d_sigma, d_rgb, d_prob)
        return d_sigma, None, d_rgb, d_prob, None, None


composite_dense = _CompositeDense.apply


class SemanticNeRFRenderer(nn.Module):

    def __init__(
        self,
        bound=1,
        cuda_ray=False,
        density_scale=1,
        num_semantic_classes=41,
    ):
        super().__init__()

        self.epoch = 1
        self.weights = np.zeros([0])
        self.weights_sum = np.zeros([0])

        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.density_scale = density_scale
        self.num_semantic_classes = num_semantic_classes

        # (xmin, ymin, zmin, xmax, ymax, zmax); only used to place samples (renderer_semantics.py:81-87)
        aabb_train = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        aabb_infer = aabb_train.clone()
        self.register_buffer("aabb_train", aabb_train)
        self.register_buffer("aabb_infer", aabb_infer)

        # extra state of the occupancy-grid path (renderer_semantics.py:89-103)
        self.cuda_ray = cuda_ray
        if cuda_ray:
            density_grid = torch.zeros([self.cascade] + [128] * 3)  # [CAS, H, H, H]
            self.register_buffer("density_grid", density_grid)
            self.mean_density = 0
            self.iter_density = 0
            step_counter = torch.zeros(16, 2, dtype=torch.int32)
            self.register_buffer("step_counter", step_counter)
            self.mean_count = 0
            self.local_step = 0

        # counter-based random numbers (see module docstring)
        self.rng_seed = 0x5EED
        self._rng_calls = 0

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    # (min, max) marching steps per wavefront round of run_cuda's inference loop.  torch-ngp uses (1, 8); measured on a
    # 640x480 view (scripts/occupancy_probe.py, B200): (1, 8) 132.5 ms, (2, 8) 88.2, (4, 8) 69.9, (8, 8) 61.2,
    # (8, 16) 60.6, (16, 32) 59.9 -- all with bit-identical images.
    wavefront_steps = (8, 16)

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_bitfield = None
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def _next_seed(self):
        self._rng_calls += 1
        return (self.rng_seed << 20) ^ self._rng_calls

    def run(self,
            rays_o,
            rays_d,
            direction_norms,
            num_steps=256,
            upsample_steps=256,
            bg_color=None,
            perturb=False,
            epoch=None,
            **kwargs):
        # rays_o, rays_d: [B, N, 3]; direction_norms: [B, N, 1]; returns image [B,N,3], depth [B,N], semantics [B,N,C]
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        direction_norms = direction_norms.contiguous().view(-1).float()
        n_rays = rays_o.shape[0]
        device = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        seed = kwargs.get("seed", None) or self._next_seed()
        ray_base = int(kwargs.get("ray_base", 0))
        t_rand, u = kwargs.get("t_rand"), kwargs.get("u")

        nears, fars = ops.near_far_from_aabb(rays_o, rays_d, aabb)
        t_total = num_steps + upsample_steps
        z_cat = torch.empty(n_rays, t_total, dtype=torch.float32, device=device)
        lin = torch.linspace(0.0, 1.0, num_steps, device=device)
        ops.sample_coarse(nears, fars, lin, z_cat, num_steps, perturb=perturb,
                          t_rand=None if t_rand is None else t_rand.float().contiguous(), seed=seed,
                          ray_base=ray_base)

        def positions(z):
            p = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
            return torch.min(torch.max(p, aabb[:3]), aabb[3:])

        z_vals = z_cat[:, :num_steps]
        xyzs = positions(z_vals)
        density_outputs = self.density(xyzs.reshape(-1, 3))
        for k, v in density_outputs.items():
            density_outputs[k] = v.view(n_rays, num_steps, -1)

        if upsample_steps > 0:
            with torch.no_grad():
                sigma_cat = torch.zeros(n_rays, t_total, dtype=torch.float32, device=device)
                sigma_cat[:, :num_steps] = density_outputs["sigma"].squeeze(-1).float()
                order = torch.empty(n_rays, t_total, dtype=torch.int32, device=device)
                ops.resample_merge(sigma_cat, z_cat, order, num_steps, upsample_steps, self.density_scale,
                                   u=None if u is None else u.float().contiguous(), seed=seed, ray_base=ray_base)
                new_xyzs = positions(z_cat[:, num_steps:])
            new_density_outputs = self.density(new_xyzs.reshape(-1, 3))
            for k, v in new_density_outputs.items():
                new_density_outputs[k] = v.view(n_rays, upsample_steps, -1)
            z_index = order.long()
            z_vals = torch.gather(z_cat, 1, z_index)
            xyzs = torch.cat([xyzs, new_xyzs], dim=1)
            xyzs = torch.gather(xyzs, 1, z_index.unsqueeze(-1).expand_as(xyzs))
            for k in density_outputs:
                tmp = torch.cat([density_outputs[k], new_density_outputs[k]], dim=1)
                density_outputs[k] = torch.gather(tmp, 1, z_index.unsqueeze(-1).expand_as(tmp))

        sigma = density_outputs["sigma"].squeeze(-1).float()
        with torch.no_grad():
            w_tmp = torch.empty_like(sigma)
            depth_tmp = torch.empty(n_rays, dtype=torch.float32, device=device)
            count_tmp = torch.empty(n_rays, dtype=torch.int32, device=device)
            mask_u8 = torch.empty(n_rays, t_total, dtype=torch.uint8, device=device)
            ops.weights_fwd(z_vals.contiguous(), sigma.detach().contiguous(), None, direction_norms,
                            self.density_scale, w_tmp, depth_tmp, count_tmp, mask_u8)
            mask = mask_u8.bool()  # weights > 1e-4 (renderer_semantics.py:249-250)

        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        for k, v in density_outputs.items():
            density_outputs[k] = v.reshape(-1, v.shape[-1])
        rgbs = self.color(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), mask=mask.reshape(-1), **density_outputs)
        rgbs = rgbs.view(n_rays, -1, 3)
        local_semantics = self.semantics(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), mask=mask.reshape(-1),
                                         **density_outputs)
        local_semantics = local_semantics.view(n_rays, -1, self.num_semantic_classes)

        depth, image, semantics = composite_dense(sigma, z_vals, rgbs, local_semantics, direction_norms,
                                                  float(self.density_scale))
        return {
            "depth": depth.view(*prefix),
            "image": image.view(*prefix, 3),
            "semantics": semantics.view(*prefix, self.num_semantic_classes),
        }

    # ------------------------------------------------------------------ occupancy-grid path (cuda_ray=True)
    # The reference allocates density_grid / step_counter (renderer_semantics.py:89-103) but ships neither run_cuda
    # nor update_extra_state; both are defined here after torch-ngp, which the reference is adapted
    # from (README.md:257): march through the occupancy grid, evaluate the heads on the packed samples, composite
    # with the ragged kernels.  Semantics are composited as probabilities with detached weights like the live path.
    def run_cuda(self, rays_o, rays_d, direction_norms, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False,
                 max_steps=1024, epoch=None, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        direction_norms = direction_norms.contiguous().view(-1).float()
        n_rays = rays_o.shape[0]
        device = rays_o.device
        c = self.num_semantic_classes
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = ops.near_far_from_aabb(rays_o, rays_d, aabb)
        bits = getattr(self, "density_bitfield", None)

        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_grid, self.mean_density, nears, fars, counter, self.mean_count,
                perturb, 128, force_all_rays, dt_gamma, bitfield=bits)
            # a subclass may offer heads + ragged compositing of the marched batch as one fused autograd node
            fused = getattr(self, "render_packed_train", None)
            fused = fused(xyzs, dirs, deltas, rays) if fused is not None else None
            if fused is not None:
                weights_sum, depth, image, semantics = fused
            else:
                sigmas, rgbs, sems = self(xyzs, dirs)
                sigmas = self.density_scale * sigmas
                weights_sum, depth, image, semantics = raymarching.composite_rays_train_semantics(
                    sigmas, rgbs.float(), sems.float(), deltas, rays, c)
        else:
            dtype = torch.float32
            weights_sum = torch.zeros(n_rays, dtype=dtype, device=device)
            depth = torch.zeros(n_rays, dtype=dtype, device=device)
            image = torch.zeros(n_rays, 3, dtype=dtype, device=device)
            semantics = torch.zeros(n_rays, c, dtype=dtype, device=device)
            n_alive = n_rays
            alive_counter = torch.zeros([1], dtype=torch.int32, device=device)
            rays_alive = torch.zeros(2, n_rays, dtype=torch.int32, device=device)  # 2 is used to loop old/new
            rays_t = torch.zeros(2, n_rays, dtype=dtype, device=device)
            step = 0
            i = 0
            while step < max_steps:
                if step == 0:
                    n_alive = n_rays
                    rays_alive[0] = torch.arange(n_alive, dtype=torch.int32, device=device)
                    rays_t[0] = nears
                else:
                    alive_counter.zero_()
                    raymarching.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2],
                                             rays_t[(i + 1) % 2], alive_counter)
                    n_alive = alive_counter.item()  # the wavefront's one host read per round
                if n_alive <= 0:
                    break
                # fewer live rays -> more steps per round (torch-ngp: between 1 and 8).  The rendered values do not depend
                # on the schedule -- every ray continues from its own t and the compositing stops a ray at T < 1e-4
                # wherever that falls inside a round; a ray wastes at most n_step - 1 evaluations, once -- so the lower
                # bound is a pure cost knob: rounds (launches + one host read each) against wasted samples.
                lo, hi = self.wavefront_steps
                n_step = max(min(n_rays // n_alive, hi), lo)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o,
                                                            rays_d, self.bound, self.density_grid, self.mean_density,
                                                            nears, fars, 128, perturb, dt_gamma, bitfield=bits)
                # a subclass may offer the three heads on packed points as one kernel chain (no autograd: inference),
                # handing over the semantic LOGITS: the soft-max is then taken inside the compositing kernel
                heads = getattr(self, "forward_packed", None) if not torch.is_grad_enabled() else None
                if heads is not None:
                    sigmas, rgbs, logits = heads(xyzs, dirs, want_logits=True)
                    raymarching.composite_rays_semantics(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2],
                                                         self.density_scale * sigmas, rgbs, None, deltas, weights_sum,
                                                         depth, image, semantics, logits=logits)
                else:
                    sigmas, rgbs, sems = self(xyzs, dirs)
                    sigmas = self.density_scale * sigmas
                    raymarching.composite_rays_semantics(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas,
                                                         rgbs.float(), sems.float(), deltas, weights_sum, depth, image,
                                                         semantics)
                step += n_step
                i += 1
        depth = depth / direction_norms
        return {
            "depth": depth.view(*prefix),
            "image": image.view(*prefix, 3),
            "semantics": semantics.view(*prefix, c),
        }

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """Refresh the occupancy grid from the current density field (call every ~16 training steps)."""
        if not self.cuda_ray:
            return
        device = self.density_grid.device
        h = self.density_grid.shape[1]
        fresh = -torch.ones_like(self.density_grid)
        coords = torch.arange(h, dtype=torch.int32, device=device)
        xs = coords.split(S)
        for xc in xs:
            for yc in xs:
                for zc in xs:
                    xx, yy, zz = torch.meshgrid(xc, yc, zc, indexing="ij")
                    cell = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)  # [n,3]
                    index = (cell[:, 0].long() * h + cell[:, 1].long()) * h + cell[:, 2].long()  # raymarching.cu:204
                    xyz = 2 * cell.float() / (h - 1) - 1
                    for cas in range(self.cascade):
                        cas_bound = min(2 ** cas, self.bound)
                        half = cas_bound / h
                        pts = xyz * (cas_bound - half) + (torch.rand_like(xyz) * 2 - 1) * half
                        sigma = self.density(pts)["sigma"].reshape(-1).detach() * self.density_scale
                        fresh[cas].view(-1)[index] = sigma
        ops.grid_update(self.density_grid, fresh.contiguous(), decay)
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        if not hasattr(self, "density_bitfield") or self.density_bitfield is None:
            self.density_bitfield = torch.zeros(self.density_grid.numel() // 32, dtype=torch.int32, device=device)
        ops.grid_packbits(self.density_grid, self.mean_density, self.density_bitfield)
        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0

    def render(self,
               rays_o,
               rays_d,
               direction_norms,
               staged=False,
               max_ray_batch=4096,
               bg_color=None,
               perturb=False,
               epoch=None,
               **kwargs):
        # rays_o, rays_d: [B, N, 3]; direction_norms: [B, N, 1]
        _run = self.run_cuda if self.cuda_ray else self.run
        B, N = rays_o.shape[:2]
        device = rays_o.device

        # never stage when cuda_ray (renderer_semantics.py:320-321)
        if staged and not self.cuda_ray:
            depth = torch.empty((B, N), device=device)
            image = torch.empty((B, N, 3), device=device)
            semantics = torch.empty((B, N, self.num_semantic_classes), device=device)
            t_rand, u = kwargs.pop("t_rand", None), kwargs.pop("u", None)
            seed = kwargs.pop("seed", None) or self._next_seed()
            # max_ray_batch bounds the reference's memory; a subclass whose run() is chunk-invariant (random numbers
            # keyed by ray index) may ask for larger chunks when no graph is recorded (self.stage_chunk)
            stage_chunk = getattr(self, "stage_chunk", None)
            if stage_chunk and not torch.is_grad_enabled():
                cap = max(int(max_ray_batch), int(stage_chunk))
                max_ray_batch = -(-N // -(-N // cap)) if N > 0 else cap  # equal parts: one workspace shape, no ragged tail
            for b in range(B):
                head = 0
                while head < N:
                    tail = min(head + max_ray_batch, N)
                    sl = slice(b * N + head, b * N + tail)
                    results_ = _run(rays_o[b:b + 1, head:tail],
                                    rays_d[b:b + 1, head:tail],
                                    direction_norms=direction_norms[b:b + 1, head:tail],
                                    bg_color=bg_color,
                                    perturb=perturb,
                                    epoch=epoch,
                                    t_rand=None if t_rand is None else t_rand[sl],
                                    u=None if u is None else u[sl],
                                    seed=seed,
                                    ray_base=b * N + head,
                                    **kwargs)
                    depth[b:b + 1, head:tail] = results_["depth"]
                    image[b:b + 1, head:tail] = results_["image"]
                    semantics[b:b + 1, head:tail] = results_["semantics"]
                    head += max_ray_batch
            results = {"depth": depth, "image": image, "semantics": semantics}
        else:
            results = _run(rays_o,
                           rays_d,
                           direction_norms=direction_norms,
                           bg_color=bg_color,
                           perturb=perturb,
                           epoch=epoch,
                           **kwargs)
        return results
