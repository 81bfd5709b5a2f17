"""torch-CPU restatement of the reference's LIVE rendering path -- TEST INFRASTRUCTURE ONLY.

Follows, function by function:

* ``near_far``        nr4seg/nerf/raymarching/src/raymarching.cu:62-115 (slab test,
                      min_near default 0.2 from raymarching/raymarching.py:16)
* ``sample_pdf``      nr4seg/nerf/renderer_semantics.py:10-46
* ``OracleHeads``     nr4seg/nerf/network_tcnn_semantics.py:102-207 with the tcnn calls
                      replaced by oracle/tcnn_spec.py and trunc_exp from activation.py:7-19
* ``run`` / ``render`` renderer_semantics.py:123-299 / :301-358

Pinned: tests/golden/make_golden.py imports the *reference* renderer and network module
in the build container (tinycudann stubbed by tcnn_spec, trimesh stubbed) and freezes
its outputs and gradients; tests/test_oracle_golden.py checks this file against them.

The only liberty taken is that the two random draws of the reference
(``torch.rand`` for the stratified jitter, renderer_semantics.py:166, and for the
inverse-CDF samples, :28) can be injected, so that CPU and GPU see the same numbers.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import tcnn_spec as spec

FLT_MAX = torch.finfo(torch.float32).max


def near_far(rays_o: torch.Tensor, rays_d: torch.Tensor, aabb: torch.Tensor, min_near: float = 0.2):
    """Ray / axis-aligned box slab intersection in float32; misses give FLT_MAX twice."""
    o = rays_o.float()
    inv = 1.0 / rays_d.float()
    t_lo = (aabb[:3].float() - o) * inv
    t_hi = (aabb[3:].float() - o) * inv
    swap = t_lo > t_hi  # the kernel's "if (near > far) swap", NaN-faithful
    t_in = torch.where(swap, t_hi, t_lo)
    t_out = torch.where(swap, t_lo, t_hi)
    near = t_in[:, 0]
    far = t_out[:, 0]
    miss = torch.zeros_like(near, dtype=torch.bool)
    for ax in (1, 2):
        miss = miss | (near > t_out[:, ax]) | (t_in[:, ax] > far)
        near = torch.where(t_in[:, ax] > near, t_in[:, ax], near)
        far = torch.where(t_out[:, ax] < far, t_out[:, ax], far)
    near = torch.where(near < min_near, torch.full_like(near, min_near), near)
    near = torch.where(miss, torch.full_like(near, FLT_MAX), near)
    far = torch.where(miss, torch.full_like(far, FLT_MAX), far)
    return near, far


def transmittance_weights(z: torch.Tensor, sigma: torch.Tensor, density_scale: float):
    """renderer_semantics.py:185-198 and :238-247 (same arithmetic twice in the reference)."""
    gaps = z[:, 1:] - z[:, :-1]
    gaps = torch.cat([gaps, torch.full_like(gaps[:, :1], 1e10)], dim=1)
    alpha = 1 - torch.exp(-gaps * density_scale * sigma)
    keep = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-15], dim=1)
    trans = torch.cumprod(keep, dim=1)[:, :-1]
    return alpha * trans, gaps


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, u: torch.Tensor | None = None):
    """Inverse-CDF sampling; ``u`` [B,n_samples] in [0,1) replaces the reference's torch.rand."""
    w = weights + 1e-5
    pdf = w / w.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)
    if u is None:
        u = torch.rand(cdf.shape[0], n_samples)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp_min(0)
    hi = hi.clamp_max(cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    return b_lo + (u - c_lo) / span * (b_hi - b_lo)


class _TruncExp(torch.autograd.Function):
    """activation.py:7-19: exp forward in fp32, backward g * exp(clamp(x, -15, 15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


class OracleHeads(torch.nn.Module):
    """density / color / semantics / forward of SemanticNeRFNetwork on the tcnn spec."""

    def __init__(self, bound: float = 4, num_semantic_classes: int = 40, geo_feat_dim: int = 15,
                 seed: int = 1337, hash_amp: float = 1e-4):
        super().__init__()
        self.bound = float(bound)
        self.num_semantic_classes = num_semantic_classes
        self.geo_feat_dim = geo_feat_dim
        self.table = spec.level_table(bound)
        self.dims_sigma = spec.mlp_dims(32, 1 + geo_feat_dim, 1)
        self.dims_color = spec.mlp_dims(16 + geo_feat_dim, 3, 2)
        self.dims_sem = spec.mlp_dims(geo_feat_dim, num_semantic_classes, 1)
        n_hash = self.table["total"] * spec.N_FEATURES
        self.encoder = torch.nn.Parameter(spec.splitmix_uniform(n_hash, seed, -hash_amp, hash_amp))
        self.sigma_net = torch.nn.Parameter(spec.xavier_mlp_init(self.dims_sigma, seed + 1))
        self.color_net = torch.nn.Parameter(spec.xavier_mlp_init(self.dims_color, seed + 2))
        self.semantics_net = torch.nn.Parameter(spec.xavier_mlp_init(self.dims_sem, seed + 3))

    # network_tcnn_semantics.py:130-144
    def density(self, x):
        x01 = (x + self.bound) / (2 * self.bound)
        enc = spec.hashgrid_forward(x01, self.encoder, self.table)
        h = spec.mlp_forward(enc, self.sigma_net, self.dims_sigma, 32, 1 + self.geo_feat_dim)
        return {"sigma": _TruncExp.apply(h[:, 0]), "geo_feat": h[:, 1:]}

    def _color_rows(self, d, geo_feat):
        sh = spec.sh4_forward((d + 1) / 2)
        h = spec.mlp_forward(torch.cat([sh, geo_feat], dim=-1), self.color_net, self.dims_color,
                             16 + self.geo_feat_dim, 3)
        return spec._round_h(torch.sigmoid(h))  # fp16 sigmoid output under autocast

    # network_tcnn_semantics.py:147-178
    def color(self, x, d, mask=None, geo_feat=None, **_):
        if mask is None:
            return self._color_rows(d, geo_feat)
        out = torch.zeros(mask.shape[0], 3, dtype=torch.float32)
        if not mask.any():
            return out
        rows = self._color_rows(d[mask], geo_feat[mask])
        return out.masked_scatter(mask.unsqueeze(1).expand_as(out), rows)

    def _sem_rows(self, geo_feat):
        h = spec.mlp_forward(geo_feat, self.semantics_net, self.dims_sem, self.geo_feat_dim,
                             self.num_semantic_classes)
        return F.softmax(h, dim=-1)

    # network_tcnn_semantics.py:180-207
    def semantics(self, x, d, mask=None, geo_feat=None, **_):
        if mask is None:
            return self._sem_rows(geo_feat)
        out = torch.zeros(mask.shape[0], self.num_semantic_classes, dtype=torch.float32)
        if not mask.any():
            return out
        rows = self._sem_rows(geo_feat[mask])
        return out.masked_scatter(mask.unsqueeze(1).expand_as(out), rows)

    # network_tcnn_semantics.py:102-128
    def forward(self, x, d):
        dens = self.density(x)
        return dens["sigma"], self._color_rows(d, dens["geo_feat"]), self._sem_rows(dens["geo_feat"])


def run(heads, rays_o, rays_d, direction_norms, *, aabb=None, num_steps=256, upsample_steps=256,
        perturb=False, density_scale=1.0, t_rand=None, u=None, return_aux=False):
    """One un-staged pass; tensors are [B,N,...] like the reference, B folded into N."""
    prefix = rays_o.shape[:-1]
    o = rays_o.contiguous().view(-1, 3).float()
    d = rays_d.contiguous().view(-1, 3).float()
    dn = direction_norms.contiguous().view(-1).float()
    n_rays = o.shape[0]
    if aabb is None:
        b = heads.bound
        aabb = torch.tensor([-b, -b, -b, b, b, b], dtype=torch.float32)

    near, far = near_far(o, d, aabb)
    near, far = near.unsqueeze(1), far.unsqueeze(1)
    z = near + (far - near) * torch.linspace(0.0, 1.0, num_steps).unsqueeze(0)
    if perturb:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        top = torch.cat([mid, z[:, -1:]], dim=1)
        bot = torch.cat([z[:, :1], mid], dim=1)
        if t_rand is None:
            t_rand = torch.rand(z.shape)
        z = bot + (top - bot) * t_rand

    def positions(zz):
        p = o.unsqueeze(1) + d.unsqueeze(1) * zz.unsqueeze(2)
        return torch.min(torch.max(p, aabb[:3]), aabb[3:])

    xyz = positions(z)
    dens = heads.density(xyz.reshape(-1, 3))
    sigma = dens["sigma"].view(n_rays, num_steps)
    geo = dens["geo_feat"].view(n_rays, num_steps, -1)
    aux = {}

    if upsample_steps > 0:
        with torch.no_grad():
            w_coarse, gaps = transmittance_weights(z, sigma, density_scale)
            z_mid = z[:, :-1] + 0.5 * gaps[:, :-1]
            z_new = sample_pdf(z_mid, w_coarse[:, 1:-1], upsample_steps, u=u).detach()
            xyz_new = positions(z_new)
        dens_new = heads.density(xyz_new.reshape(-1, 3))
        z, order = torch.sort(torch.cat([z, z_new], dim=1), dim=1)
        pick = lambda a, b: torch.gather(torch.cat([a, b], dim=1), 1,
                                         order.view(n_rays, -1, *([1] * (a.dim() - 2))).expand(-1, -1, *a.shape[2:]))
        xyz = pick(xyz, xyz_new)
        sigma = pick(sigma, dens_new["sigma"].view(n_rays, upsample_steps))
        geo = pick(geo, dens_new["geo_feat"].view(n_rays, upsample_steps, -1))
        aux.update(z_new=z_new, order=order)

    weights, _ = transmittance_weights(z, sigma, density_scale)
    mask = weights > 1e-4
    n_tot = z.shape[1]
    dirs = d.view(-1, 1, 3).expand(-1, n_tot, -1).reshape(-1, 3)
    flat = dict(geo_feat=geo.reshape(-1, geo.shape[-1]))
    rgb = heads.color(xyz.reshape(-1, 3), dirs, mask=mask.reshape(-1), **flat).view(n_rays, n_tot, 3)
    prob = heads.semantics(xyz.reshape(-1, 3), dirs, mask=mask.reshape(-1), **flat)
    prob = prob.view(n_rays, n_tot, -1)

    w_sem = torch.where(mask, weights.detach(), torch.zeros_like(weights))
    w_rgb = torch.where(mask, weights, torch.zeros_like(weights))
    depth = (w_rgb * z).sum(dim=-1) / dn
    image = (w_rgb.unsqueeze(-1) * rgb).sum(dim=-2)
    semantics = (w_sem.unsqueeze(-1) * prob).sum(dim=-2)
    out = {
        "depth": depth.view(*prefix),
        "image": image.view(*prefix, 3),
        "semantics": semantics.view(*prefix, prob.shape[-1]),
    }
    if return_aux:
        aux.update(z=z, weights=weights, mask=mask, sigma=sigma)
        out["aux"] = aux
    return out


def render(heads, rays_o, rays_d, direction_norms, *, staged=False, max_ray_batch=4096, **kw):
    """renderer_semantics.py:301-358: optional sequential chunking over rays."""
    if not staged:
        return run(heads, rays_o, rays_d, direction_norms, **kw)
    b, n = rays_o.shape[:2]
    c = heads.num_semantic_classes
    depth = torch.empty(b, n)
    image = torch.empty(b, n, 3)
    sem = torch.empty(b, n, c)
    t_rand, u = kw.pop("t_rand", None), kw.pop("u", None)
    for bi in range(b):
        for head in range(0, n, max_ray_batch):
            tail = min(head + max_ray_batch, n)
            sl = slice(bi * n + head, bi * n + tail)
            part = run(heads, rays_o[bi:bi + 1, head:tail], rays_d[bi:bi + 1, head:tail],
                       direction_norms[bi:bi + 1, head:tail],
                       t_rand=None if t_rand is None else t_rand[sl],
                       u=None if u is None else u[sl], **kw)
            depth[bi:bi + 1, head:tail] = part["depth"]
            image[bi:bi + 1, head:tail] = part["image"]
            sem[bi:bi + 1, head:tail] = part["semantics"]
    return {"depth": depth, "image": image, "semantics": sem}
