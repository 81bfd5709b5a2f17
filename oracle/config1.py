"""BASELINE.json configs[0]: volume composite of N rays x T samples (RGB + depth + C-class semantics) on synthetic
densities through the reference's PyTorch path on CPU -- TEST / BASELINE INFRASTRUCTURE ONLY.

Inputs follow SURVEY.md section 8(d) "Config 1": seed 1234; sigma = 50 U(0,1)^4, rgb = U(0,1), p = softmax(N(0,1)),
rays from the origin with normalised Gaussian directions, direction_norms = 1, bound 4, density_scale 1, no
perturbation, run(num_steps=T, upsample_steps=0).  `SyntheticHeads` plays the network: density / color / semantics
return those tensors (colour and semantics zeroed outside the mask, like the masked evaluation of
network_tcnn_semantics.py:147-207), so that run() (renderer_semantics.py:123-299, restated in live_path.run) times
the compositing alone.
"""
from __future__ import annotations

import time

import torch

from . import live_path


def make_inputs(n=4096, t=128, c=40, seed=1234):
    g = torch.Generator().manual_seed(seed)
    sigma = 50.0 * torch.rand(n * t, generator=g) ** 4
    rgb = torch.rand(n * t, 3, generator=g)
    prob = torch.softmax(torch.randn(n * t, c, generator=g), dim=-1)
    rays_o = torch.zeros(1, n, 3)
    rays_d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    dn = torch.ones(1, n, 1)
    return dict(sigma=sigma, rgb=rgb, prob=prob, rays_o=rays_o, rays_d=rays_d, direction_norms=dn, n=n, t=t, c=c)


class SyntheticHeads:
    """density / color / semantics of a 'network' that returns fixed per-sample tensors."""

    def __init__(self, inp, bound=4.0, requires_grad=False):
        self.bound = float(bound)
        self.num_semantic_classes = inp["c"]
        self.sigma = inp["sigma"].clone().requires_grad_(requires_grad)
        self.rgb = inp["rgb"].clone().requires_grad_(requires_grad)
        self.prob = inp["prob"].clone().requires_grad_(requires_grad)

    def density(self, x):
        return {"sigma": self.sigma, "geo_feat": torch.zeros(x.shape[0], 1)}

    def color(self, x, d, mask=None, geo_feat=None, **_):
        return self.rgb * mask.unsqueeze(1)

    def semantics(self, x, d, mask=None, geo_feat=None, **_):
        return self.prob * mask.unsqueeze(1)


def run_config1(inp, heads=None):
    heads = heads or SyntheticHeads(inp)
    return live_path.run(heads, inp["rays_o"], inp["rays_d"], inp["direction_norms"], num_steps=inp["t"],
                         upsample_steps=0, perturb=False)


def time_cpu(n=4096, t=128, c=40, threads=None, repeats=5, warmup=2, backward=True):
    """-> dict(fwd_ms, fwd_bwd_ms, cores): best of `repeats` after `warmup`, torch.set_num_threads(threads)."""
    import os

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = make_inputs(n, t, c)
    g = torch.Generator().manual_seed(1)
    gi, gd, gs = torch.randn(1, n, 3, generator=g), torch.randn(1, n, generator=g), torch.randn(1, n, c, generator=g)
    best_f, best_fb = float("inf"), float("inf")
    for it in range(warmup + repeats):
        heads = SyntheticHeads(inp, requires_grad=backward)
        t0 = time.perf_counter()
        out = run_config1(inp, heads)
        t1 = time.perf_counter()
        if backward:
            ((out["image"] * gi).sum() + (out["depth"] * gd).sum() + (out["semantics"] * gs).sum()).backward()
        t2 = time.perf_counter()
        if it >= warmup:
            best_f, best_fb = min(best_f, t1 - t0), min(best_fb, t2 - t0)
    return {"fwd_ms": best_f * 1e3, "fwd_bwd_ms": best_fb * 1e3, "cores": threads}
