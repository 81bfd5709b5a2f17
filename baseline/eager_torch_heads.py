"""GPU STAND-IN for the reference's tiny-cuda-nn path ("Reference GPU (tiny-cuda-nn) path"
denominator) -- NOT tiny-cuda-nn, and never reported as such.

tiny-cuda-nn is not installable here (README.md:51 installs it from git HEAD; no network), so BASELINE.json's
"10x the reference GPU path" target has no measurable denominator.  What can be measured on the same B200 is the
reference-SHAPED step with the network heads written the way a PyTorch user would without tcnn: the hash-grid
encoding as index arithmetic + gathers, the three MLPs as fp16 cuBLAS matmuls, SH in elementwise ops, autograd for the
backward -- driven through the reference's call structure (`run()` calling `density` / `color(mask)` /
`semantics(mask)`, network_tcnn_semantics.py:130-207, renderer_semantics.py:123-299; the renderer's own arithmetic
is this repo's `SemanticNeRFRenderer.run` generic form).  bench.py reports it under the key `stand_in`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from ucsa_neural_rendering_b200 import ops
from ucsa_neural_rendering_b200.nerf.renderer_semantics import SemanticNeRFRenderer

PRIME_Y, PRIME_Z = 2654435761, 805459861


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


class EagerTorchNetwork(SemanticNeRFRenderer):
    """Same architecture and parameter counts as SemanticNeRFNetwork, heads in eager PyTorch (fp16 compute)."""

    def __init__(self, bound=4, num_semantic_classes=40, **kw):
        super().__init__(bound, cuda_ray=False, density_scale=1, num_semantic_classes=num_semantic_classes)
        g = ops.make_grid_desc(bound)
        self.levels = [(float(g.scale[l]), int(g.res[l]), int(g.entries[l]), int(g.offset[l]), bool(g.hashed[l]))
                       for l in range(16)]
        n = int(g.total_entries)
        self.table = torch.nn.Parameter((torch.rand(n, 2) * 2 - 1) * 1e-4)

        def mlp(dims):
            return torch.nn.ParameterList(
                [torch.nn.Parameter((torch.rand(o, i) * 2 - 1) * math.sqrt(6.0 / (i + o))) for i, o in zip(dims[:-1], dims[1:])])

        self.sigma_w = mlp([32, 64, 16])
        self.color_w = mlp([32, 64, 64, 16])
        self.sem_w = mlp([16, 64, 48])

    @staticmethod
    def _mlp(x, ws):
        h = x
        for i, w in enumerate(ws):
            h = F.linear(h, w.half())
            if i < len(ws) - 1:
                h = torch.relu(h)
        return h

    def _encode(self, x01):
        table = self.table.half()
        feats = []
        for scale, res, entries, offset, hashed in self.levels:
            pos = x01 * scale + 0.5
            cell = torch.floor(pos)
            frac = pos - cell
            cell = cell.long()
            acc = 0
            for c in range(8):
                cx = cell[:, 0] + (c & 1)
                cy = cell[:, 1] + ((c >> 1) & 1)
                cz = cell[:, 2] + ((c >> 2) & 1)
                if hashed:
                    idx = (cx ^ ((cy * PRIME_Y) & 0xFFFFFFFF) ^ ((cz * PRIME_Z) & 0xFFFFFFFF)) & (entries - 1)
                else:
                    idx = (cx + cy * res + cz * res * res) % entries
                w = torch.ones_like(frac[:, 0])
                for d in range(3):
                    w = w * (frac[:, d] if (c >> d) & 1 else 1 - frac[:, d])
                acc = acc + w[:, None] * table[idx + offset].float()
            feats.append(acc.half())
        return torch.cat(feats, dim=1)

    def density(self, x):
        x01 = (x + self.bound) / (2 * self.bound)
        h = self._mlp(self._encode(x01), self.sigma_w)
        return {"sigma": _TruncExp.apply(h[:, 0]), "geo_feat": h[:, 1:]}

    @staticmethod
    def _sh4(d):
        x, y, z = d[:, 0], d[:, 1], d[:, 2]
        xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
        return torch.stack([
            torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
            -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
            0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
            0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
            2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
            0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
            1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], dim=1).half()

    def color(self, x, d, mask=None, geo_feat=None, **_):
        rgbs = torch.zeros(mask.shape[0], 3, dtype=torch.float32, device=x.device)
        if not mask.any():
            return rgbs
        d, geo = d[mask], geo_feat[mask]
        ones = torch.ones(d.shape[0], 1, dtype=torch.float16, device=x.device)
        h = self._mlp(torch.cat([self._sh4(d), geo, ones], dim=-1), self.color_w)[:, :3]
        rgbs[mask] = torch.sigmoid(h).float()
        return rgbs

    def semantics(self, x, d, mask=None, geo_feat=None, **_):
        c = self.num_semantic_classes
        sem = torch.zeros(mask.shape[0], c, dtype=torch.float32, device=x.device)
        if not mask.any():
            return sem
        geo = geo_feat[mask]
        ones = torch.ones(geo.shape[0], 1, dtype=torch.float16, device=x.device)
        h = self._mlp(torch.cat([geo, ones], dim=-1), self.sem_w)[:, :c]
        sem[mask] = F.softmax(h.float(), dim=-1)
        return sem
