"""Frozen CPU restatement of the tiny-cuda-nn arithmetic on the reference's hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: tiny-cuda-nn
is installed by the reference from git HEAD (README.md:51) and is not vendored, so
there is no tcnn build or golden vector to pin against.  What is restated here is
the published tcnn / Instant-NGP algorithm, anchored on the reference's call sites:

* HashGrid     -- nr4seg/nerf/network_tcnn_semantics.py:34-46,108,134
                  (16 levels, F=2, 2^19 entries, base 16, per_level_scale from bound)
* FullyFusedMLP-- network_tcnn_semantics.py:48-58 (sigma 32->64->16),
                  :74-84 (colour 31->64->64->3), :90-100 (semantics 15->64->C)
* SH degree 4  -- network_tcnn_semantics.py:64-70,117,165

Spec decisions frozen here (DESIGN.md section "tcnn arithmetic spec"):

* level scale   s_l = float32(exp2(l * log2f(pls)) * 16 - 1) (exp2 in double, one rounding), res_l = ceil(s_l)+1,
                entries_l = min(round_up(res_l^3, 8), 2^19); dense index
                x + y*res + z*res^2 (mod entries_l) when res^3 <= entries_l, otherwise
                (x*1) ^ (y*2654435761) ^ (z*805459861) (uint32) mod entries_l.
* position      pos = fma(s_l, x01, 0.5); cell = floor(pos); frac = pos - cell.
* interpolation fp32 accumulate of w_c * half(table[idx_c]) over the 8 corners in corner
                order c = 0..7 (bit d of c selects +1 on axis d); one rounding to fp16.
* MLP           weights fp16 (cast from the fp32 master), y = W x with W stored
                [out, in] row-major, layer after layer in one flat buffer; fp32
                accumulate; ReLU then rounding to fp16 between layers; no bias; input
                padded to a multiple of 16 with the constant 1.0 (tcnn's identity
                encoding padding), output padded to a multiple of 16 and sliced.
* SH4           the 16 real SH basis polynomials of (2*d01 - 1), fp32 -> fp16.
"""
from __future__ import annotations

import math

import numpy as np
import torch

N_LEVELS = 16
N_FEATURES = 2
LOG2_HASHMAP = 19
BASE_RES = 16
PRIME_Y = np.uint32(2654435761)
PRIME_Z = np.uint32(805459861)


def per_level_scale(bound: float) -> float:
    """network_tcnn_semantics.py:34."""
    return float(np.exp2(np.log2(2048 * bound / 16) / (16 - 1)))


def level_table(bound: float):
    """Per-level (scale f32, res, entries, offset, hashed) for the reference config."""
    pls = np.float32(per_level_scale(bound))
    log2_pls = np.float32(np.log2(pls))
    scales, res, entries, offsets, hashed = [], [], [], [], []
    off = 0
    for lvl in range(N_LEVELS):
        s = np.float32(np.exp2(np.float64(lvl) * np.float64(log2_pls)) * np.float64(BASE_RES) - np.float64(1.0))
        r = int(math.ceil(float(s))) + 1
        n = min(((r ** 3 + 7) // 8) * 8, 1 << LOG2_HASHMAP)
        scales.append(s)
        res.append(r)
        entries.append(n)
        offsets.append(off)
        hashed.append(r ** 3 > n)
        off += n
    return {
        "scale": np.asarray(scales, dtype=np.float32),
        "res": np.asarray(res, dtype=np.uint32),
        "entries": np.asarray(entries, dtype=np.uint32),
        "offset": np.asarray(offsets, dtype=np.uint32),
        "hashed": np.asarray(hashed, dtype=bool),
        "total": off,
    }


def _fma32(a: np.ndarray, b: np.ndarray, c) -> np.ndarray:
    """float32 fused multiply-add emulated through float64 (product exact in f64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(np.float32)


def grid_cells(x01: np.ndarray, table, level: int):
    """Integer cell, fractional part of one level.  x01: [S,3] float32 in [0,1]."""
    s = np.float32(table["scale"][level])
    pos = _fma32(np.full_like(x01, s), x01, 0.5)
    cell = np.floor(pos)
    frac = (pos - cell).astype(np.float32)
    return cell.astype(np.uint32), frac


def corner_index(cell: np.ndarray, table, level: int, corner: int) -> np.ndarray:
    """Index (within the level) of one of the 8 corners; uint32 wrap-around arithmetic."""
    cx = cell[:, 0] + np.uint32(corner & 1)
    cy = cell[:, 1] + np.uint32((corner >> 1) & 1)
    cz = cell[:, 2] + np.uint32((corner >> 2) & 1)
    n = np.uint32(table["entries"][level])
    if table["hashed"][level]:
        with np.errstate(over="ignore"):
            idx = cx ^ (cy * PRIME_Y) ^ (cz * PRIME_Z)
    else:
        r = np.uint32(table["res"][level])
        with np.errstate(over="ignore"):
            idx = cx + cy * r + cz * r * r
    return (idx % n).astype(np.uint32)


def corner_weight(frac: np.ndarray, corner: int) -> np.ndarray:
    one = np.float32(1.0)
    w = np.ones(frac.shape[0], dtype=np.float32)
    for d in range(3):
        f = frac[:, d]
        w = (w * (f if (corner >> d) & 1 else (one - f))).astype(np.float32)
    return w


def hash_indices(x01: np.ndarray, table) -> np.ndarray:
    """[S,16,8] uint32 global entry index (level offset included) -- the bit-exact contract."""
    x01 = np.ascontiguousarray(x01, dtype=np.float32)
    out = np.empty((x01.shape[0], N_LEVELS, 8), dtype=np.uint32)
    for lvl in range(N_LEVELS):
        cell, _ = grid_cells(x01, table, lvl)
        for c in range(8):
            out[:, lvl, c] = corner_index(cell, table, lvl, c) + np.uint32(table["offset"][lvl])
    return out


class _RoundH(torch.autograd.Function):
    """Round to fp16 and come back to fp32.  Straight-through: the gradient passes unchanged in fp32 (a plain
    ``t.half().float()`` would round the *gradient* to fp16 at the cast, which is an artefact of the emulation)."""

    @staticmethod
    def forward(ctx, t):
        return t.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _round_h(t: torch.Tensor) -> torch.Tensor:
    return _RoundH.apply(t)


class _HashGridFn(torch.autograd.Function):
    """All-torch gather / scatter of the grid (no numpy round-trips): per level one [S,8] index block, one gather and
    -- backward -- one index_add_ into the level's slice of the gradient.  Same arithmetic and summation order
    (corner 0..7, fp32) as the numpy helpers above, which stay the bit-exact index contract."""

    @staticmethod
    def forward(ctx, x01, params, table):
        xs = x01.detach().to(torch.float32).cpu()
        tab_h = params.detach().half().float().view(-1, N_FEATURES)
        s = xs.shape[0]
        out = torch.empty(s, N_LEVELS * N_FEATURES, dtype=torch.float32)
        saved_idx, saved_w = [], []
        m32 = 0xFFFFFFFF
        for lvl in range(N_LEVELS):
            scale = float(table["scale"][lvl])
            pos = (xs.double() * scale + 0.5).float()  # float32 fma: the product is exact in float64
            cell_f = torch.floor(pos)
            frac = pos - cell_f
            cell = cell_f.long() & m32
            n = int(table["entries"][lvl])
            r = int(table["res"][lvl])
            idx = torch.empty(s, 8, dtype=torch.int64)
            w = torch.empty(s, 8, dtype=torch.float32)
            for c in range(8):
                cx = (cell[:, 0] + (c & 1)) & m32
                cy = (cell[:, 1] + ((c >> 1) & 1)) & m32
                cz = (cell[:, 2] + ((c >> 2) & 1)) & m32
                if table["hashed"][lvl]:
                    i = cx ^ ((cy * int(PRIME_Y)) & m32) ^ ((cz * int(PRIME_Z)) & m32)
                else:
                    i = (cx + ((cy * r) & m32) + ((cz * ((r * r) & m32)) & m32)) & m32
                idx[:, c] = i % n + int(table["offset"][lvl])
                wc = torch.ones(s, dtype=torch.float32)
                for d in range(3):
                    f = frac[:, d]
                    wc = wc * (f if (c >> d) & 1 else (1.0 - f))
                w[:, c] = wc
            vals = tab_h[idx.view(-1)].view(s, 8, N_FEATURES)
            acc = torch.zeros(s, N_FEATURES, dtype=torch.float32)
            for c in range(8):
                acc = acc + w[:, c:c + 1] * vals[:, c]
            out[:, lvl * N_FEATURES:(lvl + 1) * N_FEATURES] = acc
            saved_idx.append(idx)
            saved_w.append(w)
        ctx.saved_idx, ctx.saved_w, ctx.n_params = saved_idx, saved_w, params.numel()
        return out.half().float()

    @staticmethod
    def backward(ctx, g):
        grad = torch.zeros(ctx.n_params // N_FEATURES, N_FEATURES, dtype=torch.float32)
        for lvl in range(N_LEVELS):
            gl = g[:, lvl * N_FEATURES:(lvl + 1) * N_FEATURES]
            contrib = (ctx.saved_w[lvl].unsqueeze(2) * gl.unsqueeze(1)).reshape(-1, N_FEATURES)
            grad.index_add_(0, ctx.saved_idx[lvl].view(-1), contrib)
        return None, grad.view(-1), None


def hashgrid_forward(x01: torch.Tensor, params: torch.Tensor, table) -> torch.Tensor:
    """x01 [S,3] f32 in [0,1]; params [total*2] f32 master.  Returns [S,32] f32 holding
    fp16-representable values (level-major feature order l0f0,l0f1,l1f0,...).  Both roundings (table -> fp16,
    output -> fp16) are straight-through for the gradient, like _round_h."""
    return _HashGridFn.apply(x01, params, table)


def sh4_forward(d01: torch.Tensor) -> torch.Tensor:
    """[K,3] f32 in [0,1] -> [K,16] f32 holding fp16-representable values."""
    v = d01.float() * 2.0 - 1.0
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    cols = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return _round_h(torch.stack(cols, dim=1))


def pad16(n: int) -> int:
    return ((n + 15) // 16) * 16


def mlp_dims(n_in: int, n_out: int, n_hidden_layers: int, width: int = 64):
    return [pad16(n_in)] + [width] * n_hidden_layers + [pad16(n_out)]


def mlp_param_count(dims) -> int:
    return sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def mlp_forward(x: torch.Tensor, params: torch.Tensor, dims, n_in: int, n_out: int) -> torch.Tensor:
    """x [S,n_in] (fp16-representable); returns [S,n_out] f32 holding fp16 values."""
    s = x.shape[0]
    h = _round_h(x.float())
    if dims[0] > n_in:
        h = torch.cat([h, torch.ones(s, dims[0] - n_in, dtype=torch.float32)], dim=1)
    off = 0
    n_layers = len(dims) - 1
    for i in range(n_layers):
        fi, fo = dims[i], dims[i + 1]
        w = _round_h(params[off:off + fi * fo]).view(fo, fi)
        off += fi * fo
        h = h @ w.t()
        if i < n_layers - 1:
            h = torch.relu(h)
        h = _round_h(h)
    return h[:, :n_out]


def splitmix_uniform(n: int, seed: int, lo: float, hi: float) -> torch.Tensor:
    """Portable counter-based U(lo,hi): splitmix64 finaliser of (index, seed); float32."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = i + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32))


def xavier_mlp_init(dims, seed: int) -> torch.Tensor:
    parts = []
    for i, (fi, fo) in enumerate(zip(dims[:-1], dims[1:])):
        s = math.sqrt(6.0 / (fi + fo))
        parts.append(splitmix_uniform(fi * fo, seed * 16 + i, -s, s))
    return torch.cat(parts)
