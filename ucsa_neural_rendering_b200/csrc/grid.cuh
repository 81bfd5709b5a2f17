// Multiresolution hash-grid device functions (rows a4/a5 of SURVEY.md section 8).
// Arithmetic contract (frozen in oracle/tcnn_spec.py, DESIGN.md "tcnn arithmetic spec"):
//   pos = fma(scale_l, x01, 0.5); cell = floor(pos); frac = pos - cell;
//   corner c adds bit d of c on axis d; weight = ((1*wx)*wy)*wz in fp32;
//   index = dense x + y*res + z*res^2  or  x ^ y*2654435761 ^ z*805459861, then mod entries_l.
#pragma once

#include "common.cuh"

namespace ucsa {

struct LevelGeom {
  float scale;
  uint32_t res;
  uint32_t entries;
  uint32_t offset;
  uint32_t hashed;
};

__device__ __forceinline__ LevelGeom level_geom(const ucsa_grid_desc& g, int l) {
  return LevelGeom{g.scale[l], g.res[l], g.entries[l], g.offset[l], g.hashed[l]};
}

// number of entries of the leading dense levels (= offset of the first hashed level)
__host__ __device__ __forceinline__ uint32_t dense_entry_count(const ucsa_grid_desc& g) {
  for (int l = 0; l < UCSA_GRID_LEVELS; ++l)
    if (g.hashed[l]) return g.offset[l];
  return g.total_entries;
}

// Host-side check every entry point that takes a descriptor runs first: the kernels reduce a hash modulo the level
// size with a mask, and address 16-byte groups of entries.
inline bool grid_desc_ok(const ucsa_grid_desc* g) {
  if (g == nullptr) return false;
  for (int l = 0; l < UCSA_GRID_LEVELS; ++l) {
    if (g->entries[l] == 0u || g->entries[l] % 8u != 0u || g->offset[l] % 8u != 0u || g->res[l] < 2u) return false;
    if (g->hashed[l] && (g->entries[l] & (g->entries[l] - 1u)) != 0u) return false;
  }
  return true;
}
#define UCSA_REQUIRE_GRID(g, who) \
  UCSA_REQUIRE(grid_desc_ok(g), who ": bad grid descriptor (hashed levels need 2^k entries, sizes multiples of 8)")

struct Cell {
  uint32_t c[3];
  float f[3];
};

__device__ __forceinline__ Cell locate(const LevelGeom& lv, const float x01[3]) {
  Cell out;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(lv.scale, x01[d], 0.5f);
    const float fl = floorf(pos);
    out.c[d] = static_cast<uint32_t>(fl);
    out.f[d] = pos - fl;
  }
  return out;
}

__device__ __forceinline__ uint32_t corner_entry(const LevelGeom& lv, const Cell& cell, int corner) {
  const uint32_t x = cell.c[0] + (corner & 1);
  const uint32_t y = cell.c[1] + ((corner >> 1) & 1);
  const uint32_t z = cell.c[2] + ((corner >> 2) & 1);
  uint32_t idx;
  if (lv.hashed) {
    idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
    idx &= lv.entries - 1u;  // hashed levels hold 2^k entries (the cap 2^log2_hashmap_size; grid_desc_ok checks it)
  } else {
    // dense: coordinates are <= res, so idx <= res + res^2 + res^3 < 2 * entries (entries >= res^3, res >= 2):
    // the modulo is one conditional subtraction
    idx = x + y * lv.res + z * lv.res * lv.res;
    if (idx >= lv.entries) idx -= lv.entries;
  }
  return lv.offset + idx;
}

__device__ __forceinline__ float corner_weight(const Cell& cell, int corner) {
  float w = 1.0f;
#pragma unroll
  for (int d = 0; d < 3; ++d) w *= ((corner >> d) & 1) ? cell.f[d] : (1.0f - cell.f[d]);
  return w;
}

// The four (y, z) corner pairs of a cell: entries of the x corner (e0) and of the x+1 corner (e1), bit 0 of the pair
// index = +1 in y, bit 1 = +1 in z.  Same arithmetic as corner_entry, with the level-uniform work hoisted.
__device__ __forceinline__ void pair_entries(const LevelGeom& lv, const Cell& cell, uint32_t e0[4], uint32_t e1[4]) {
  if (lv.hashed) {
    const uint32_t hy = cell.c[1] * 2654435761u, hz = cell.c[2] * 805459861u;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t a = ((p & 1) ? hy + 2654435761u : hy) ^ ((p & 2) ? hz + 805459861u : hz);
      const uint32_t i = cell.c[0] ^ a, j = (cell.c[0] + 1u) ^ a;
      e0[p] = lv.offset + (i & (lv.entries - 1u));  // 2^k entries (grid_desc_ok)
      e1[p] = lv.offset + (j & (lv.entries - 1u));
    }
  } else {
    const uint32_t sy = lv.res, sz = lv.res * lv.res;
    const uint32_t base = cell.c[0] + cell.c[1] * sy + cell.c[2] * sz;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t i = base + ((p & 1) ? sy : 0u) + ((p & 2) ? sz : 0u);
      uint32_t j = i + 1u;
      if (i >= lv.entries) i -= lv.entries;
      if (j >= lv.entries) j -= lv.entries;
      e0[p] = lv.offset + i;
      e1[p] = lv.offset + j;
    }
  }
}

// One level: trilinear interpolation of the fp16 table, fp32 accumulate, split into the gather (issue) and the
// arithmetic (finish) so that a caller can keep the loads of several levels in flight.  `keep` = l2_policy_keep().
// The gather is bound by L1 wavefronts (one per distinct 128-byte line per load instruction), so the x and x+1
// corners of a pair come from ONE 16-byte load whenever they share an aligned group of four entries: dense levels
// index x contiguously, and hashed levels XOR x in with prime 1, so entry(x+1) = entry(x) ^ (x ^ (x+1)) stays in the
// group unless x % 4 == 3.  A quarter of the pairs need a second (predicated) 4-byte load.  `table` is 16-byte
// aligned and level offsets are multiples of 8 entries.
struct LevelGather {
  uint4 q[4];         // the aligned group of four entries holding e0 of each pair
  uint32_t extra[4];  // e1 when it lies outside that group
  uint32_t sel;       // per pair p, bits 8p..8p+4: (e0 & 3) | (e1 & 3) << 2 | outside << 4
  float f[3];
};

__device__ __forceinline__ void issue_level(const __half2* __restrict__ table, const LevelGeom& lv,
                                            const float x01[3], uint64_t keep, LevelGather& g) {
  const Cell cell = locate(lv, x01);
  uint32_t e0[4], e1[4];
  pair_entries(lv, cell, e0, e1);
  g.sel = 0u;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t group = e0[p] & ~3u;
    const bool outside = (e1[p] & ~3u) != group;
    g.q[p] = ld_keep_b128(table + group, keep);
    g.extra[p] = 0u;
    if (outside) g.extra[p] = ld_keep_b32(table + e1[p], keep);
    g.sel |= ((e0[p] & 3u) | ((e1[p] & 3u) << 2) | (outside ? 16u : 0u)) << (8 * p);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) g.f[d] = cell.f[d];
}

__device__ __forceinline__ float2 finish_level(const LevelGather& g) {
  Cell cell;
#pragma unroll
  for (int d = 0; d < 3; ++d) cell.f[d] = g.f[d];
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {  // corner order 0..7 = pairs (0,0) (1,0) (0,1) (1,1) in (y, z), x inside: as before
    const uint32_t s = g.sel >> (8 * p);
    const uint32_t b0 = pick_word(g.q[p], s & 3u);
    const uint32_t b1 = (s & 16u) ? g.extra[p] : pick_word(g.q[p], (s >> 2) & 3u);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&b0));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&b1));
    const float w0 = corner_weight(cell, 2 * p), w1 = corner_weight(cell, 2 * p + 1);
    a0 = fmaf(w0, f0.x, a0);
    a1 = fmaf(w0, f0.y, a1);
    a0 = fmaf(w1, f1.x, a0);
    a1 = fmaf(w1, f1.y, a1);
  }
  return make_float2(a0, a1);
}

__device__ __forceinline__ float2 interp_level(const __half2* __restrict__ table, const LevelGeom& lv,
                                               const float x01[3], uint64_t keep) {
  LevelGather g;
  issue_level(table, lv, x01, keep, g);
  return finish_level(g);
}

// Scatter of one level's gradient: grad_table[entry] += w_c * (g0, g1).  The x and x+1 corners of a pair land in
// adjacent entries whenever x is even (dense levels: index = x + ...; hashed levels: x enters the XOR with prime 1),
// and then go out as one 16-byte vector reduction.  `keep` = l2_policy_keep().
__device__ __forceinline__ void scatter_level(float* __restrict__ grad_table, const LevelGeom& lv,
                                              const float x01[3], float g0, float g1, uint64_t keep) {
  if (g0 == 0.f && g1 == 0.f) return;
  const Cell cell = locate(lv, x01);
  uint32_t e0[4], e1[4];
  pair_entries(lv, cell, e0, e1);  // pair p = corners 2p (x) and 2p+1 (x+1)
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float w0 = corner_weight(cell, 2 * p), w1 = corner_weight(cell, 2 * p + 1);
    if ((e0[p] ^ e1[p]) == 1u) {
      float* q = grad_table + 2ull * (e0[p] & ~1u);
      const bool first = (e0[p] & 1u) == 0u;  // one reduction, operands selected (less code than two call sites)
      const float wa = first ? w0 : w1, wb = first ? w1 : w0;
      red_keep_f32x4(q, wa * g0, wa * g1, wb * g0, wb * g1, keep);
    } else {
      red_keep_f32x2(grad_table + 2ull * e0[p], w0 * g0, w0 * g1, keep);
      red_keep_f32x2(grad_table + 2ull * e1[p], w1 * g0, w1 * g1, keep);
    }
  }
}

// Warp-collective scatter: neighbouring lanes hold neighbouring samples of a ray, and a run of them often sits in ONE
// cell (16 samples per cell at level 0 of the 256+256 configuration; about two per cell at resolution 446; after some
// training the importance samples crowd around the surfaces and share cells at every level).  The eight corner
// contributions are summed over each run of equal-cell lanes with a segmented shuffle reduction and only the head
// lane of a run issues reductions: fewer red.global operations, which are what bounds the backward pass (L2 request
// rate, about 1.3 LSU cycles per lane and operation).  The reduction stops at the first distance no run spans, and a
// warp without any two neighbours in one cell skips it altogether, so it is applied at EVERY level
// (measured, bench.py config 2, ms per density_bwd launch by finest merged resolution: 128: 0.960, 195: 0.935,
// 295: 0.910, 446: 0.870).  Every lane of the warp must call this; `active` = the lane has a sample with a
// non-zero gradient.
constexpr uint32_t kRunMaxRes = 0xffffu;  // default: all levels (UCSA_RUN_MAX_RES overrides, bring-up knob)

__device__ __forceinline__ void scatter_level_runs(float* __restrict__ grad_table, const LevelGeom& lv,
                                                   const float x01[3], float g0, float g1, bool active,
                                                   uint64_t keep) {
  const int lane = threadIdx.x & 31;
  const Cell cell = locate(lv, x01);
  // cell coordinates are <= 8192 (finest level): x, y in one word, z (and the inactive marker) in the other
  const uint32_t key_xy = cell.c[0] | (cell.c[1] << 16);
  const uint32_t key_z = active ? cell.c[2] : (0x80000000u | lane);
  // bit L of `links`: lanes L and L+1 are in the same cell
  const uint32_t next_xy = __shfl_down_sync(kFullMask, key_xy, 1);
  const uint32_t next_z = __shfl_down_sync(kFullMask, key_z, 1);
  const uint32_t links = __ballot_sync(kFullMask, lane < 31 && next_xy == key_xy && next_z == key_z);
  float c[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = active ? corner_weight(cell, k) : 0.f;
    c[2 * k] = w * g0;
    c[2 * k + 1] = w * g1;
  }
  bool head = active;
  if (links != 0u) {  // warp-uniform
    const uint32_t run = links >> lane;  // bit d-1 .. : links from this lane onwards
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      const bool take = (run & ((1u << d) - 1u)) == ((1u << d) - 1u);  // lanes L .. L+d all share the cell
      if (__ballot_sync(kFullMask, take) == 0u) break;  // no run reaches this far (nor any longer distance)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float o = __shfl_down_sync(kFullMask, c[i], d);
        if (take) c[i] += o;
      }
    }
    head = active && (lane == 0 || ((links >> (lane - 1)) & 1u) == 0u);
  }
  if (!head) return;
  uint32_t e0[4], e1[4];
  pair_entries(lv, cell, e0, e1);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float a0 = c[4 * p], a1 = c[4 * p + 1], b0 = c[4 * p + 2], b1 = c[4 * p + 3];  // x corner, x+1 corner
    if ((e0[p] ^ e1[p]) == 1u) {
      float* q = grad_table + 2ull * (e0[p] & ~1u);
      if (e0[p] & 1u) red_keep_f32x4(q, b0, b1, a0, a1, keep);
      else red_keep_f32x4(q, a0, a1, b0, b1, keep);
    } else {
      red_keep_f32x2(grad_table + 2ull * e0[p], a0, a1, keep);
      red_keep_f32x2(grad_table + 2ull * e1[p], b0, b1, keep);
    }
  }
}

// Sample position of slot k of ray n, exactly as renderer_semantics.py:171-173 then
// network_tcnn_semantics.py:133 evaluate it in eager fp32 (no FMA contraction).
__device__ __forceinline__ void sample_x01(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                           const float* __restrict__ aabb, float z, uint32_t n, float bound,
                                           float x01[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float p = __fadd_rn(rays_o[3 * n + d], __fmul_rn(rays_d[3 * n + d], z));
    p = fminf(fmaxf(p, aabb[d]), aabb[3 + d]);
    x01[d] = __fdiv_rn(__fadd_rn(p, bound), 2.0f * bound);
  }
}

}  // namespace ucsa
