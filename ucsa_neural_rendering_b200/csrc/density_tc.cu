// density(): hash-grid encode + sigma MLP (32->64->16, tcgen05) + trunc_exp, forward and backward.
// Rows a4/a5/a6/a8/a15 of SURVEY.md section 8 (network_tcnn_semantics.py:130-144, activation.py:7-19).
//
// Forward, per 128-sample tile of a persistent CTA (thread t = sample t = TMEM lane t):
//   gather 16 levels x 8 corners (fp16x2, L1/L2-resident table) -> encoded row straight into the shared-memory
//   A tile -> UMMA 128x64x32 -> TMEM -> ReLU/fp16 -> A tile of layer 2 -> UMMA 128x16x64 -> TMEM -> h, sigma.
//   The encoded features and hidden activations reach HBM only as the copies saved for the backward pass.
// Backward: dL/dh tile -> UMMA data gradients back to dL/d(encoded), weight gradients accumulated in TMEM over
//   all tiles of the CTA, and the hash-table gradient scattered with vector reductions (red.global.add.v2.f32).
// Several CTAs per SM (TMEM: 64 / 128 columns each) overlap one tile's gather/scatter phase with another's MMAs.
// The hash table and its gradient are tagged L2 evict_last, the per-sample streams evict_first (common.cuh).
//
// Both kernels are bound by the L2 REQUEST rate (scripts/gather_probe.cu: 285 G gathers/s, 192 G reductions/s, the
// same for 4-, 8- and 16-byte operations), hence one 16-byte request per (x, x+1) corner pair, same-cell runs merged
// per warp before they are reduced (grid.cuh) - and by instruction fetch when the 16 levels are unrolled (127 KB /
// 290 KB of SASS, "no instruction" the largest stall), hence the ROLLED level loops below.
#include <cstdlib>

#include "grid.cuh"
#include "mlp_umma.cuh"

namespace ucsa {
namespace {

using umma::Tile;

struct DensityArgs {
  const float* xyz;     // [S,3] or null
  const float* rays_o;  // [N,3]
  const float* rays_d;
  const float* aabb;
  const float* z_cat;   // [N,T]
  uint32_t n_rays, t, k0, span;  // span = k1-k0
  uint64_t n_samples;            // n_rays*span
  float bound;
  ucsa_grid_desc grid;
  // occupancy-grid mode (ucsa_grid_density): sample s = cell s of a [cascades, H, H, H] grid, jittered inside the cell
  uint32_t cell_h;  // 0 = off
  uint64_t cell_seed;
};

__device__ __forceinline__ uint64_t locate_sample(const DensityArgs& a, uint64_t s, float x01[3]) {
  if (a.cell_h != 0u) {
    // cell (cas, x, y, z) in the marching kernels' order (raymarching.cu:204: index = cas*H^3 + x*H^2 + y*H + z); point =
    // cell centre of cascade cas (half extent min(2^cas, bound)) + uniform jitter of one cell (torch-ngp's update rule)
    const uint32_t hh = a.cell_h;
    const uint64_t cube = static_cast<uint64_t>(hh) * hh * hh;
    const uint32_t cas = static_cast<uint32_t>(s / cube);
    const uint32_t idx = static_cast<uint32_t>(s % cube);
    const uint32_t c3[3] = {idx / (hh * hh), (idx / hh) % hh, idx % hh};
    const float cas_bound = fminf(exp2f(static_cast<float>(cas)), a.bound);
    const float half = cas_bound / static_cast<float>(hh);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float centre = 2.0f * static_cast<float>(c3[d]) / static_cast<float>(hh - 1) - 1.0f;
      const float jitter = uniform01(a.cell_seed, static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32) * 4u + d, 7u);
      const float p = centre * (cas_bound - half) + (2.0f * jitter - 1.0f) * half;
      x01[d] = __fdiv_rn(__fadd_rn(p, a.bound), 2.0f * a.bound);
    }
    return s;
  }
  if (a.xyz != nullptr) {
#pragma unroll
    for (int d = 0; d < 3; ++d) x01[d] = __fdiv_rn(__fadd_rn(a.xyz[3 * s + d], a.bound), 2.0f * a.bound);
    return s;
  }
  uint32_t n, k;
  if ((a.n_samples >> 32) == 0) {  // (kernel-uniform) 32-bit division: the 64-bit form is a ~100-instruction routine
    const uint32_t s32 = static_cast<uint32_t>(s);
    n = s32 / a.span;
    k = a.k0 + (s32 - n * a.span);
  } else {
    n = static_cast<uint32_t>(s / a.span);
    k = a.k0 + static_cast<uint32_t>(s % a.span);
  }
  const uint64_t flat = static_cast<uint64_t>(n) * a.t + k;
  sample_x01(a.rays_o, a.rays_d, a.aabb, a.z_cat[flat], n, a.bound, x01);
  return flat;
}

// Tile-layout saved activations (enc, hid): block index of the 128-sample tile starting at sample s0.  Requires
// k0, span and t to be multiples of 128 in ray mode (checked on the host), so that a tile never straddles rays or
// passes and the coarse pass, the fine pass and the backward pass agree on the block of every sample.
__device__ __forceinline__ uint64_t saved_block(const DensityArgs& a, uint64_t s0) {
  if (a.xyz != nullptr || a.cell_h != 0u) return s0 / 128;
  const uint64_t n = s0 / a.span;
  return (n * a.t + a.k0 + (s0 % a.span)) / 128;
}

constexpr uint32_t kW1Bytes = 64 / 8 * Tile<32>::kGroupBytes;  // [64][32]
constexpr uint32_t kW2Bytes = 16 / 8 * Tile<64>::kGroupBytes;  // [16][64]
constexpr uint32_t kFwdTmemCols = 64;   // layer-2 accumulator re-uses the columns of layer 1 once they are read
constexpr uint32_t kBwdTmemCols = 128;
constexpr int kFwdCtasPerSm = 6, kBwdCtasPerSm = 4;  // fwd: 0.478 ms (4) / 0.459 ms (6) per step with the rolled level loop

template <bool TILED>  // TILED: enc / hid are tile-layout buffers written with bulk copies (else row-major rows)
__global__ void __launch_bounds__(128, kFwdCtasPerSm)
density_fwd_tc_kernel(const DensityArgs a, const __half2* __restrict__ table, const __half* __restrict__ w_sigma,
                      float* __restrict__ sigma, __half* __restrict__ h, __half* __restrict__ enc,
                      __half* __restrict__ hid) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* w1 = smem;
  unsigned char* w2 = w1 + kW1Bytes;
  // the hidden tile re-uses the bytes of the encoded tile: layer 1 has consumed it (commit waited) before the
  // epilogue writes layer 1's activations.  Less shared memory per CTA = more L1 for the coarse grid levels.
  unsigned char* t_enc = w2 + kW2Bytes;
  unsigned char* t_hid = t_enc;
  unsigned char* tail = t_hid + Tile<64>::kBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 8);
  umma::load_weight_tile<32>(w1, w_sigma, 64);
  umma::load_weight_tile<64>(w2, w_sigma + 64 * 32, 16);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kFwdTmemCols);
  const uint32_t s_enc = umma::smem_u32(t_enc), s_hid = umma::smem_u32(t_hid);
  const uint32_t s_w1 = umma::smem_u32(w1), s_w2 = umma::smem_u32(w2);
  constexpr uint32_t kAcc1 = 0, kAcc2 = 0;
  const uint64_t keep = l2_policy_keep(), stream = l2_policy_stream();

  const int row = threadIdx.x;
  const uint64_t n_tiles = (a.n_samples + 127) / 128;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t s = tile * 128 + row;
    const bool valid = s < a.n_samples;
    uint64_t flat = 0;
    if (valid) {
      float x01[3];
      flat = locate_sample(a, s, x01);
      // Two levels per trip of a ROLLED loop: fully unrolled the kernel was 127 KB of SASS and "no instruction" its
      // largest stall; eight gathers in flight per thread are enough to reach the L2 request rate
      // (scripts/gather_probe.cu: the rate is flat from 4 loads per thread on).
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        LevelGather g[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) issue_level(table, level_geom(a.grid, 2 * c + j), x01, keep, g[j]);
        uint2 o;
        {
          const float2 f0 = finish_level(g[0]), f1 = finish_level(g[1]);
          const __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f1.x, f1.y);
          o.x = *reinterpret_cast<const uint32_t*>(&h0);
          o.y = *reinterpret_cast<const uint32_t*>(&h1);
        }
        // levels 2c, 2c+1 = columns 4c .. 4c+3 of the encoded row = half of the 16-byte chunk c / 2
        reinterpret_cast<uint2*>(Tile<32>::chunk(t_enc, row, c >> 1))[c & 1] = o;
        if (!TILED && enc != nullptr) *reinterpret_cast<uint2*>(enc + flat * 32 + c * 4) = o;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) *Tile<32>::chunk(t_enc, row, c) = make_uint4(0, 0, 0, 0);
    }
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<32, 64>(ctx.tmem + kAcc1, s_enc, s_w1);
      if (TILED && enc != nullptr) {  // the commit waits for the copy's reads: the epilogue overwrites the tile
        umma::bulk_store(umma::tile_block<32>(enc, saved_block(a, tile * 128)), s_enc, Tile<32>::kBytes, stream);
        umma::bulk_store_fence_reads();
      }
      umma::commit(ctx.bar);
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, true>(ctx, kAcc1 + c0, t_hid, c0);
    if (!TILED && hid != nullptr && valid) {
#pragma unroll
      for (int c = 0; c < 8; ++c) st_stream(hid + flat * 64 + c * 8, *Tile<64>::chunk(t_hid, row, c), stream);
    }
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<64, 16>(ctx.tmem + kAcc2, s_hid, s_w2);
      if (TILED && hid != nullptr) {
        umma::bulk_store(umma::tile_block<64>(hid, saved_block(a, tile * 128)), s_hid, Tile<64>::kBytes, stream);
        umma::bulk_store_fence_reads();
      }
      umma::commit(ctx.bar);
    }
    ctx.wait();
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(kAcc2), v);
    if (valid) {
      H8 lo, hi;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lo.h[i] = __float2half_rn(v[i]);
        hi.h[i] = __float2half_rn(v[8 + i]);
      }
      if (h != nullptr) {
        st_stream(h + flat * 16, lo.v, stream);
        st_stream(h + flat * 16 + 8, hi.v, stream);
      }
      st_stream_f32(sigma + flat, expf(__half2float(lo.h[0])), stream);  // trunc_exp forward, fp32
    }
  }
  umma::ctx_free(ctx, kFwdTmemCols);
}

// [64 x N] weight-gradient accumulator (UMMA M = 64) -> global fp32 (see mlp_umma.cuh)
template <int N, bool TRANSPOSED>
__device__ __forceinline__ void flush_wgrad(const umma::Ctx& ctx, uint32_t col0, float* __restrict__ grad, int ld,
                                            float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(col0 + c0), v);
    if (lane < 16) {
      const int m = warp * 16 + lane;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        atomicAdd(grad + (TRANSPOSED ? (c0 + i) * ld + m : m * ld + c0 + i), v[i] * scale);
    }
  }
}

template <bool TILED>
__global__ void __launch_bounds__(128)
density_bwd_tc_kernel(const DensityArgs a, const __half* __restrict__ w_sigma, const __half* __restrict__ h,
                      const __half* __restrict__ enc, const __half* __restrict__ hid,
                      const float* __restrict__ d_sigma, const __half* __restrict__ dh, const __half* __restrict__ dh2,
                      const uint8_t* __restrict__ use_geo, float loss_scale, float* __restrict__ grad_table,
                      float* __restrict__ grad_replicas, uint32_t n_replicas, float* __restrict__ grad_w,
                      uint32_t run_max_res) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* w1 = smem;
  unsigned char* w2 = w1 + kW1Bytes;
  unsigned char* t_enc = w2 + kW2Bytes;
  unsigned char* t_hid = t_enc + Tile<32>::kBytes;
  unsigned char* t_dhid = t_hid + Tile<64>::kBytes;
  // Dense (coarse) levels: every ray of a batch starts at the camera, so a handful of coarse cells receive most of
  // the adds and the L2 atomic units serialise on them.  Each CTA therefore adds into one of n_replicas private
  // copies of the dense part of the table; ucsa_reduce_grad_replicas folds them afterwards.
  const uint32_t dense_entries = dense_entry_count(a.grid);
  float* dense_base = (grad_replicas != nullptr && n_replicas > 0)
                          ? grad_replicas + static_cast<size_t>(blockIdx.x % n_replicas) * 2ull * dense_entries
                          : grad_table;
  unsigned char* t_dout = t_dhid + Tile<64>::kBytes;
  unsigned char* tail = t_dout + Tile<16>::kBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* ld_bar = reinterpret_cast<uint64_t*>(tail + 8);  // completion of the bulk loads of enc / hid
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 16);
  uint32_t ld_phase = 0;
  if (threadIdx.x == 64) umma::mbar_init(ld_bar, 1);
  umma::load_weight_tile<32>(w1, w_sigma, 64);
  umma::load_weight_tile<64>(w2, w_sigma + 64 * 32, 16);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kBwdTmemCols);
  const uint64_t keep = l2_policy_keep(), stream = l2_policy_stream();
  const uint32_t s_enc = umma::smem_u32(t_enc), s_hid = umma::smem_u32(t_hid), s_dhid = umma::smem_u32(t_dhid),
                 s_dout = umma::smem_u32(t_dout), s_w1 = umma::smem_u32(w1), s_w2 = umma::smem_u32(w2);
  constexpr uint32_t kAcc = 0, kG1 = 64, kG2 = 96;  // scratch 64 | dW1 [64x32] | dW2^T [64x16]

  const int row = threadIdx.x;
  const float inv_scale = 1.0f / loss_scale;
  const uint64_t n_tiles = (a.n_samples + 127) / 128;
  bool first = true;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t s = tile * 128 + row;
    const bool valid = s < a.n_samples;
    if (!first) ctx.wait();  // the weight-gradient MMAs of the previous tile have read the tiles
    if (TILED && threadIdx.x == 0) {
      const uint64_t block = saved_block(a, tile * 128);
      umma::mbar_expect_tx(ld_bar, Tile<32>::kBytes + Tile<64>::kBytes);
      umma::bulk_load(s_enc, umma::tile_block<32>(enc, block), Tile<32>::kBytes, ld_bar, stream);
      umma::bulk_load(s_hid, umma::tile_block<64>(hid, block), Tile<64>::kBytes, ld_bar, stream);
    }
    float x01[3] = {0.f, 0.f, 0.f};
    if (valid) {
      const uint64_t flat = locate_sample(a, s, x01);
      // dL/dh: element 0 through trunc_exp (activation.py:16-19), elements 1..15 = dL/dgeo_feat from the heads
      H8 lo, hi;
      if (use_geo != nullptr && dh != nullptr && use_geo[flat]) {
        lo.v = ld_stream(dh + flat * 16, stream);
        hi.v = ld_stream(dh + flat * 16 + 8, stream);
        if (dh2 != nullptr) {  // second share of dL/dgeo_feat (the semantic head's), summed like the in-place form did
          H8 lo2, hi2;
          lo2.v = ld_stream(dh2 + flat * 16, stream);
          hi2.v = ld_stream(dh2 + flat * 16 + 8, stream);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            lo.h[i] = __float2half_rn(__half2float(lo.h[i]) + __half2float(lo2.h[i]));
            hi.h[i] = __float2half_rn(__half2float(hi.h[i]) + __half2float(hi2.h[i]));
          }
        }
      } else {
        lo.v = make_uint4(0, 0, 0, 0);
        hi.v = make_uint4(0, 0, 0, 0);
      }
      const float h0 = __half2float(h[flat * 16]);
      const float gs = d_sigma != nullptr ? d_sigma[flat] : 0.f;
      lo.h[0] = __float2half_rn(gs * expf(fminf(fmaxf(h0, -15.f), 15.f)) * loss_scale);
      *Tile<16>::chunk(t_dout, row, 0) = lo.v;
      *Tile<16>::chunk(t_dout, row, 1) = hi.v;
      if (!TILED) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *Tile<64>::chunk(t_hid, row, c) = ld_stream(hid + flat * 64 + c * 8, stream);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *Tile<32>::chunk(t_enc, row, c) = ld_stream(enc + flat * 32 + c * 8, stream);
      }
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *Tile<16>::chunk(t_dout, row, 0) = z;
      *Tile<16>::chunk(t_dout, row, 1) = z;
      if (!TILED) {
#pragma unroll
        for (int c = 0; c < 8; ++c) *Tile<64>::chunk(t_hid, row, c) = z;
#pragma unroll
        for (int c = 0; c < 4; ++c) *Tile<32>::chunk(t_enc, row, c) = z;
      }
    }
    ctx.publish();
    if (TILED) {
      umma::mbar_wait(ld_bar, ld_phase);
      ld_phase ^= 1u;
    }
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<16, 64>(ctx.tmem + kAcc, s_dout, s_w2);  // d(hidden) = dL/dh . W2
      umma::commit(ctx.bar);
      umma::issue_wgrad<16>(ctx.tmem + kG2, s_hid, s_dout, first);  // d(W2)^T = hidden^T . dL/dh
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, false>(ctx, kAcc + c0, t_dhid, c0, t_hid);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<64, 32>(ctx.tmem + kAcc, s_dhid, s_w1);  // d(encoded) = d(hidden) . W1
      umma::commit(ctx.bar);
      umma::issue_wgrad<32>(ctx.tmem + kG1, s_dhid, s_enc, first);  // d(W1) = d(hidden)^T . encoded
    }
    ctx.wait();
    // dL/d(encoded) of this thread's sample -> (g0, g1) per level, parked in the hidden-activation tile (free by now:
    // every product that read it has completed) so that the level loop below can stay ROLLED.  Unrolled, with both
    // scatter variants per level, the kernel was 290 KB of SASS and "no instruction" its second largest stall.
    float2* stash = reinterpret_cast<float2*>(t_hid);  // [16 levels][128 threads]
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[16];
      umma::tmem_ld16(ctx.lane_addr(kAcc + 16 * half), v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        stash[(8 * half + i) * 128 + row] = make_float2(round_h(v[2 * i]) * inv_scale, round_h(v[2 * i + 1]) * inv_scale);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::commit(ctx.bar);  // covers the two weight-gradient products; waited on at the top of the next tile
    }
    if (grad_table != nullptr) {
      // two rolled loops, one scatter variant each (resolutions grow with the level): every loop body stays small
      int l = 0;
#pragma unroll 1
      for (; l < UCSA_GRID_LEVELS && a.grid.res[l] <= run_max_res; ++l) {  // coarse: merge same-cell runs per warp
        const LevelGeom lv = level_geom(a.grid, l);
        const float2 gl = stash[l * 128 + row];
        scatter_level_runs(lv.hashed ? grad_table : dense_base, lv, x01, gl.x, gl.y,
                           valid && (gl.x != 0.f || gl.y != 0.f), keep);
      }
#pragma unroll 1
      for (; l < UCSA_GRID_LEVELS; ++l) {  // fine: every sample has its own cell
        const LevelGeom lv = level_geom(a.grid, l);
        const float2 gl = stash[l * 128 + row];
        if (valid) scatter_level(lv.hashed ? grad_table : dense_base, lv, x01, gl.x, gl.y, keep);
      }
    }
    // the stash lives in the hidden tile, which the next tile refills (bulk copy or row stores of other threads)
    umma::fence_async_smem();
    __syncthreads();
    first = false;
  }
  ctx.wait();
  flush_wgrad<32, false>(ctx, kG1, grad_w, 32, inv_scale);
  flush_wgrad<16, true>(ctx, kG2, grad_w + 64 * 32, 64, inv_scale);
  umma::ctx_free(ctx, kBwdTmemCols);
}

constexpr size_t kFwdSmem = kW1Bytes + kW2Bytes + Tile<64>::kBytes + 64;
constexpr size_t kBwdSmem = kW1Bytes + kW2Bytes + Tile<32>::kBytes + 2 * Tile<64>::kBytes + Tile<16>::kBytes + 64;

int fill_args(DensityArgs& a, const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
              const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
              const ucsa_grid_desc* grid) {
  UCSA_REQUIRE_GRID(grid, "density");
  UCSA_REQUIRE(bound > 0.f, "density: bound must be positive");
  if (xyz != nullptr) {
    a = DensityArgs{xyz, nullptr, nullptr, nullptr, nullptr, n_rays, 1u, 0u, 1u, n_rays, bound, *grid, 0u, 0ull};
    return UCSA_OK;
  }
  UCSA_REQUIRE(rays_o && rays_d && aabb6 && z_cat, "density: rays_o/rays_d/aabb/z_cat required without xyz");
  UCSA_REQUIRE(k0 < k1 && k1 <= t, "density: bad slot range [%u,%u) of %u", k0, k1, t);
  a = DensityArgs{nullptr, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1 - k0,
                  static_cast<uint64_t>(n_rays) * (k1 - k0), bound, *grid, 0u, 0ull};
  return UCSA_OK;
}

int check_tiled(const DensityArgs& a, int tiled, const char* who) {
  if (!tiled || a.xyz != nullptr) return UCSA_OK;
  UCSA_REQUIRE(a.t % 128 == 0 && a.k0 % 128 == 0 && a.span % 128 == 0,
               "%s: tile-layout enc/hid need T, k0 and k1-k0 to be multiples of 128 (T=%u, k0=%u, span=%u)", who, a.t,
               a.k0, a.span);
  return UCSA_OK;
}

uint32_t persistent_grid(uint64_t n_samples, int ctas_per_sm) {
  const uint64_t tiles = (n_samples + 127) / 128;
  const uint64_t cap = static_cast<uint64_t>(kNumSMs) * ctas_per_sm;
  return static_cast<uint32_t>(tiles < cap ? tiles : cap);
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_density_fwd(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                                const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1,
                                float bound, const void* table_h, const ucsa_grid_desc* grid_host,
                                const void* w_sigma_h, float* sigma, void* h, void* enc, void* hid,
                                int tiled, void* stream) {
  DensityArgs a;
  if (int rc = fill_args(a, xyz, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1, bound, grid_host)) return rc;
  UCSA_REQUIRE(table_h && w_sigma_h && sigma && h, "density_fwd: null table/weights/outputs");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(table_h) & 15u) == 0, "density_fwd: the fp16 table must be 16-byte aligned");
  if (a.n_samples == 0) return UCSA_OK;
  if (int rc = check_tiled(a, tiled, "density_fwd")) return rc;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_fwd_tc_kernel<false>), kFwdSmem, "density_fwd_tc_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_fwd_tc_kernel<true>), kFwdSmem, "density_fwd_tc_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {  // tuning knob (bring-up): UCSA_DFWD_CTAS overrides the resident CTAs per SM
    const char* e = getenv("UCSA_DFWD_CTAS");
    ctas_per_sm = e ? atoi(e) : kFwdCtasPerSm;
    if (ctas_per_sm < 1 || ctas_per_sm > 8) ctas_per_sm = kFwdCtasPerSm;
    // tuning knob (bring-up): UCSA_DFWD_CARVEOUT = shared-memory carve-out in percent of the SM's 228 KB; what is left
    // is L1, which serves the gathers of the coarse levels (measured: 8 KB more shared memory per CTA, i.e. the next
    // carve-out step, cost 55 % of this kernel's speed)
    if (const char* c = getenv("UCSA_DFWD_CARVEOUT")) {
      const int pct = atoi(c);
      if (pct >= 0 && pct <= 100) {
        cudaFuncSetAttribute(density_fwd_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(density_fwd_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      }
    }
  }
  auto kernel = tiled ? density_fwd_tc_kernel<true> : density_fwd_tc_kernel<false>;
  kernel<<<persistent_grid(a.n_samples, ctas_per_sm), 128, kFwdSmem, as_stream(stream)>>>(
      a, static_cast<const __half2*>(table_h), static_cast<const __half*>(w_sigma_h), sigma,
      static_cast<__half*>(h), static_cast<__half*>(enc), static_cast<__half*>(hid));
  return check_launch("density_fwd");
}

extern "C" int ucsa_grid_density(const void* table_h, const ucsa_grid_desc* grid_host, const void* w_sigma_h,
                                 float bound, uint32_t cascades, uint32_t grid_h, uint64_t seed, float* sigma_cells,
                                 void* stream) {
  UCSA_REQUIRE_GRID(grid_host, "grid_density");
  UCSA_REQUIRE(table_h && w_sigma_h && sigma_cells, "grid_density: null pointer");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(table_h) & 15u) == 0, "grid_density: the fp16 table must be 16-byte aligned");
  UCSA_REQUIRE(bound > 0.f && cascades >= 1 && cascades <= 16 && grid_h >= 2 && grid_h <= 1024,
               "grid_density: bad occupancy-grid geometry");
  DensityArgs a{nullptr, nullptr, nullptr, nullptr, nullptr, 0u, 1u, 0u, 1u,
                static_cast<uint64_t>(cascades) * grid_h * grid_h * grid_h, bound, *grid_host, grid_h, seed};
  if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_fwd_tc_kernel<false>), kFwdSmem,
                                "density_fwd_tc_kernel"))
    return rc;
  density_fwd_tc_kernel<false><<<persistent_grid(a.n_samples, kFwdCtasPerSm), 128, kFwdSmem, as_stream(stream)>>>(
      a, static_cast<const __half2*>(table_h), static_cast<const __half*>(w_sigma_h), sigma_cells, nullptr, nullptr,
      nullptr);
  return check_launch("grid_density");
}

extern "C" int ucsa_density_bwd(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                                const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1,
                                float bound, const ucsa_grid_desc* grid_host, const void* w_sigma_h,
                                const void* h, const void* enc, const void* hid, int tiled, const float* d_sigma,
                                const void* dh, const void* dh2, const uint8_t* use_geo, float loss_scale, float* grad_table,
                                float* grad_replicas, uint32_t n_replicas, float* grad_w_sigma, void* stream) {
  DensityArgs a;
  if (int rc = fill_args(a, xyz, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1, bound, grid_host)) return rc;
  UCSA_REQUIRE(w_sigma_h && h && enc && hid && grad_w_sigma, "density_bwd: null saved tensors / outputs");
  UCSA_REQUIRE(loss_scale > 0.f, "density_bwd: loss_scale must be positive");
  if (a.n_samples == 0) return UCSA_OK;
  if (int rc = check_tiled(a, tiled, "density_bwd")) return rc;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_bwd_tc_kernel<false>), kBwdSmem, "density_bwd_tc_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_bwd_tc_kernel<true>), kBwdSmem, "density_bwd_tc_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  static uint32_t run_max_res = 0;
  if (run_max_res == 0) {  // tuning knob (bring-up): finest resolution that still merges same-cell runs per warp
    const char* e = getenv("UCSA_RUN_MAX_RES");
    run_max_res = e ? static_cast<uint32_t>(atoi(e)) : kRunMaxRes;
    if (run_max_res < 1) run_max_res = kRunMaxRes;
  }
  static int bwd_ctas = 0;
  if (bwd_ctas == 0) {  // tuning knob (bring-up): UCSA_DBWD_CTAS caps the resident CTAs per SM (grid size)
    const char* e = getenv("UCSA_DBWD_CTAS");
    bwd_ctas = e ? atoi(e) : kBwdCtasPerSm;
    if (bwd_ctas < 1 || bwd_ctas > kBwdCtasPerSm) bwd_ctas = kBwdCtasPerSm;
  }
  auto kernel = tiled ? density_bwd_tc_kernel<true> : density_bwd_tc_kernel<false>;
  kernel<<<persistent_grid(a.n_samples, bwd_ctas), 128, kBwdSmem, as_stream(stream)>>>(
      a, static_cast<const __half*>(w_sigma_h), static_cast<const __half*>(h), static_cast<const __half*>(enc),
      static_cast<const __half*>(hid), d_sigma, static_cast<const __half*>(dh), static_cast<const __half*>(dh2), use_geo,
      loss_scale, grad_table,
      grad_replicas, n_replicas, grad_w_sigma, run_max_res);
  return check_launch("density_bwd");
}

namespace ucsa {
namespace {
// grad_table[i] += sum_r replicas[r][i] over the dense part; replicas are left zeroed for the next step
__global__ void reduce_replicas_kernel(float* __restrict__ replicas, uint32_t n_replicas, uint32_t n_floats,
                                       float* __restrict__ grad_table) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n_floats) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t r = 0; r < n_replicas; ++r) {
    float4* p = reinterpret_cast<float4*>(replicas + static_cast<size_t>(r) * n_floats + i);
    const float4 v = *p;
    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    *p = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4* g = reinterpret_cast<float4*>(grad_table + i);
  float4 cur = *g;
  cur.x += acc.x, cur.y += acc.y, cur.z += acc.z, cur.w += acc.w;
  *g = cur;
}
}  // namespace
}  // namespace ucsa

extern "C" int ucsa_reduce_grad_replicas(float* grad_replicas, uint32_t n_replicas, const ucsa_grid_desc* grid_host,
                                         float* grad_table, void* stream) {
  UCSA_REQUIRE(grad_replicas && grid_host && grad_table, "reduce_grad_replicas: null pointer");
  UCSA_REQUIRE_GRID(grid_host, "reduce_grad_replicas");
  const uint32_t n_floats = 2u * dense_entry_count(*grid_host);  // entries are multiples of 8 -> divisible by 4
  if (n_replicas == 0 || n_floats == 0) return UCSA_OK;
  reduce_replicas_kernel<<<ceil_div(n_floats / 4, 256), 256, 0, as_stream(stream)>>>(grad_replicas, n_replicas, n_floats,
                                                                                    grad_table);
  return check_launch("reduce_grad_replicas");
}
