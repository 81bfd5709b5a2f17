// Module-level operators: the stand-alone forms of the pieces the reference exposes as attributes of
// SemanticNeRFNetwork (self.encoder, self.encoder_dir, self.sigma_net / color_net / semantics_net;
// network_tcnn_semantics.py:36-100) for callers that use them outside render().
#include "grid.cuh"
#include "mlp_simt.cuh"
#include "sh4.cuh"

namespace ucsa {
namespace {

__global__ void hashgrid_fwd_kernel(const float* __restrict__ x01, uint32_t n, const __half2* __restrict__ table,
                                    const ucsa_grid_desc grid, __half* __restrict__ enc) {
  // thread = (sample, level): consecutive threads share the sample, so its 64-byte output row is contiguous
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<uint64_t>(n) * UCSA_GRID_LEVELS) return;
  const uint32_t s = static_cast<uint32_t>(i / UCSA_GRID_LEVELS), l = static_cast<uint32_t>(i % UCSA_GRID_LEVELS);
  const float x[3] = {x01[3ull * s], x01[3ull * s + 1], x01[3ull * s + 2]};
  const float2 f = interp_level(table, level_geom(grid, l), x, l2_policy_keep());
  reinterpret_cast<__half2*>(enc)[i] = __floats2half2_rn(f.x, f.y);
}

__global__ void hashgrid_bwd_kernel(const float* __restrict__ x01, uint32_t n, const ucsa_grid_desc grid,
                                    const __half* __restrict__ d_enc, float inv_scale,
                                    float* __restrict__ grad_table) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<uint64_t>(n) * UCSA_GRID_LEVELS) return;
  const uint32_t s = static_cast<uint32_t>(i / UCSA_GRID_LEVELS), l = static_cast<uint32_t>(i % UCSA_GRID_LEVELS);
  const float x[3] = {x01[3ull * s], x01[3ull * s + 1], x01[3ull * s + 2]};
  const float2 g = __half22float2(reinterpret_cast<const __half2*>(d_enc)[i]);
  scatter_level(grad_table, level_geom(grid, l), x, g.x * inv_scale, g.y * inv_scale, l2_policy_keep());
}

__global__ void hashgrid_indices_kernel(const float* __restrict__ x01, uint32_t n, const ucsa_grid_desc grid,
                                        uint32_t* __restrict__ idx) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<uint64_t>(n) * UCSA_GRID_LEVELS) return;
  const uint32_t s = static_cast<uint32_t>(i / UCSA_GRID_LEVELS), l = static_cast<uint32_t>(i % UCSA_GRID_LEVELS);
  const float x[3] = {x01[3ull * s], x01[3ull * s + 1], x01[3ull * s + 2]};
  const LevelGeom lv = level_geom(grid, l);
  const Cell cell = locate(lv, x);
#pragma unroll
  for (int c = 0; c < 8; ++c) idx[i * 8 + c] = corner_entry(lv, cell, c);
}

__global__ void sh4_fwd_kernel(const float* __restrict__ d01, uint32_t n, __half* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // module input is already (d+1)/2 (network_tcnn_semantics.py:116)
  float sh[16];
  sh4_from01(d01[3ull * i], d01[3ull * i + 1], d01[3ull * i + 2], sh);
  H8 a, b;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a.h[k] = __float2half_rn(sh[k]);
    b.h[k] = __float2half_rn(sh[8 + k]);
  }
  *reinterpret_cast<uint4*>(out + 16ull * i) = a.v;
  *reinterpret_cast<uint4*>(out + 16ull * i + 8) = b.v;
}

// ------------------------------------------------------------------ stand-alone MLP, D3 == 0 -> two layers
template <int W>
__device__ __forceinline__ void row_in(__half* row, const __half* __restrict__ src) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(row + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
}
template <int W>
__device__ __forceinline__ void row_out(__half* __restrict__ dst, const __half* row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(row + i);
}
template <int W>
__device__ __forceinline__ void row_zero(__half* row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(row + i) = make_uint4(0, 0, 0, 0);
}

template <int D0, int D1, int D2, int D3>
struct MlpShape {
  static constexpr bool kThree = D3 > 0;
  static constexpr int kOut = kThree ? D3 : D2;
  static constexpr int kActs = kThree ? D1 + D2 : D1;
  static constexpr int kParams = D0 * D1 + D1 * D2 + (kThree ? D2 * D3 : 0);
  static constexpr int kFwdHalves = tile_ld(D0) + tile_ld(D1) + tile_ld(D2) + (kThree ? tile_ld(D3) : 0);
  static constexpr size_t kFwdSmem = kParams * sizeof(float) + kTileRows * kFwdHalves * sizeof(__half);
  static constexpr size_t kBwdSmem = kParams * sizeof(float) + kTileRows * 2 * kFwdHalves * sizeof(__half);
};

template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(kTileRows)
mlp_fwd_kernel(const __half* __restrict__ x, uint32_t n, const __half* __restrict__ w, __half* __restrict__ y,
               __half* __restrict__ acts) {
  using S = MlpShape<D0, D1, D2, D3>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wf = reinterpret_cast<float*>(smem_raw);
  __half* t0 = reinterpret_cast<__half*>(wf + S::kParams);
  __half* t1 = t0 + kTileRows * tile_ld(D0);
  __half* t2 = t1 + kTileRows * tile_ld(D1);
  __half* t3 = t2 + kTileRows * tile_ld(D2);
  load_weights_f32(wf, w, S::kParams);
  __syncthreads();
  __half* r0 = t0 + threadIdx.x * tile_ld(D0);
  __half* r1 = t1 + threadIdx.x * tile_ld(D1);
  __half* r2 = t2 + threadIdx.x * tile_ld(D2);
  const uint32_t n_tiles = (n + kTileRows - 1) / kTileRows;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * kTileRows + threadIdx.x;
    if (r >= n) continue;
    row_in<D0>(r0, x + static_cast<uint64_t>(r) * D0);
    dense_row_fwd<D0, D1, true>(wf, r0, r1);
    if (acts != nullptr) row_out<D1>(acts + static_cast<uint64_t>(r) * S::kActs, r1);
    if (S::kThree) {
      __half* r3 = t3 + threadIdx.x * tile_ld(D3 > 0 ? D3 : 16);
      dense_row_fwd<D1, D2, true>(wf + D0 * D1, r1, r2);
      if (acts != nullptr) row_out<D2>(acts + static_cast<uint64_t>(r) * S::kActs + D1, r2);
      dense_row_fwd<D2, (D3 > 0 ? D3 : 16), false>(wf + D0 * D1 + D1 * D2, r2, r3);
      row_out<(D3 > 0 ? D3 : 16)>(y + static_cast<uint64_t>(r) * S::kOut, r3);
    } else {
      dense_row_fwd<D1, D2, false>(wf + D0 * D1, r1, r2);
      row_out<D2>(y + static_cast<uint64_t>(r) * S::kOut, r2);
    }
  }
}

template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(kTileRows)
mlp_bwd_kernel(const __half* __restrict__ x, uint32_t n, const __half* __restrict__ w,
               const __half* __restrict__ acts, const __half* __restrict__ dy, float inv_scale,
               __half* __restrict__ dx, float* __restrict__ grad_w) {
  using S = MlpShape<D0, D1, D2, D3>;
  constexpr int DL = S::kOut;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wf = reinterpret_cast<float*>(smem_raw);
  __half* x_t = reinterpret_cast<__half*>(wf + S::kParams);
  __half* a1_t = x_t + kTileRows * tile_ld(D0);
  __half* a2_t = a1_t + kTileRows * tile_ld(D1);            // second hidden (three-layer) or dy (two-layer)
  __half* dy_t = a2_t + kTileRows * tile_ld(D2);            // dy (three-layer)
  __half* d1_t = dy_t + kTileRows * (S::kThree ? tile_ld(DL) : 0);
  __half* d2_t = d1_t + kTileRows * tile_ld(D1);
  load_weights_f32(wf, w, S::kParams);
  __syncthreads();
  __half* xr = x_t + threadIdx.x * tile_ld(D0);
  __half* a1 = a1_t + threadIdx.x * tile_ld(D1);
  __half* a2 = a2_t + threadIdx.x * tile_ld(D2);
  __half* dyr = dy_t + threadIdx.x * tile_ld(DL);
  __half* d1 = d1_t + threadIdx.x * tile_ld(D1);
  __half* d2 = d2_t + threadIdx.x * tile_ld(D2);
  WGrad<D0, D1> g0;
  WGrad<D1, D2> g1;
  WGrad<(S::kThree ? D2 : 64), (S::kThree ? DL : 16)> g2;  // unused for two layers
  g0.clear(); g1.clear(); g2.clear();
  const uint32_t n_tiles = (n + kTileRows - 1) / kTileRows;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * kTileRows + threadIdx.x;
    if (r < n) {
      row_in<D0>(xr, x + static_cast<uint64_t>(r) * D0);
      row_in<D1>(a1, acts + static_cast<uint64_t>(r) * S::kActs);
      float g_in[D0];
      if (S::kThree) {
        row_in<D2>(a2, acts + static_cast<uint64_t>(r) * S::kActs + D1);
        row_in<DL>(dyr, dy + static_cast<uint64_t>(r) * DL);
        float t2[D2];
        dense_row_bwd<D2, DL>(wf + D0 * D1 + D1 * D2, dyr, t2);
        store_row_masked<D2, true>(t2, a2, d2);
        float t1[D1];
        dense_row_bwd<D1, D2>(wf + D0 * D1, d2, t1);
        store_row_masked<D1, true>(t1, a1, d1);
      } else {
        row_in<D2>(a2, dy + static_cast<uint64_t>(r) * D2);  // a2 row holds dy
        float t1[D1];
        dense_row_bwd<D1, D2>(wf + D0 * D1, a2, t1);
        store_row_masked<D1, true>(t1, a1, d1);
      }
      dense_row_bwd<D0, D1>(wf, d1, g_in);
      if (dx != nullptr) {
#pragma unroll
        for (int k = 0; k < D0; k += 8) {
          H8 o;
#pragma unroll
          for (int i = 0; i < 8; ++i) o.h[i] = __float2half_rn(g_in[k + i]);
          *reinterpret_cast<uint4*>(dx + static_cast<uint64_t>(r) * D0 + k) = o.v;
        }
      }
    } else {
      row_zero<D0>(xr); row_zero<D1>(a1); row_zero<D2>(a2); row_zero<D1>(d1);
      if (S::kThree) { row_zero<DL>(dyr); row_zero<D2>(d2); }
    }
    __syncthreads();
    g0.add_tile(d1_t, x_t);
    if (S::kThree) {
      g1.add_tile(d2_t, a1_t);
      g2.add_tile(dy_t, a2_t);
    } else {
      g1.add_tile(a2_t, a1_t);
    }
    __syncthreads();
  }
  g0.flush(grad_w, inv_scale);
  g1.flush(grad_w + D0 * D1, inv_scale);
  if (S::kThree) g2.flush(grad_w + D0 * D1 + D1 * D2, inv_scale);
}

uint32_t mlp_grid(uint32_t n, int per_sm) {
  const uint32_t tiles = (n + kTileRows - 1) / kTileRows;
  const uint32_t cap = kNumSMs * per_sm;
  return tiles < cap ? tiles : cap;
}

template <int D0, int D1, int D2, int D3>
int launch_mlp_fwd(const void* x, uint32_t n, const void* w, void* y, void* acts, cudaStream_t st) {
  using S = MlpShape<D0, D1, D2, D3>;
  if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(mlp_fwd_kernel<D0, D1, D2, D3>), S::kFwdSmem, "mlp_fwd_kernel")) return rc;
  mlp_fwd_kernel<D0, D1, D2, D3><<<mlp_grid(n, 2), kTileRows, S::kFwdSmem, st>>>(
      static_cast<const __half*>(x), n, static_cast<const __half*>(w), static_cast<__half*>(y),
      static_cast<__half*>(acts));
  return check_launch("mlp_fwd");
}
template <int D0, int D1, int D2, int D3>
int launch_mlp_bwd(const void* x, uint32_t n, const void* w, const void* acts, const void* dy, float inv_scale,
                   void* dx, float* grad_w, cudaStream_t st) {
  using S = MlpShape<D0, D1, D2, D3>;
  if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(mlp_bwd_kernel<D0, D1, D2, D3>), S::kBwdSmem, "mlp_bwd_kernel")) return rc;
  mlp_bwd_kernel<D0, D1, D2, D3><<<mlp_grid(n, 1), kTileRows, S::kBwdSmem, st>>>(
      static_cast<const __half*>(x), n, static_cast<const __half*>(w), static_cast<const __half*>(acts),
      static_cast<const __half*>(dy), inv_scale, static_cast<__half*>(dx), grad_w);
  return check_launch("mlp_bwd");
}

int shape_id(const uint32_t* d, uint32_t n_layers) {
  if (n_layers == 2 && d[0] == 32 && d[1] == 64 && d[2] == 16) return 0;                 // sigma net
  if (n_layers == 3 && d[0] == 32 && d[1] == 64 && d[2] == 64 && d[3] == 16) return 1;   // colour net
  if (n_layers == 2 && d[0] == 16 && d[1] == 64 && d[2] == 48) return 2;                 // semantic net
  return -1;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_hashgrid_fwd(const float* x01, uint32_t n, const void* table_h, const ucsa_grid_desc* grid,
                                 void* enc, void* stream) {
  UCSA_REQUIRE(x01 && table_h && grid && enc, "hashgrid_fwd: null pointer");
  UCSA_REQUIRE_GRID(grid, "hashgrid_fwd");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(table_h) & 15u) == 0, "hashgrid_fwd: the fp16 table must be 16-byte aligned");
  if (n == 0) return UCSA_OK;
  const uint64_t total = static_cast<uint64_t>(n) * UCSA_GRID_LEVELS;
  hashgrid_fwd_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      x01, n, static_cast<const __half2*>(table_h), *grid, static_cast<__half*>(enc));
  return check_launch("hashgrid_fwd");
}

extern "C" int ucsa_hashgrid_bwd(const float* x01, uint32_t n, const ucsa_grid_desc* grid, const void* d_enc,
                                 float inv_loss_scale, float* grad_table, void* stream) {
  UCSA_REQUIRE(x01 && grid && d_enc && grad_table, "hashgrid_bwd: null pointer");
  UCSA_REQUIRE_GRID(grid, "hashgrid_bwd");
  if (n == 0) return UCSA_OK;
  const uint64_t total = static_cast<uint64_t>(n) * UCSA_GRID_LEVELS;
  hashgrid_bwd_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      x01, n, *grid, static_cast<const __half*>(d_enc), inv_loss_scale, grad_table);
  return check_launch("hashgrid_bwd");
}

extern "C" int ucsa_hashgrid_indices(const float* x01, uint32_t n, const ucsa_grid_desc* grid, uint32_t* idx,
                                     void* stream) {
  UCSA_REQUIRE(x01 && grid && idx, "hashgrid_indices: null pointer");
  UCSA_REQUIRE_GRID(grid, "hashgrid_indices");
  if (n == 0) return UCSA_OK;
  const uint64_t total = static_cast<uint64_t>(n) * UCSA_GRID_LEVELS;
  hashgrid_indices_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x01, n, *grid, idx);
  return check_launch("hashgrid_indices");
}

extern "C" int ucsa_sh4_fwd(const float* d01, uint32_t n, void* out_h, void* stream) {
  UCSA_REQUIRE(d01 && out_h, "sh4_fwd: null pointer");
  if (n == 0) return UCSA_OK;
  sh4_fwd_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d01, n, static_cast<__half*>(out_h));
  return check_launch("sh4_fwd");
}

extern "C" int ucsa_mlp_fwd_simt(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims, uint32_t n_layers,
                            void* y_h, void* acts_h, void* stream) {
  UCSA_REQUIRE(x_h && w_h && dims && y_h, "mlp_fwd: null pointer");
  if (n == 0) return UCSA_OK;
  switch (shape_id(dims, n_layers)) {
    case 0: return launch_mlp_fwd<32, 64, 16, 0>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    case 1: return launch_mlp_fwd<32, 64, 64, 16>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    case 2: return launch_mlp_fwd<16, 64, 48, 0>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    default:
      set_error("mlp_fwd: unsupported layer widths (supported: 32-64-16, 32-64-64-16, 16-64-48)");
      return UCSA_ERR_UNSUPPORTED;
  }
}

extern "C" int ucsa_mlp_bwd_simt(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims, uint32_t n_layers,
                            const void* acts_h, const void* dy_h, float inv_loss_scale, void* dx_h, float* grad_w,
                            void* stream) {
  UCSA_REQUIRE(x_h && w_h && dims && acts_h && dy_h && grad_w, "mlp_bwd: null pointer");
  if (n == 0) return UCSA_OK;
  switch (shape_id(dims, n_layers)) {
    case 0: return launch_mlp_bwd<32, 64, 16, 0>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    case 1: return launch_mlp_bwd<32, 64, 64, 16>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    case 2: return launch_mlp_bwd<16, 64, 48, 0>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    default:
      set_error("mlp_bwd: unsupported layer widths (supported: 32-64-16, 32-64-64-16, 16-64-48)");
      return UCSA_ERR_UNSUPPORTED;
  }
}
