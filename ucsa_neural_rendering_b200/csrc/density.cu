// density() on CUDA cores only: hash-grid encode + sigma MLP (32->64->16) + trunc_exp, forward and backward.
// On-device cross-check of density_tc.cu (the tcgen05 production kernels); exported as ucsa_density_*_simt.
// Rows a4/a5/a6/a8/a15 of SURVEY.md section 8 (network_tcnn_semantics.py:130-144, activation.py:7-19).
//
// One kernel per direction; the encoded features never leave the SM between the gather and the MLP.
// Persistent grid (a multiple of the SM count), one 128-sample tile per CTA iteration, one thread per
// sample.  Consecutive threads take consecutive samples of the same ray, so the coarse levels of the grid
// are read with high L1 locality.
#include "grid.cuh"
#include "mlp_simt.cuh"

namespace ucsa {
namespace {

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
};

// sample s of the launch -> flat index n*T+k and the [0,1]^3 position
__device__ __forceinline__ uint64_t locate_sample(const DensityArgs& a, uint64_t s, float x01[3]) {
  if (a.xyz != nullptr) {
#pragma unroll
    for (int d = 0; d < 3; ++d) x01[d] = __fdiv_rn(__fadd_rn(a.xyz[3 * s + d], a.bound), 2.0f * a.bound);
    return s;
  }
  const uint32_t n = static_cast<uint32_t>(s / a.span);
  const uint32_t k = a.k0 + static_cast<uint32_t>(s % a.span);
  const uint64_t flat = static_cast<uint64_t>(n) * a.t + k;
  sample_x01(a.rays_o, a.rays_d, a.aabb, a.z_cat[flat], n, a.bound, x01);
  return flat;
}

constexpr int kEncLd = tile_ld(32), kHidLd = tile_ld(64), kOutLd = tile_ld(16);

template <int W>
__device__ __forceinline__ void row_to_global(__half* __restrict__ dst, const __half* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(row + i);
}
template <int W>
__device__ __forceinline__ void global_to_row(__half* __restrict__ row, const __half* __restrict__ src) {
#pragma unroll
  for (int i = 0; i < W; i += 8)
    *reinterpret_cast<uint4*>(row + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
}
template <int W>
__device__ __forceinline__ void zero_row(__half* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(row + i) = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kTileRows)
density_fwd_kernel(const DensityArgs a, const __half2* __restrict__ table, const __half* __restrict__ w_sigma,
                   float* __restrict__ sigma, __half* __restrict__ h, __half* __restrict__ enc,
                   __half* __restrict__ hid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w1 = reinterpret_cast<float*>(smem_raw);  // [64][32]
  float* w2 = w1 + 64 * 32;                        // [16][64]
  __half* enc_t = reinterpret_cast<__half*>(w2 + 16 * 64);
  __half* hid_t = enc_t + kTileRows * kEncLd;
  __half* out_t = hid_t + kTileRows * kHidLd;
  load_weights_f32(w1, w_sigma, UCSA_SIGMA_PARAMS);
  __syncthreads();
  const uint64_t keep = l2_policy_keep();

  __half* enc_row = enc_t + threadIdx.x * kEncLd;
  __half* hid_row = hid_t + threadIdx.x * kHidLd;
  __half* out_row = out_t + threadIdx.x * kOutLd;
  const uint64_t n_tiles = (a.n_samples + kTileRows - 1) / kTileRows;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t s = tile * kTileRows + threadIdx.x;
    if (s >= a.n_samples) continue;  // rows are thread-private: no barrier inside the loop
    float x01[3];
    const uint64_t flat = locate_sample(a, s, x01);
#pragma unroll 4
    for (int l = 0; l < UCSA_GRID_LEVELS; ++l) {
      const float2 f = interp_level(table, level_geom(a.grid, l), x01, keep);
      *reinterpret_cast<__half2*>(enc_row + 2 * l) = __floats2half2_rn(f.x, f.y);
    }
    dense_row_fwd<32, 64, true>(w1, enc_row, hid_row);
    dense_row_fwd<64, 16, false>(w2, hid_row, out_row);
    sigma[flat] = expf(__half2float(out_row[0]));  // trunc_exp forward, fp32
    row_to_global<16>(h + flat * 16, out_row);
    if (enc != nullptr) row_to_global<32>(enc + flat * 32, enc_row);
    if (hid != nullptr) row_to_global<64>(hid + flat * 64, hid_row);
  }
}

__global__ void __launch_bounds__(kTileRows)
density_bwd_kernel(const DensityArgs a, const __half* __restrict__ w_sigma, const __half* __restrict__ h,
                   const __half* __restrict__ enc, const __half* __restrict__ hid,
                   const float* __restrict__ d_sigma, const __half* __restrict__ dh,
                   const uint8_t* __restrict__ use_geo, float loss_scale, float* __restrict__ grad_table,
                   float* __restrict__ grad_w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w1 = reinterpret_cast<float*>(smem_raw);
  float* w2 = w1 + 64 * 32;
  __half* enc_t = reinterpret_cast<__half*>(w2 + 16 * 64);
  __half* hid_t = enc_t + kTileRows * kEncLd;
  __half* dhid_t = hid_t + kTileRows * kHidLd;
  __half* dout_t = dhid_t + kTileRows * kHidLd;
  load_weights_f32(w1, w_sigma, UCSA_SIGMA_PARAMS);
  __syncthreads();
  const uint64_t keep = l2_policy_keep();

  __half* enc_row = enc_t + threadIdx.x * kEncLd;
  __half* hid_row = hid_t + threadIdx.x * kHidLd;
  __half* dhid_row = dhid_t + threadIdx.x * kHidLd;
  __half* dout_row = dout_t + threadIdx.x * kOutLd;
  WGrad<32, 64> g1;
  WGrad<64, 16> g2;
  g1.clear();
  g2.clear();
  const float inv_scale = 1.0f / loss_scale;
  const uint64_t n_tiles = (a.n_samples + kTileRows - 1) / kTileRows;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t s = tile * kTileRows + threadIdx.x;
    if (s < a.n_samples) {
      float x01[3];
      const uint64_t flat = locate_sample(a, s, x01);
      // dL/dh: element 0 through trunc_exp, elements 1..15 = dL/dgeo_feat from the heads
      H8 lo, hi;
      if (use_geo != nullptr && dh != nullptr && use_geo[flat]) {
        lo.v = __ldg(reinterpret_cast<const uint4*>(dh + flat * 16));
        hi.v = __ldg(reinterpret_cast<const uint4*>(dh + flat * 16 + 8));
      } else {
        lo.v = make_uint4(0, 0, 0, 0);
        hi.v = make_uint4(0, 0, 0, 0);
      }
      const float h0 = __half2float(h[flat * 16]);
      const float gs = d_sigma != nullptr ? d_sigma[flat] : 0.f;
      lo.h[0] = __float2half_rn(gs * expf(fminf(fmaxf(h0, -15.f), 15.f)) * loss_scale);
      *reinterpret_cast<uint4*>(dout_row) = lo.v;
      *reinterpret_cast<uint4*>(dout_row + 8) = hi.v;
      global_to_row<64>(hid_row, hid + flat * 64);
      global_to_row<32>(enc_row, enc + flat * 32);

      float dx64[64];
      dense_row_bwd<64, 16>(w2, dout_row, dx64);
      store_row_masked<64, true>(dx64, hid_row, dhid_row);
      float dx32[32];
      dense_row_bwd<32, 64>(w1, dhid_row, dx32);
      if (grad_table != nullptr) {
#pragma unroll
        for (int l = 0; l < UCSA_GRID_LEVELS; ++l) {
          scatter_level(grad_table, level_geom(a.grid, l), x01, round_h(dx32[2 * l]) * inv_scale,
                        round_h(dx32[2 * l + 1]) * inv_scale, keep);
        }
      }
    } else {
      zero_row<16>(dout_row);
      zero_row<64>(hid_row);
      zero_row<64>(dhid_row);
      zero_row<32>(enc_row);
    }
    __syncthreads();
    g2.add_tile(dout_t, hid_t);
    g1.add_tile(dhid_t, enc_t);
    __syncthreads();
  }
  g1.flush(grad_w, inv_scale);
  g2.flush(grad_w + 64 * 32, inv_scale);
}

constexpr size_t kFwdSmem = UCSA_SIGMA_PARAMS * sizeof(float) +
                            kTileRows * (kEncLd + kHidLd + kOutLd) * sizeof(__half);
constexpr size_t kBwdSmem = UCSA_SIGMA_PARAMS * sizeof(float) +
                            kTileRows * (kEncLd + 2 * kHidLd + kOutLd) * sizeof(__half);

int fill_args(DensityArgs& a, const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
              const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
              const ucsa_grid_desc* grid) {
  UCSA_REQUIRE_GRID(grid, "density");
  UCSA_REQUIRE(bound > 0.f, "density: bound must be positive");
  if (xyz != nullptr) {
    a = DensityArgs{xyz, nullptr, nullptr, nullptr, nullptr, n_rays, 1u, 0u, 1u, n_rays, bound, *grid};
    return UCSA_OK;
  }
  UCSA_REQUIRE(rays_o && rays_d && aabb6 && z_cat, "density: rays_o/rays_d/aabb/z_cat required without xyz");
  UCSA_REQUIRE(k0 < k1 && k1 <= t, "density: bad slot range [%u,%u) of %u", k0, k1, t);
  a = DensityArgs{nullptr, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1 - k0,
                  static_cast<uint64_t>(n_rays) * (k1 - k0), bound, *grid};
  return UCSA_OK;
}

uint32_t persistent_grid(uint64_t n_samples, int ctas_per_sm) {
  const uint64_t tiles = (n_samples + kTileRows - 1) / kTileRows;
  const uint64_t cap = static_cast<uint64_t>(kNumSMs) * ctas_per_sm;
  return static_cast<uint32_t>(tiles < cap ? tiles : cap);
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_density_fwd_simt(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                                const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1,
                                float bound, const void* table_h, const ucsa_grid_desc* grid_host,
                                const void* w_sigma_h, float* sigma, void* h, void* enc, void* hid,
                                void* stream) {
  DensityArgs a;
  if (int rc = fill_args(a, xyz, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1, bound, grid_host)) return rc;
  UCSA_REQUIRE(table_h && w_sigma_h && sigma && h, "density_fwd: null table/weights/outputs");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(table_h) & 15u) == 0, "density_fwd_simt: the fp16 table must be 16-byte aligned");
  if (a.n_samples == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_fwd_kernel), kFwdSmem, "density_fwd_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  density_fwd_kernel<<<persistent_grid(a.n_samples, 4), kTileRows, kFwdSmem, as_stream(stream)>>>(
      a, static_cast<const __half2*>(table_h), static_cast<const __half*>(w_sigma_h), sigma,
      static_cast<__half*>(h), static_cast<__half*>(enc), static_cast<__half*>(hid));
  return check_launch("density_fwd");
}

extern "C" int ucsa_density_bwd_simt(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                                const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1,
                                float bound, const ucsa_grid_desc* grid_host, const void* w_sigma_h,
                                const void* h, const void* enc, const void* hid, const float* d_sigma,
                                const void* dh, const uint8_t* use_geo, float loss_scale, float* grad_table,
                                float* grad_w_sigma, void* stream) {
  DensityArgs a;
  if (int rc = fill_args(a, xyz, rays_o, rays_d, aabb6, z_cat, n_rays, t, k0, k1, bound, grid_host)) return rc;
  UCSA_REQUIRE(w_sigma_h && h && enc && hid && grad_w_sigma, "density_bwd: null saved tensors / outputs");
  UCSA_REQUIRE(loss_scale > 0.f, "density_bwd: loss_scale must be positive");
  if (a.n_samples == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(density_bwd_kernel), kBwdSmem, "density_bwd_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  density_bwd_kernel<<<persistent_grid(a.n_samples, 3), kTileRows, kBwdSmem, as_stream(stream)>>>(
      a, static_cast<const __half*>(w_sigma_h), static_cast<const __half*>(h), static_cast<const __half*>(enc),
      static_cast<const __half*>(hid), d_sigma, static_cast<const __half*>(dh), use_geo, loss_scale, grad_table,
      grad_w_sigma);
  return check_launch("density_bwd");
}
