// CUDA-core cross-check of heads_tc.cu (exported as ucsa_heads_*_simt).
// Colour and semantic heads on the masked-in samples: SH degree-4 direction encoding, colour MLP
// (SH16 + geo15 + 1 -> 64 -> 64 -> 3, sigmoid) and semantic MLP (geo15 + 1 -> 64 -> C).
// Rows a7, a12, a13 and their backward of SURVEY.md section 8 (network_tcnn_semantics.py:147-207).
//
// The reference evaluates the heads only where w > 1e-4 (boolean-mask gather / scatter, :159-161,:174);
// here the K surviving rows arrive as the compact list `sel` (row -> n*T+slot) built by compact_masked.
#include "mlp_simt.cuh"
#include "sh4.cuh"

namespace ucsa {
namespace {

constexpr int kSemOut = UCSA_MAX_CLASSES;  // semantic output layer is always 48 wide (pad16 of 33..48 classes)
constexpr int kColorW1 = 0, kColorW2 = 64 * 32, kColorW3 = kColorW2 + 64 * 64;  // offsets in w_color
constexpr int kSemW1 = 0, kSemW2 = 64 * 16;                                      // offsets in w_sem
constexpr int kSemParams = 64 * 16 + kSemOut * 64;

constexpr int kLd16 = tile_ld(16), kLd32 = tile_ld(32), kLd48 = tile_ld(48), kLd64 = tile_ld(64);

// colour input row: [SH(16) | geo_feat(15) | 1];  semantic input row: [geo_feat(15) | 1]
__device__ __forceinline__ void build_inputs(const float* __restrict__ rays_d, const __half* __restrict__ h,
                                             uint32_t flat, uint32_t t, __half* in_c, __half* in_s) {
  const uint32_t n = flat / t;
  float sh[16];
  sh4_eval(rays_d[3 * n + 0], rays_d[3 * n + 1], rays_d[3 * n + 2], sh);
  H8 a, b;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a.h[i] = __float2half_rn(sh[i]);
    b.h[i] = __float2half_rn(sh[8 + i]);
  }
  *reinterpret_cast<uint4*>(in_c) = a.v;
  *reinterpret_cast<uint4*>(in_c + 8) = b.v;
  H8 lo, hi, g0, g1;
  lo.v = __ldg(reinterpret_cast<const uint4*>(h + static_cast<uint64_t>(flat) * 16));
  hi.v = __ldg(reinterpret_cast<const uint4*>(h + static_cast<uint64_t>(flat) * 16 + 8));
#pragma unroll
  for (int i = 0; i < 7; ++i) g0.h[i] = lo.h[i + 1];
  g0.h[7] = hi.h[0];
#pragma unroll
  for (int i = 0; i < 7; ++i) g1.h[i] = hi.h[i + 1];
  g1.h[7] = __float2half_rn(1.0f);
  *reinterpret_cast<uint4*>(in_c + 16) = g0.v;
  *reinterpret_cast<uint4*>(in_c + 24) = g1.v;
  if (in_s != nullptr) {
    *reinterpret_cast<uint4*>(in_s) = g0.v;
    *reinterpret_cast<uint4*>(in_s + 8) = g1.v;
  }
}

template <int W>
__device__ __forceinline__ void copy_row_out(__half* __restrict__ dst, const __half* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(row + i);
}
template <int W>
__device__ __forceinline__ void copy_row_in(__half* __restrict__ row, const __half* __restrict__ src) {
#pragma unroll
  for (int i = 0; i < W; i += 8)
    *reinterpret_cast<uint4*>(row + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
}
template <int W>
__device__ __forceinline__ void clear_row(__half* __restrict__ row) {
#pragma unroll
  for (int i = 0; i < W; i += 8) *reinterpret_cast<uint4*>(row + i) = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kTileRows)
heads_fwd_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                 const float* __restrict__ rays_d, const __half* __restrict__ h,
                 const __half* __restrict__ w_color, const __half* __restrict__ w_sem, float* __restrict__ rgb,
                 __half* __restrict__ logits, __half* __restrict__ hc1, __half* __restrict__ hc2,
                 __half* __restrict__ hs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wc = reinterpret_cast<float*>(smem_raw);
  float* ws = wc + UCSA_COLOR_PARAMS;
  __half* in_c_t = reinterpret_cast<__half*>(ws + kSemParams);
  __half* in_s_t = in_c_t + kTileRows * kLd32;
  __half* h1_t = in_s_t + kTileRows * kLd16;
  __half* h2_t = h1_t + kTileRows * kLd64;
  __half* out_t = h2_t + kTileRows * kLd64;
  load_weights_f32(wc, w_color, UCSA_COLOR_PARAMS);
  load_weights_f32(ws, w_sem, kSemParams);
  __syncthreads();

  __half* in_c = in_c_t + threadIdx.x * kLd32;
  __half* in_s = in_s_t + threadIdx.x * kLd16;
  __half* h1 = h1_t + threadIdx.x * kLd64;
  __half* h2 = h2_t + threadIdx.x * kLd64;
  __half* out = out_t + threadIdx.x * kLd16;
  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + kTileRows - 1) / kTileRows;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * kTileRows + threadIdx.x;
    if (r >= k_rows) continue;  // thread-private rows, no barrier in the loop
    build_inputs(rays_d, h, static_cast<uint32_t>(sel[r]), t, in_c, in_s);
    dense_row_fwd<32, 64, true>(wc + kColorW1, in_c, h1);
    dense_row_fwd<64, 64, true>(wc + kColorW2, h1, h2);
    dense_row_fwd<64, 16, false>(wc + kColorW3, h2, out);
    if (hc1 != nullptr) copy_row_out<64>(hc1 + static_cast<uint64_t>(r) * 64, h1);
    if (hc2 != nullptr) copy_row_out<64>(hc2 + static_cast<uint64_t>(r) * 64, h2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __half2float(out[c]);
      rgb[static_cast<uint64_t>(r) * 3 + c] = round_h(1.0f / (1.0f + expf(-x)));  // fp16 sigmoid under autocast
    }
    dense_row_fwd<16, 64, true>(ws + kSemW1, in_s, h1);
    if (hs != nullptr) copy_row_out<64>(hs + static_cast<uint64_t>(r) * 64, h1);
    dense_row_fwd<64, kSemOut, false>(ws + kSemW2, h1, h2);  // 48 logits fit in the 64-wide row
    copy_row_out<kSemOut>(logits + static_cast<uint64_t>(r) * kSemOut, h2);
  }
}

__global__ void __launch_bounds__(kTileRows)
heads_bwd_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                 const float* __restrict__ rays_d, const __half* __restrict__ h,
                 const __half* __restrict__ w_color, const __half* __restrict__ w_sem,
                 const float* __restrict__ rgb, const __half* __restrict__ hc1, const __half* __restrict__ hc2,
                 const __half* __restrict__ hs, const float* __restrict__ d_rgb,
                 const float* __restrict__ d_logits, float loss_scale, __half* __restrict__ dh,
                 float* __restrict__ grad_w_color, float* __restrict__ grad_w_sem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wc = reinterpret_cast<float*>(smem_raw);
  float* ws = wc + UCSA_COLOR_PARAMS;
  __half* in_c_t = reinterpret_cast<__half*>(ws + kSemParams);  // colour phase           | semantic phase
  __half* h1_t = in_c_t + kTileRows * kLd32;                    // hc1                    | hs
  __half* h2_t = h1_t + kTileRows * kLd64;                      // hc2                    | d_logits (ld 56)
  __half* dpre_t = h2_t + kTileRows * kLd64;                    // d(colour pre-sigmoid)  | semantic input (ld 24)
  __half* dh2_t = dpre_t + kTileRows * kLd16;                   // d hc2                  | d hs
  __half* dh1_t = dh2_t + kTileRows * kLd64;                    // d hc1
  load_weights_f32(wc, w_color, UCSA_COLOR_PARAMS);
  load_weights_f32(ws, w_sem, kSemParams);
  __syncthreads();

  __half* in_c = in_c_t + threadIdx.x * kLd32;
  __half* h1 = h1_t + threadIdx.x * kLd64;
  __half* h2 = h2_t + threadIdx.x * kLd64;
  __half* dlog = h2_t + threadIdx.x * kLd48;
  __half* dpre = dpre_t + threadIdx.x * kLd16;
  __half* in_s = dpre;
  __half* dh2 = dh2_t + threadIdx.x * kLd64;
  __half* dh1 = dh1_t + threadIdx.x * kLd64;

  WGrad<32, 64> gc1;
  WGrad<64, 64> gc2;
  WGrad<64, 16> gc3;
  WGrad<16, 64> gs1;
  WGrad<64, kSemOut> gs2;
  gc1.clear(); gc2.clear(); gc3.clear(); gs1.clear(); gs2.clear();

  const float inv_scale = 1.0f / loss_scale;
  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + kTileRows - 1) / kTileRows;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * kTileRows + threadIdx.x;
    const bool valid = r < k_rows;
    const uint32_t flat = valid ? static_cast<uint32_t>(sel[r]) : 0u;
    float d_geo[15];
    // ---------------- colour phase
    if (valid) {
      build_inputs(rays_d, h, flat, t, in_c, nullptr);
      copy_row_in<64>(h1, hc1 + static_cast<uint64_t>(r) * 64);
      copy_row_in<64>(h2, hc2 + static_cast<uint64_t>(r) * 64);
      H8 lo, hi;
      lo.v = make_uint4(0, 0, 0, 0);
      hi.v = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float s = rgb[static_cast<uint64_t>(r) * 3 + c];
        lo.h[c] = __float2half_rn(d_rgb[static_cast<uint64_t>(r) * 3 + c] * s * (1.0f - s) * loss_scale);
      }
      *reinterpret_cast<uint4*>(dpre) = lo.v;
      *reinterpret_cast<uint4*>(dpre + 8) = hi.v;
      float dx64[64];
      dense_row_bwd<64, 16>(wc + kColorW3, dpre, dx64);
      store_row_masked<64, true>(dx64, h2, dh2);
      dense_row_bwd<64, 64>(wc + kColorW2, dh2, dx64);
      store_row_masked<64, true>(dx64, h1, dh1);
      float dx32[32];
      dense_row_bwd<32, 64>(wc + kColorW1, dh1, dx32);
#pragma unroll
      for (int i = 0; i < 15; ++i) d_geo[i] = round_h(dx32[16 + i]);
    } else {
      clear_row<32>(in_c); clear_row<64>(h1); clear_row<64>(h2); clear_row<16>(dpre);
      clear_row<64>(dh2); clear_row<64>(dh1);
    }
    __syncthreads();
    gc3.add_tile(dpre_t, h2_t);
    gc2.add_tile(dh2_t, h1_t);
    gc1.add_tile(dh1_t, in_c_t);
    __syncthreads();
    // ---------------- semantic phase (buffers re-used, see the table above)
    if (valid) {
      build_inputs(rays_d, h, flat, t, in_c, in_s);
      copy_row_in<64>(h1, hs + static_cast<uint64_t>(r) * 64);
#pragma unroll
      for (int c0 = 0; c0 < kSemOut; c0 += 8) {
        H8 o;
        const float4 a = __ldg(reinterpret_cast<const float4*>(d_logits + static_cast<uint64_t>(r) * kSemOut + c0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(d_logits + static_cast<uint64_t>(r) * kSemOut + c0 + 4));
        o.h[0] = __float2half_rn(a.x * loss_scale); o.h[1] = __float2half_rn(a.y * loss_scale);
        o.h[2] = __float2half_rn(a.z * loss_scale); o.h[3] = __float2half_rn(a.w * loss_scale);
        o.h[4] = __float2half_rn(b.x * loss_scale); o.h[5] = __float2half_rn(b.y * loss_scale);
        o.h[6] = __float2half_rn(b.z * loss_scale); o.h[7] = __float2half_rn(b.w * loss_scale);
        *reinterpret_cast<uint4*>(dlog + c0) = o.v;
      }
      float dx64[64];
      dense_row_bwd<64, kSemOut>(ws + kSemW2, dlog, dx64);
      store_row_masked<64, true>(dx64, h1, dh2);
      float dx16[16];
      dense_row_bwd<16, 64>(ws + kSemW1, dh2, dx16);
      // dL/dgeo_feat = colour part + semantic part, handed to density_bwd as fp16 (scaled)
      H8 lo, hi;
      lo.h[0] = __float2half_rn(0.f);
#pragma unroll
      for (int i = 0; i < 7; ++i) lo.h[i + 1] = __float2half_rn(d_geo[i] + round_h(dx16[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) hi.h[i] = __float2half_rn(d_geo[7 + i] + round_h(dx16[7 + i]));
      *reinterpret_cast<uint4*>(dh + static_cast<uint64_t>(flat) * 16) = lo.v;
      *reinterpret_cast<uint4*>(dh + static_cast<uint64_t>(flat) * 16 + 8) = hi.v;
    } else {
      clear_row<16>(in_s); clear_row<64>(h1); clear_row<kSemOut>(dlog); clear_row<64>(dh2);
    }
    __syncthreads();
    gs2.add_tile(h2_t, h1_t);   // d_logits tile (ld 56) x hs tile
    gs1.add_tile(dh2_t, dpre_t);  // d hs tile x semantic input tile (ld 24)
    __syncthreads();
  }
  gc1.flush(grad_w_color + kColorW1, inv_scale);
  gc2.flush(grad_w_color + kColorW2, inv_scale);
  gc3.flush(grad_w_color + kColorW3, inv_scale);
  gs1.flush(grad_w_sem + kSemW1, inv_scale);
  gs2.flush(grad_w_sem + kSemW2, inv_scale);
}

constexpr size_t kHeadsWeightsSmem = (UCSA_COLOR_PARAMS + kSemParams) * sizeof(float);
constexpr size_t kHeadsFwdSmem =
    kHeadsWeightsSmem + kTileRows * (kLd32 + kLd16 + 2 * kLd64 + kLd16) * sizeof(__half);
constexpr size_t kHeadsBwdSmem =
    kHeadsWeightsSmem + kTileRows * (kLd32 + 2 * kLd64 + kLd16 + 2 * kLd64) * sizeof(__half);

uint32_t heads_grid(uint32_t k_max, int ctas_per_sm) {
  const uint32_t tiles = (k_max + kTileRows - 1) / kTileRows;
  const uint32_t cap = kNumSMs * ctas_per_sm;
  return tiles < cap ? tiles : cap;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_heads_fwd_simt(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t,
                              uint32_t k_max, const float* rays_d, const void* h, const void* w_color_h,
                              const void* w_sem_h, uint32_t n_classes, float* rgb, void* logits, void* hc1,
                              void* hc2, void* hs, void* stream) {
  UCSA_REQUIRE(sel && ray_off && rays_d && h && w_color_h && w_sem_h && rgb && logits, "heads_fwd: null pointer");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "heads_fwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  if (k_max == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_fwd_kernel), kHeadsFwdSmem, "heads_fwd_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  heads_fwd_kernel<<<heads_grid(k_max, 2), kTileRows, kHeadsFwdSmem, as_stream(stream)>>>(
      sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_color_h),
      static_cast<const __half*>(w_sem_h), rgb, static_cast<__half*>(logits), static_cast<__half*>(hc1),
      static_cast<__half*>(hc2), static_cast<__half*>(hs));
  return check_launch("heads_fwd");
}

extern "C" int ucsa_heads_bwd_simt(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t,
                              uint32_t k_max, const float* rays_d, const void* h, const void* w_color_h,
                              const void* w_sem_h, uint32_t n_classes, const float* rgb, const void* hc1,
                              const void* hc2, const void* hs, const float* d_rgb, const float* d_logits,
                              float loss_scale, void* dh, float* grad_w_color, float* grad_w_sem, void* stream) {
  UCSA_REQUIRE(sel && ray_off && rays_d && h && w_color_h && w_sem_h && rgb && hc1 && hc2 && hs && d_rgb &&
                   d_logits && dh && grad_w_color && grad_w_sem,
               "heads_bwd: null pointer");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "heads_bwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  UCSA_REQUIRE(loss_scale > 0.f, "heads_bwd: loss_scale must be positive");
  if (k_max == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_bwd_kernel), kHeadsBwdSmem, "heads_bwd_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  heads_bwd_kernel<<<heads_grid(k_max, 1), kTileRows, kHeadsBwdSmem, as_stream(stream)>>>(
      sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_color_h),
      static_cast<const __half*>(w_sem_h), rgb, static_cast<const __half*>(hc1), static_cast<const __half*>(hc2),
      static_cast<const __half*>(hs), d_rgb, d_logits, loss_scale, static_cast<__half*>(dh), grad_w_color,
      grad_w_sem);
  return check_launch("heads_bwd");
}
