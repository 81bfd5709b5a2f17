// Front-to-back compositing of colour and semantics over the compact list of masked-in samples, forward and
// backward.  Row a14 / a15 of SURVEY.md section 8 (renderer_semantics.py:262-285).
//
// One warp per ray.  A row's 48 fp16 logits are one 96-byte load (lane l holds classes 2l, 2l+1), the
// soft-max is two warp reductions, and the class accumulators stay in registers for the whole ray.
#include "common.cuh"

namespace ucsa {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kLogitsLd = UCSA_MAX_CLASSES;

struct RowProb {
  float p0, p1;
};

__device__ __forceinline__ RowProb row_softmax(const __half* __restrict__ logits, int r, int lane, int n_classes) {
  float x0 = -INFINITY, x1 = -INFINITY;
  if (2 * lane < kLogitsLd) {
    const float2 f = __half22float2(
        __ldg(reinterpret_cast<const __half2*>(logits + static_cast<uint64_t>(r) * kLogitsLd) + lane));
    if (2 * lane < n_classes) x0 = f.x;
    if (2 * lane + 1 < n_classes) x1 = f.y;
  }
  const float m = warp_max(fmaxf(x0, x1));
  const float e0 = 2 * lane < n_classes ? expf(x0 - m) : 0.f;
  const float e1 = 2 * lane + 1 < n_classes ? expf(x1 - m) : 0.f;
  const float s = warp_sum(e0 + e1);
  return RowProb{e0 / s, e1 / s};
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
composite_fwd_kernel(const int32_t* __restrict__ ray_off, const float* __restrict__ w_sel,
                     const float* __restrict__ rgb, const __half* __restrict__ logits, uint32_t n_rays,
                     int n_classes, float* __restrict__ image, float* __restrict__ semantics) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (n >= n_rays) return;
  const int r0 = ray_off[n], r1 = ray_off[n + 1];
  float s0 = 0.f, s1 = 0.f, col = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float w = w_sel[r];
    const RowProb p = row_softmax(logits, r, lane, n_classes);
    s0 = fmaf(w, p.p0, s0);
    s1 = fmaf(w, p.p1, s1);
    if (lane < 3) col = fmaf(w, rgb[static_cast<uint64_t>(r) * 3 + lane], col);
  }
  if (lane < 3) image[static_cast<uint64_t>(n) * 3 + lane] = col;
  if (2 * lane < n_classes) semantics[static_cast<uint64_t>(n) * n_classes + 2 * lane] = s0;
  if (2 * lane + 1 < n_classes) semantics[static_cast<uint64_t>(n) * n_classes + 2 * lane + 1] = s1;
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
composite_bwd_kernel(const int32_t* __restrict__ ray_off, const float* __restrict__ w_sel,
                     const float* __restrict__ z_sel, const float* __restrict__ rgb,
                     const __half* __restrict__ logits, const float* __restrict__ g_image,
                     const float* __restrict__ g_depth, const float* __restrict__ g_sem,
                     const float* __restrict__ dnorm, uint32_t n_rays, int n_classes, float* __restrict__ d_rgb,
                     float* __restrict__ d_logits, float* __restrict__ d_w_sel) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (n >= n_rays) return;
  const int r0 = ray_off[n], r1 = ray_off[n + 1];
  const float gs0 = 2 * lane < n_classes ? g_sem[static_cast<uint64_t>(n) * n_classes + 2 * lane] : 0.f;
  const float gs1 = 2 * lane + 1 < n_classes ? g_sem[static_cast<uint64_t>(n) * n_classes + 2 * lane + 1] : 0.f;
  const float gi = lane < 3 ? g_image[static_cast<uint64_t>(n) * 3 + lane] : 0.f;
  const float gd = g_depth[n] / dnorm[n];
  for (int r = r0; r < r1; ++r) {
    const float w = w_sel[r];
    const RowProb p = row_softmax(logits, r, lane, n_classes);
    // semantics_n = sum_r w_r * softmax(l_r); the weights are detached on this branch (:270)
    const float q0 = w * gs0, q1 = w * gs1;
    const float dot = warp_sum(q0 * p.p0 + q1 * p.p1);
    if (2 * lane < kLogitsLd) {
      float2 o;
      o.x = p.p0 * (q0 - dot);
      o.y = p.p1 * (q1 - dot);
      *reinterpret_cast<float2*>(d_logits + static_cast<uint64_t>(r) * kLogitsLd + 2 * lane) = o;
    }
    float c = 0.f;
    if (lane < 3) {
      c = rgb[static_cast<uint64_t>(r) * 3 + lane];
      d_rgb[static_cast<uint64_t>(r) * 3 + lane] = w * gi;
    }
    const float dw = warp_sum(gi * c);
    if (lane == 0) d_w_sel[r] = dw + gd * z_sel[r];
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_composite_fwd(const int32_t* ray_off, const float* w_sel, const float* rgb, const void* logits,
                                  uint32_t n_rays, uint32_t n_classes, float* image, float* semantics,
                                  void* stream) {
  UCSA_REQUIRE(ray_off && w_sel && rgb && logits && image && semantics, "composite_fwd: null pointer");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "composite_fwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  if (n_rays == 0) return UCSA_OK;
  composite_fwd_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, 0, as_stream(stream)>>>(
      ray_off, w_sel, rgb, static_cast<const __half*>(logits), n_rays, static_cast<int>(n_classes), image,
      semantics);
  return check_launch("composite_fwd");
}

extern "C" int ucsa_composite_bwd(const int32_t* ray_off, const int32_t* sel, const float* w_sel,
                                  const float* z_sel, const float* rgb, const void* logits, const float* g_image,
                                  const float* g_depth, const float* g_semantics, const float* direction_norms,
                                  uint32_t n_rays, uint32_t n_classes, float* d_rgb, float* d_logits,
                                  float* d_w_sel, void* stream) {
  (void)sel;
  UCSA_REQUIRE(ray_off && w_sel && z_sel && rgb && logits && g_image && g_depth && g_semantics &&
                   direction_norms && d_rgb && d_logits && d_w_sel,
               "composite_bwd: null pointer");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "composite_bwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  if (n_rays == 0) return UCSA_OK;
  composite_bwd_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, 0, as_stream(stream)>>>(
      ray_off, w_sel, z_sel, rgb, static_cast<const __half*>(logits), g_image, g_depth, g_semantics,
      direction_norms, n_rays, static_cast<int>(n_classes), d_rgb, d_logits, d_w_sel);
  return check_launch("composite_bwd");
}
