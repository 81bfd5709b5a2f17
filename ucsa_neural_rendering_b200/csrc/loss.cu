// NeRF training losses and their gradients in one kernel (row f2 of SURVEY.md section 8:
// nr4seg/lightning/joint_train_lightning_net.py:199-221 and :503-507).
//
//   invalid_n   = (sum_c semantics[n,c] == 0)            -> semantics := 1 (uniform), label := -1 (ignored)
//   loss_color  = mean_{n,c} (image - gt_rgb)^2
//   loss_sem    = mean_n  [label != -1] * -log(semantics[n,label] / sum_c semantics[n,c] + 1e-15)
//   loss_depth  = mean_{n: gt_depth != 0} | depth / uom - gt_depth |
//   total       = (loss_color + w_sem * loss_sem + w_depth * loss_depth) * global_scale
// Outputs: loss[4] = (total, color, sem, depth) and d total / d image, d depth, d semantics.
// One warp per ray (coalesced over the class axis), kLossCtas CTAs; every CTA recounts the valid depths (N reads out of
// L2), the last CTA to finish adds the per-CTA partial sums in a fixed order, so the losses are deterministic.
// The partial sums and the done-counter live in a caller-owned scratch block (UCSA_LOSS_SCRATCH_BYTES, zero-filled
// once; the kernel re-arms it), so launches of different engines / streams never share state.
#include "common.cuh"

namespace ucsa {
namespace {

constexpr int kLossCtas = 148, kLossThreads = 512;  // one CTA per SM: 2368 warps, <= 2 rays each at 4096 rays
static_assert(UCSA_LOSS_SCRATCH_BYTES >= (kLossCtas * 3 + 1) * 4, "loss scratch too small");

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float tot = 0.f;
  if (wid == 0) {
    tot = lane < (blockDim.x >> 5) ? scratch[lane] : 0.f;
    tot = warp_sum(tot);
    if (lane == 0) scratch[0] = tot;
  }
  __syncthreads();
  return scratch[0];
}

__global__ void __launch_bounds__(kLossThreads)
nerf_loss_kernel(const float* __restrict__ image, const float* __restrict__ depth, const float* __restrict__ sem,
                 const __half* __restrict__ gt_rgb_h, const float* __restrict__ gt_rgb_f,
                 const int64_t* __restrict__ labels, const float* __restrict__ gt_depth, uint32_t n, uint32_t c,
                 float uom, float w_sem, float w_depth, float global_scale, float* __restrict__ loss,
                 float* __restrict__ g_image, float* __restrict__ g_depth, float* __restrict__ g_sem,
                 float* __restrict__ g_loss_partial, unsigned int* __restrict__ g_loss_done) {
  __shared__ float scratch[32];
  __shared__ bool is_last;
  float cnt = 0.f;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) cnt += gt_depth[i] != 0.f ? 1.f : 0.f;
  const float n_valid = block_sum(cnt, scratch);  // exact: a count below 2^24
  const float inv_n = 1.0f / static_cast<float>(n);
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  float s_color = 0.f, s_sem = 0.f, s_depth = 0.f;  // lane 0 carries the per-ray terms
  for (uint32_t i = warp; i < n; i += n_warps) {
    if (lane < 3) {  // colour
      const float gt = gt_rgb_h != nullptr ? __half2float(gt_rgb_h[3ull * i + lane]) : gt_rgb_f[3ull * i + lane];
      const float diff = image[3ull * i + lane] - gt;
      s_color += diff * diff;
      g_image[3ull * i + lane] = 2.0f * diff * inv_n * (1.0f / 3.0f) * global_scale;
    }
    if (lane == 0) {  // depth
      const float gd = gt_depth[i];
      float gdepth = 0.f;
      if (gd != 0.f) {
        const float diff = depth[i] / uom - gd;
        s_depth += fabsf(diff);
        gdepth = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) / (uom * n_valid) * w_depth * global_scale;
      }
      g_depth[i] = gdepth;
    }
    // semantics
    const float* s = sem + static_cast<uint64_t>(i) * c;
    float* gs = g_sem + static_cast<uint64_t>(i) * c;
    float part = 0.f;
    for (uint32_t k = lane; k < c; k += 32) part += s[k];
    const float tot = warp_sum(part);
    const int64_t label = labels[i];
    if (tot == 0.f || label < 0 || label >= static_cast<int64_t>(c)) {
      for (uint32_t k = lane; k < c; k += 32) gs[k] = 0.f;  // ignored ray: no gradient
    } else {
      const float sl = s[label];
      const float p = sl / tot;
      if (lane == 0) s_sem += -logf(p + 1e-15f);
      const float coeff = -1.0f / (p + 1e-15f) * inv_n * w_sem * global_scale;
      for (uint32_t k = lane; k < c; k += 32)
        gs[k] = coeff * ((k == static_cast<uint32_t>(label) ? 1.0f / tot : 0.f) - sl / (tot * tot));
    }
  }
  const float t_color = block_sum(s_color, scratch);
  const float t_sem = block_sum(s_sem, scratch);
  const float t_depth = block_sum(s_depth, scratch);
  if (threadIdx.x == 0) {
    g_loss_partial[blockIdx.x * 3 + 0] = t_color;
    g_loss_partial[blockIdx.x * 3 + 1] = t_sem;
    g_loss_partial[blockIdx.x * 3 + 2] = t_depth;
    __threadfence();
    is_last = atomicAdd(g_loss_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float sums[3] = {0.f, 0.f, 0.f};
  if (threadIdx.x < 3) {
    const volatile float* part = g_loss_partial;
    for (uint32_t b = 0; b < gridDim.x; ++b) sums[threadIdx.x] += part[b * 3 + threadIdx.x];
    scratch[threadIdx.x] = sums[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float l_color = scratch[0] * inv_n * (1.0f / 3.0f);
    const float l_sem = scratch[1] * inv_n;
    const float l_depth = n_valid > 0.f ? scratch[2] / n_valid : 0.f;
    loss[0] = (l_color + w_sem * l_sem + w_depth * l_depth) * global_scale;
    loss[1] = l_color;
    loss[2] = l_sem;
    loss[3] = l_depth;
    *g_loss_done = 0;  // re-arm for the next launch
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_nerf_loss(const float* image, const float* depth, const float* semantics, const void* gt_rgb_h,
                              const float* gt_rgb_f, const int64_t* labels, const float* gt_depth, uint32_t n_rays,
                              uint32_t n_classes, float one_m_to_scene_uom, float weight_semantics,
                              float weight_depth, float global_scale, float* loss4, float* g_image, float* g_depth,
                              float* g_semantics, void* scratch, void* stream) {
  UCSA_REQUIRE(image && depth && semantics && labels && gt_depth && loss4 && g_image && g_depth && g_semantics &&
                   scratch,
               "nerf_loss: null pointer");
  UCSA_REQUIRE((gt_rgb_h != nullptr) != (gt_rgb_f != nullptr), "nerf_loss: give gt_rgb as fp16 or as fp32");
  UCSA_REQUIRE(n_rays >= 1 && n_classes >= 1 && one_m_to_scene_uom > 0.f, "nerf_loss: bad sizes");
  nerf_loss_kernel<<<kLossCtas, kLossThreads, 0, as_stream(stream)>>>(image, depth, semantics, static_cast<const __half*>(gt_rgb_h),
                                                      gt_rgb_f, labels, gt_depth, n_rays, n_classes,
                                                      one_m_to_scene_uom, weight_semantics, weight_depth, global_scale,
                                                      loss4, g_image, g_depth, g_semantics,
                                                      static_cast<float*>(scratch),
                                                      reinterpret_cast<unsigned int*>(static_cast<float*>(scratch) +
                                                                                      kLossCtas * 3));
  return check_launch("nerf_loss");
}
