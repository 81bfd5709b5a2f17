// Parameter plumbing: fp32 master -> fp16 working copy, and the fused Adam step of row f1 of SURVEY.md
// section 8 (joint_train_lightning_net.py:897-919: Adam, betas (0.9,0.99), eps 1e-15, weight decay on the
// MLP group; GradScaler unscale / inf-skip semantics of :46,509-513 folded into the same pass).
#include "common.cuh"

namespace ucsa {
namespace {

__global__ void cast_kernel(const float* __restrict__ src, uint64_t n, __half* __restrict__ dst) {
  const uint64_t i = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (uint64_t k = i; k < n; ++k) dst[k] = __float2half_rn(src[k]);
  }
}

// torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected moments).
struct AdamCoef {
  float lr_over_bc1, b1, b2, eps, wd, ginv, bc2_sqrt;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCoef& c) {
  float grad = g * c.ginv;
  if (c.wd != 0.f) grad = fmaf(c.wd, p, grad);
  m = fmaf(c.b1, m, (1.0f - c.b1) * grad);
  v = fmaf(c.b2, v, (1.0f - c.b2) * grad * grad);
  const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  p = p - c.lr_over_bc1 * (m / denom);
}

// Four parameters per thread (16-byte accesses); the bias corrections are formed once per CTA, because with the
// step count on the device (CUDA-graph replay) they cost two powf per evaluation.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __half* __restrict__ p_h, uint64_t n, float lr, float b1, float b2, float eps, float wd, float ginv,
            const float* __restrict__ found_inf, float bc1, float bc2_sqrt, const int32_t* __restrict__ step_dev) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: skip the step on overflow
  __shared__ float bc[2];
  if (threadIdx.x == 0) {
    if (step_dev != nullptr) {  // step count lives on the device: bias corrections from it
      const float st = static_cast<float>(*step_dev);
      bc1 = 1.0f - powf(b1, st);
      bc2_sqrt = sqrtf(1.0f - powf(b2, st));
    }
    bc[0] = bc1;
    bc[1] = bc2_sqrt;
  }
  __syncthreads();
  const AdamCoef c{lr / bc[0], b1, b2, eps, wd, ginv, bc[1]};
  const uint64_t i = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 pp = *reinterpret_cast<const float4*>(p + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, c);
    adam_one(pp.y, gg.y, mm.y, vv.y, c);
    adam_one(pp.z, gg.z, mm.z, vv.z, c);
    adam_one(pp.w, gg.w, mm.w, vv.w, c);
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (p_h != nullptr) {
      const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p_h + i) = o;
    }
  } else {
    for (uint64_t k = i; k < n; ++k) {
      float pk = p[k], mk = m[k], vk = v[k];
      adam_one(pk, g[k], mk, vk, c);
      p[k] = pk;
      m[k] = mk;
      v[k] = vk;
      if (p_h != nullptr) p_h[k] = __float2half_rn(pk);
    }
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_cast_f32_to_f16(const float* src, uint64_t n, void* dst_h, void* stream) {
  UCSA_REQUIRE(src && dst_h, "cast_f32_to_f16: null pointer");
  UCSA_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst_h) % 8 == 0,
               "cast_f32_to_f16: buffers must be 16-byte (src) / 8-byte (dst) aligned");
  if (n == 0) return UCSA_OK;
  cast_kernel<<<ceil_div((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(src, n, static_cast<__half*>(dst_h));
  return check_launch("cast_f32_to_f16");
}

extern "C" int ucsa_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_h,
                              uint64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale_inv, const float* found_inf, uint32_t step, const int32_t* step_dev,
                              void* stream) {
  UCSA_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
  UCSA_REQUIRE(step >= 1 || step_dev != nullptr, "adam_step: step counts from 1");
  UCSA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                 reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(param_h) & 7u) == 0,
               "adam_step: buffers must be 16-byte aligned (fp16 copy: 8-byte)");
  if (n == 0) return UCSA_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<ceil_div((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq,
                                                               static_cast<__half*>(param_h), n, lr, beta1, beta2,
                                                               eps, weight_decay, grad_scale_inv, found_inf, bc1,
                                                               sqrtf(bc2), step_dev);
  return check_launch("adam_step");
}
