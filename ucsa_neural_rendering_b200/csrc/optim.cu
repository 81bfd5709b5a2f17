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
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, __half* __restrict__ p_h, uint64_t n, float lr, float b1,
                            float b2, float eps, float wd, float ginv, const float* __restrict__ found_inf,
                            float bc1, float bc2_sqrt, const int32_t* __restrict__ step_dev) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: skip the step on overflow
  if (step_dev != nullptr) {  // step count lives on the device (CUDA-graph replay): bias corrections from it
    const float st = static_cast<float>(*step_dev);
    bc1 = 1.0f - powf(b1, st);
    bc2_sqrt = sqrtf(1.0f - powf(b2, st));
  }
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float grad = g[i] * ginv;
  const float param = p[i];
  if (wd != 0.f) grad = fmaf(wd, param, grad);
  const float mi = fmaf(b1, m[i], (1.0f - b1) * grad);
  const float vi = fmaf(b2, v[i], (1.0f - b2) * grad * grad);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  const float out = param - (lr / bc1) * (mi / denom);
  p[i] = out;
  if (p_h != nullptr) p_h[i] = __float2half_rn(out);
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
  if (n == 0) return UCSA_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq,
                                                               static_cast<__half*>(param_h), n, lr, beta1, beta2,
                                                               eps, weight_decay, grad_scale_inv, found_inf, bc1,
                                                               sqrtf(bc2), step_dev);
  return check_launch("adam_step");
}
