// Transmittance / weight arithmetic shared by the compositing kernels.
// renderer_semantics.py:238-247:  delta_s = z_{s+1}-z_s (1e10 for the last), alpha = 1-exp(-delta*ds*sigma),
// T_s = prod_{j<s} (1 - alpha_j + 1e-15), w_s = alpha_s * T_s.  One warp owns one ray; lanes take consecutive
// samples of a 32-sample chunk, a multiplicative warp scan plus a carried product gives T.
#pragma once

#include "common.cuh"

namespace ucsa {

struct SampleTerms {
  float expo;   // exp(-delta*ds*sigma)
  float alpha;  // 1 - expo
  float keep;   // 1 - alpha + 1e-15
  float delta;
};

__device__ __forceinline__ SampleTerms sample_terms(const float* zs, const float* sg, uint32_t s, uint32_t t,
                                                    float density_scale) {
  SampleTerms o;
  o.delta = s + 1 < t ? __fsub_rn(zs[s + 1], zs[s]) : kLastDelta;
  o.expo = expf(__fmul_rn(__fmul_rn(-o.delta, density_scale), sg[s]));
  o.alpha = 1.0f - o.expo;
  o.keep = __fadd_rn(__fsub_rn(1.0f, o.alpha), kTransEps);
  return o;
}

// Transmittance in front of sample `s = base + lane`; `carry` is the product over all earlier chunks and is
// advanced to include this chunk.  Lanes past the end must pass keep = 1.
__device__ __forceinline__ float chunk_transmittance(float keep, float& carry, int lane) {
  const float incl = warp_scan_mul(keep, lane);
  float excl = __shfl_up_sync(kFullMask, incl, 1);
  if (lane == 0) excl = 1.0f;
  const float trans = carry * excl;
  carry *= __shfl_sync(kFullMask, incl, 31);
  return trans;
}

// Exclusive suffix sum inside a chunk processed back to front: returns sum over samples after `lane` in this
// chunk plus `carry` (sum over all later chunks); advances carry.
__device__ __forceinline__ float chunk_suffix(float v, float& carry, int lane) {
  const float incl = warp_rscan_add(v, lane);
  float excl = __shfl_down_sync(kFullMask, incl, 1);
  if (lane == 31) excl = 0.f;
  const float out = carry + excl;
  carry += __shfl_sync(kFullMask, incl, 0);
  return out;
}

// d(sigma_s) given g_s = dL/dw_s (0 when masked out), T_s and R_s = sum_{k>s} g_k w_k
__device__ __forceinline__ float sigma_grad(const SampleTerms& st, float density_scale, float g, float trans,
                                            float suffix) {
  const float d_alpha = g * trans - suffix / st.keep;
  return d_alpha * st.expo * (st.delta * density_scale);
}

__device__ __forceinline__ int warp_scan_add_i32(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int up = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

}  // namespace ucsa
