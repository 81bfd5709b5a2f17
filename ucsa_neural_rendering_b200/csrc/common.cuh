// Shared helpers for libucsa_nerf.so (sm_100a).  Not part of the ABI.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/ucsa_nerf.h"

namespace ucsa {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr float kMaskThreshold = 1e-4f;  // renderer_semantics.py:249-250
constexpr float kLastDelta = 1e10f;      // renderer_semantics.py:187,239
constexpr float kTransEps = 1e-15f;      // renderer_semantics.py:193,244

void set_error(const char* fmt, ...);
int check_launch(const char* what);
// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (the attribute is per device / context);
// checked, and rejected above the 227 KB a CTA can have.
int set_max_dyn_smem(const void* fn, size_t bytes, const char* what);

#define UCSA_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::ucsa::set_error(__VA_ARGS__);           \
      return UCSA_ERR_INVALID_ARGUMENT;         \
    }                                           \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline uint32_t ceil_div(uint64_t a, uint64_t b) { return static_cast<uint32_t>((a + b - 1) / b); }

// ------------------------------------------------------------------ warp primitives
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
// inclusive scans over the 32 lanes
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float up = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v *= up;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float up = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += up;
  }
  return v;
}
// inclusive suffix sum: lane i gets sum over lanes >= i
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float dn = __shfl_down_sync(kFullMask, v, o);
    if (lane + o < 32) v += dn;
  }
  return v;
}

// ------------------------------------------------------------------ counter-based uniform numbers
// Used only when the caller does not inject its own random buffers (the reference draws them with
// torch.rand, renderer_semantics.py:28,166; there is no stream to be compatible with).
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float uniform01(uint64_t seed, uint32_t ray, uint32_t k, uint32_t stream) {
  uint64_t key = (static_cast<uint64_t>(ray) << 32) | (static_cast<uint64_t>(stream) << 24) | k;
  return static_cast<float>(mix64(mix64(seed) ^ key) >> 40) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------ fp16 helpers
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

union H8 {  // eight halves as one 16-byte vector
  uint4 v;
  __half2 h2[4];
  __half h[8];
};

__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// ------------------------------------------------------------------ L2 residency control
// The 25 MB fp16 hash table (forward) and the 52 MB fp32 gradient table (backward) are re-used by every sample and
// fit the 126 MB L2; the per-sample activation streams (hundreds of MB per step) are touched once.  Streams are
// tagged evict_first, tables evict_last, so that the streams do not push the tables out to HBM.
__device__ __forceinline__ uint64_t l2_policy_stream() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_stream(const void* ptr, uint64_t policy) {
  uint4 v;  // (not volatile: a read-only load may be scheduled freely, which keeps many of them in flight)
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void st_stream(void* ptr, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void st_stream_f32(float* ptr, float v, uint64_t policy) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_keep_b32(const void* ptr, uint64_t policy) {
  uint32_t v;
  asm("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const void* ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ uint4 ld_keep_b128(const void* ptr, uint64_t policy) {
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(ptr), "l"(policy));
  return v;
}
// word k (0..3) of a 16-byte vector without indexing registers dynamically
__device__ __forceinline__ uint32_t pick_word(const uint4& q, uint32_t k) {
  const uint32_t lo = (k & 1u) ? q.y : q.x, hi = (k & 1u) ? q.w : q.z;
  return (k & 2u) ? hi : lo;
}
__device__ __forceinline__ void red_keep_f32x2(float* addr, float a, float b, uint64_t policy) {
  asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(a), "f"(b), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void red_keep_f32x4(float* addr, float a, float b, float c, float d, uint64_t policy) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d), "l"(policy)
               : "memory");
}

}  // namespace ucsa
