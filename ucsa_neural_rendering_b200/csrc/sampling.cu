// Ray set-up and sample placement of the LIVE path: AABB slab test, uniform / stratified coarse samples,
// inverse-CDF importance resampling and the merge of the two sorted runs.
// Rows a2, a3, a9, a10 of SURVEY.md section 8.
#include "common.cuh"

namespace ucsa {
namespace {

// ---------------------------------------------------------------- a2: raymarching.cu:62-115
__global__ void near_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                const float* __restrict__ aabb, uint32_t n_rays, float min_near,
                                float* __restrict__ nears, float* __restrict__ fars) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rays) return;
  float t_near = 0.f, t_far = 0.f;
  bool hit = true;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float o = rays_o[3 * n + ax];
    const float inv = 1.0f / rays_d[3 * n + ax];
    float lo = (aabb[ax] - o) * inv;
    float hi = (aabb[3 + ax] - o) * inv;
    if (lo > hi) {
      const float tmp = lo;
      lo = hi;
      hi = tmp;
    }
    if (ax == 0) {
      t_near = lo;
      t_far = hi;
    } else if (hit) {
      if (t_near > hi || lo > t_far) {
        hit = false;
      } else {
        if (lo > t_near) t_near = lo;
        if (hi < t_far) t_far = hi;
      }
    }
  }
  if (!hit) {
    nears[n] = fars[n] = 3.402823466e+38f;  // FLT_MAX, as std::numeric_limits<float>::max()
    return;
  }
  if (t_near < min_near) t_near = min_near;
  nears[n] = t_near;
  fars[n] = t_far;
}

// ---------------------------------------------------------------- a3: renderer_semantics.py:154-168
__device__ __forceinline__ float coarse_z(float near, float far, float lin) {
  return __fadd_rn(near, __fmul_rn(__fsub_rn(far, near), lin));
}

__global__ void sample_coarse_kernel(const float* __restrict__ nears, const float* __restrict__ fars,
                                     const float* __restrict__ lin, const float* __restrict__ t_rand,
                                     uint64_t seed, const int32_t* __restrict__ step_dev, uint32_t ray_base, int perturb,
                                     uint32_t n_rays, uint32_t tc, uint32_t t, float* __restrict__ z_cat) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<uint64_t>(n_rays) * tc) return;
  if (step_dev != nullptr) seed += 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(*step_dev);
  const uint32_t n = static_cast<uint32_t>(i / tc), k = static_cast<uint32_t>(i % tc);
  const float near = nears[n], far = fars[n];
  float z = coarse_z(near, far, lin[k]);
  if (perturb) {
    const float z_prev = k > 0 ? coarse_z(near, far, lin[k - 1]) : z;
    const float z_next = k + 1 < tc ? coarse_z(near, far, lin[k + 1]) : z;
    const float lower = k > 0 ? __fmul_rn(0.5f, __fadd_rn(z, z_prev)) : z;
    const float upper = k + 1 < tc ? __fmul_rn(0.5f, __fadd_rn(z_next, z)) : z;
    const float r = t_rand != nullptr ? t_rand[i] : uniform01(seed, ray_base + n, k, 0u);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), r));
  }
  z_cat[static_cast<uint64_t>(n) * t + k] = z;
}

// ---------------------------------------------------------------- a9/a10: renderer_semantics.py:182-222
// One WARP per ray, several rays per CTA and no CTA-wide barrier: the cumulative product / sum run sequentially in
// lane 0 (3*Tc dependent flops per ray, which keeps the summation order of torch.cumprod / torch.cumsum on the
// CPU) while the other warps of the SM are in their parallel phases.  Shared memory per warp:
// zc[Tc] sg[Tc] wt[Tc] cdf[Tc] zn[Tf] zs[Tf] (floats), padded to 16 bytes.
__global__ void __launch_bounds__(256)
resample_merge_kernel(const float* __restrict__ sigma, float* __restrict__ z_cat, const float* __restrict__ u,
                      uint64_t seed, const int32_t* __restrict__ step_dev, uint32_t ray_base, uint32_t tc, uint32_t tf,
                      float density_scale, int32_t* __restrict__ order, uint32_t n_rays, uint32_t warp_floats) {
  extern __shared__ __align__(16) float sm_all[];
  if (step_dev != nullptr) seed += 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(*step_dev);
  const uint32_t n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= n_rays) return;  // whole warps leave; nothing below synchronises across warps
  float* sm = sm_all + static_cast<size_t>(threadIdx.x >> 5) * warp_floats;
  float* zc = sm;
  float* sg = zc + tc;
  float* wt = sg + tc;
  float* cdf = wt + tc;
  float* zn = cdf + tc;
  float* zs = zn + tf;
  const uint32_t t = tc + tf;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  const int tid = threadIdx.x & 31, nt = 32;

  for (uint32_t k = tid; k < tc; k += nt) {
    zc[k] = z_cat[row + k];
    sg[k] = sigma[row + k];
  }
  __syncwarp();
  // alpha_k -> wt, and the transmittance factor (1 - alpha_k) + 1e-15 -> cdf (scratch for now)
  for (uint32_t k = tid; k < tc; k += nt) {
    const float delta = k + 1 < tc ? __fsub_rn(zc[k + 1], zc[k]) : kLastDelta;
    const float alpha = 1.0f - expf(__fmul_rn(__fmul_rn(-delta, density_scale), sg[k]));
    wt[k] = alpha;
    cdf[k] = __fadd_rn(__fsub_rn(1.0f, alpha), kTransEps);
  }
  __syncwarp();
  // The three scans below run in lane 0, in index order, so that every product / sum rounds exactly like
  // torch.cumprod / cumsum on the CPU.  Only the one dependent operation per element stays on the chain: each batch
  // of eight operands is loaded before the chain touches it and everything else is done by all lanes in between.
  if (tid == 0) {  // sg[k] := transmittance in front of sample k
    float trans = 1.0f;
    for (uint32_t k = 0; k < tc; k += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = k + i < tc ? cdf[k + i] : 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float before = trans;
        trans *= f[i];
        f[i] = before;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k + i < tc) sg[k + i] = f[i];
    }
  }
  __syncwarp();
  for (uint32_t k = tid; k < tc; k += nt) {
    const float w = __fmul_rn(wt[k], sg[k]);
    wt[k] = w;
    sg[k] = __fadd_rn(w, 1e-5f);  // pdf numerator over weights[1:-1] (no FMA contraction: w is rounded first)
  }
  __syncwarp();
  float total = 0.f;
  if (tid == 0) {
    for (uint32_t k = 1; k + 1 < tc; k += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = k + i + 1 < tc ? sg[k + i] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k + i + 1 < tc) total += f[i];
    }
  }
  total = __shfl_sync(kFullMask, total, 0);
  for (uint32_t k = tid; k < tc; k += nt) sg[k] = __fdiv_rn(sg[k], total);
  __syncwarp();
  if (tid == 0) {  // cdf = [0, cumsum(pdf)]  (tc-1 entries)
    float run = 0.f;
    cdf[0] = 0.f;
    for (uint32_t k = 1; k + 1 < tc; k += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = k + i + 1 < tc ? sg[k + i] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k + i + 1 < tc) run += f[i];
        f[i] = run;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k + i + 1 < tc) cdf[k + i] = f[i];
    }
  }
  __syncwarp();
  const uint32_t n_cdf = tc - 1;  // bins (mid-points) and cdf entries
  for (uint32_t j = tid; j < tf; j += nt) {
    const float uj = u != nullptr ? u[static_cast<uint64_t>(n) * tf + j] : uniform01(seed, ray_base + n, j, 1u);
    // searchsorted(cdf, u, right=True): number of entries <= u
    uint32_t lo = 0, hi = n_cdf;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
    }
    const uint32_t below = lo > 0 ? lo - 1 : 0;
    const uint32_t above = lo < n_cdf - 1 ? lo : n_cdf - 1;
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.0f;
    const float frac = __fdiv_rn(__fsub_rn(uj, cdf[below]), denom);
    // bins: z_mid_k = z_k + 0.5 * (z_{k+1} - z_k)
    const float b_lo = __fadd_rn(zc[below], __fmul_rn(0.5f, __fsub_rn(zc[below + 1], zc[below])));
    const float b_hi = __fadd_rn(zc[above], __fmul_rn(0.5f, __fsub_rn(zc[above + 1], zc[above])));
    const float z_new = __fadd_rn(b_lo, __fmul_rn(frac, __fsub_rn(b_hi, b_lo)));
    zn[j] = z_new;
  }
  __syncwarp();
  // Stable rank of every fine sample among the fine samples: #{k < j : z_k <= z_j} + #{k > j : z_k < z_j}.
  // O(Tf^2) compares per ray, but one compare per element on 16-byte broadcast reads of shared memory.
  const float4* zn4 = reinterpret_cast<const float4*>(zn);  // zn starts 16*Tc bytes into shared memory
  for (uint32_t j = tid; j < tf; j += nt) {
    const float v = zn[j];
    const uint32_t jg = j >> 2, n_groups = tf >> 2;
    uint32_t rank = 0;
    for (uint32_t g = 0; g < jg; ++g) {
      const float4 w = zn4[g];
      rank += (w.x <= v ? 1u : 0u) + (w.y <= v ? 1u : 0u) + (w.z <= v ? 1u : 0u) + (w.w <= v ? 1u : 0u);
    }
    for (uint32_t k = jg * 4; k < min(jg * 4 + 4, tf); ++k) {  // the group holding j itself
      const float w = zn[k];
      rank += (w < v || (w == v && k < j)) ? 1u : 0u;
    }
    for (uint32_t g = jg + 1; g < n_groups; ++g) {
      const float4 w = zn4[g];
      rank += (w.x < v ? 1u : 0u) + (w.y < v ? 1u : 0u) + (w.z < v ? 1u : 0u) + (w.w < v ? 1u : 0u);
    }
    for (uint32_t k = max(n_groups * 4, jg * 4 + 4); k < tf; ++k) rank += zn[k] < v ? 1u : 0u;
    zs[rank] = v;
  }
  __syncwarp();
  // The fine samples are stored in ascending order (slot Tc + r = r-th smallest): the set of samples is what
  // sample_pdf drew, only their order inside the cat buffer differs from torch.cat([z_vals, new_z_vals]), which no
  // result depends on (every consumer goes through `order`).  Neighbouring threads of the density kernels then
  // touch neighbouring cells in the fine pass as well.
  for (uint32_t r = tid; r < tf; r += nt) {
    const float v = zs[r];
    z_cat[row + tc + r] = v;
    // position in the merged order: fine samples go after coarse samples of equal depth
    uint32_t lo = 0, hi = tc;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (zc[mid] <= v) lo = mid + 1; else hi = mid;
    }
    order[row + r + lo] = static_cast<int32_t>(tc + r);
  }
  for (uint32_t k = tid; k < tc; k += nt) {
    const float v = zc[k];
    uint32_t lo = 0, hi = tf;  // number of fine samples strictly below v
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (zs[mid] < v) lo = mid + 1; else hi = mid;
    }
    order[row + k + lo] = static_cast<int32_t>(k);
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb6,
                                       uint32_t n_rays, float min_near, float* nears, float* fars,
                                       void* stream) {
  UCSA_REQUIRE(rays_o && rays_d && aabb6 && nears && fars, "near_far_from_aabb: null pointer");
  if (n_rays == 0) return UCSA_OK;
  near_far_kernel<<<ceil_div(n_rays, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, aabb6, n_rays, min_near,
                                                                        nears, fars);
  return check_launch("near_far_from_aabb");
}

extern "C" int ucsa_sample_coarse(const float* nears, const float* fars, const float* lin, const float* t_rand,
                                  uint64_t seed, const int32_t* step_dev, uint32_t ray_base, int perturb,
                                  uint32_t n_rays, uint32_t tc, uint32_t t, float* z_cat, void* stream) {
  UCSA_REQUIRE(nears && fars && lin && z_cat, "sample_coarse: null pointer");
  UCSA_REQUIRE(tc >= 1 && tc <= t, "sample_coarse: need 1 <= Tc <= T");
  if (n_rays == 0) return UCSA_OK;
  const uint64_t total = static_cast<uint64_t>(n_rays) * tc;
  sample_coarse_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(nears, fars, lin, t_rand, seed, step_dev,
                                                                            ray_base, perturb, n_rays, tc, t,
                                                                            z_cat);
  return check_launch("sample_coarse");
}

extern "C" int ucsa_resample_merge(const float* sigma, float* z_cat, const float* u, uint64_t seed,
                                   const int32_t* step_dev, uint32_t ray_base, uint32_t n_rays, uint32_t tc,
                                   uint32_t tf, float density_scale, int32_t* order, void* stream) {
  UCSA_REQUIRE(sigma && z_cat && order, "resample_merge: null pointer");
  UCSA_REQUIRE(tc >= 3 && tf >= 1 && tc <= 4096 && tf <= 4096, "resample_merge: need 3 <= Tc <= 4096, 1 <= Tf <= 4096");
  if (n_rays == 0) return UCSA_OK;
  const uint32_t warp_floats = (4u * tc + 2u * tf + 3u) & ~3u;
  const size_t warp_bytes = warp_floats * sizeof(float);
  uint32_t warps = static_cast<uint32_t>((48u * 1024u) / warp_bytes);  // rays per CTA
  warps = warps > 8 ? 8 : (warps < 1 ? 1 : warps);
  const size_t smem = warps * warp_bytes;
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaFuncSetAttribute(resample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  resample_merge_kernel<<<ceil_div(n_rays, warps), 32 * warps, smem, as_stream(stream)>>>(
      sigma, z_cat, u, seed, step_dev, ray_base, tc, tf, density_scale, order, n_rays, warp_floats);
  return check_launch("resample_merge");
}
