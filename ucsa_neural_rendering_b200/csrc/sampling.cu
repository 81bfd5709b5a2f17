// Ray set-up and sample placement of the LIVE path: AABB slab test, uniform / stratified coarse samples,
// inverse-CDF importance resampling and the merge of the two sorted runs.
// Rows a2, a3, a9, a10 of SURVEY.md section 8.
#include "common.cuh"
#include "weights.cuh"

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
// One WARP per ray, several rays per CTA and no CTA-wide barrier; the cumulative product / sums are warp scans.
// Shared memory per warp: zc[Tc] sg[Tc] wt[Tc] (the cdf is formed in place over wt) zn[Tf] zs[Tf] (4-byte words) and
// bucket[Tf] member[Tf] (16-bit: Tc, Tf <= 4096), padded to 16 bytes: 6 KB per ray at 256 + 256 samples, so that the
// 28 rays an SM receives of a 4096-ray batch are resident at once (at 8 KB per ray the batch took 1.15 waves).
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
  float* cdf = wt;  // cdf[k] replaces wt[k] (same lane, same iteration)
  float* zn = wt + tc;
  float* zs = zn + tf;
  uint16_t* bucket = reinterpret_cast<uint16_t*>(zs + tf);  // coarse interval of every fine sample
  uint16_t* member = bucket + tf;                           // fine samples grouped by bucket
  const uint32_t t = tc + tf;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  const int tid = threadIdx.x & 31, nt = 32;

  for (uint32_t k0 = tid; k0 < tc; k0 += 8 * nt) {  // eight chunks of loads in flight per lane
    float zv[8], sv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t k = k0 + i * nt;
      zv[i] = k < tc ? z_cat[row + k] : 0.f;
      sv[i] = k < tc ? sigma[row + k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t k = k0 + i * nt;
      if (k < tc) {
        zc[k] = zv[i];
        sg[k] = sv[i];
      }
    }
  }
  __syncwarp();
  // coarse weights w_k = alpha_k * prod_{j<k} (1 - alpha_j + 1e-15) by a multiplicative warp scan over chunks of 32
  // samples (the scan of weights_fwd_kernel), then pdf = (w[1:-1] + 1e-5) / sum and cdf = [0, cumsum(pdf)] by an
  // additive one.  (The rounding order differs from a sequential cumprod / cumsum - as it does between torch's CPU and
  // CUDA scans; inverse-CDF sampling is insensitive to it except inside near-empty bins, see tests.)
  float carry = 1.0f, part = 0.f;
  for (uint32_t base = 0; base < tc; base += 32) {
    const uint32_t k = base + tid;
    float alpha = 0.f, keep = 1.0f;
    if (k < tc) {
      const float delta = k + 1 < tc ? __fsub_rn(zc[k + 1], zc[k]) : kLastDelta;
      alpha = 1.0f - expf(__fmul_rn(__fmul_rn(-delta, density_scale), sg[k]));
      keep = __fadd_rn(__fsub_rn(1.0f, alpha), kTransEps);
    }
    const float trans = chunk_transmittance(keep, carry, tid);
    if (k < tc) {
      const float w = __fmul_rn(alpha, trans);
      const float num = __fadd_rn(w, 1e-5f);  // pdf numerator over weights[1:-1]
      wt[k] = num;
      if (k >= 1 && k + 1 < tc) part += num;
    }
  }
  const float total = warp_sum(part);
  __syncwarp();
  float run = 0.f;
  if (tid == 0) cdf[0] = 0.f;
  for (uint32_t base = 1; base + 1 < tc; base += 32) {  // cdf[k] = sum_{1 <= j <= k} pdf_j, k = 1 .. tc-2
    const uint32_t k = base + tid;
    const bool in = k + 1 < tc;
    const float pdf = in ? __fdiv_rn(wt[k], total) : 0.f;
    const float incl = warp_scan_add(pdf, tid);
    if (in) cdf[k] = run + incl;
    run += __shfl_sync(kFullMask, incl, 31);
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
  // Stable rank of every fine sample among the fine samples, by counting instead of Tf^2 compares: a sample's bucket
  // is the coarse interval it falls into (lo = number of coarse samples <= z, exact compares on the final z values,
  // hence monotone in z); rank = samples in lower buckets + samples of the same bucket that sort before it
  // ((z, index) order).  Buckets hold about one sample each unless the pdf is sharply peaked, where the work
  // degrades gracefully towards the quadratic loop within the crowded buckets only.
  int* cnt = reinterpret_cast<int*>(sg);  // Tc + 1 counters over the (now free) sg | wt arrays
  for (uint32_t b = tid; b <= tc; b += nt) cnt[b] = 0;
  __syncwarp();
  for (uint32_t j = tid; j < tf; j += nt) {
    const float v = zn[j];
    uint32_t lo = 0, hi = tc;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (zc[mid] <= v) lo = mid + 1; else hi = mid;
    }
    bucket[j] = static_cast<uint16_t>(lo);
    atomicAdd(&cnt[lo], 1);
  }
  __syncwarp();
  {  // exclusive scan of the Tc + 1 counters: contiguous chunk per lane, warp scan of the chunk sums
    const uint32_t per = (tc + 1 + 31) / 32, b0 = tid * per, b1 = min(b0 + per, tc + 1);
    int sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += cnt[b];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(kFullMask, incl, o);
      if (tid >= o) incl += up;
    }
    int run = incl - sum;
    for (uint32_t b = b0; b < b1; ++b) {
      const int c = cnt[b];
      cnt[b] = run;
      run += c;
    }
  }
  __syncwarp();
  for (uint32_t j = tid; j < tf; j += nt) member[atomicAdd(&cnt[bucket[j]], 1)] = static_cast<uint16_t>(j);
  __syncwarp();  // now cnt[b] = end of bucket b = start of bucket b + 1
  for (uint32_t j = tid; j < tf; j += nt) {
    const float v = zn[j];
    const int b = bucket[j];
    const int s0 = b > 0 ? cnt[b - 1] : 0, s1 = cnt[b];
    int rank = s0;
    for (int sl = s0; sl < s1; ++sl) {
      const uint32_t k = static_cast<uint32_t>(member[sl]);
      const float w = zn[k];
      rank += (w < v || (w == v && k < j)) ? 1 : 0;
    }
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
  const uint32_t warp_floats = (3u * tc + 3u * tf + 3u) & ~3u;
  const size_t warp_bytes = warp_floats * sizeof(float);
  uint32_t warps = static_cast<uint32_t>((24u * 1024u) / warp_bytes);  // rays per CTA: 4 at 256 + 256 samples
  warps = warps > 8 ? 8 : (warps < 1 ? 1 : warps);
  const size_t smem = warps * warp_bytes;
  if (smem > 48 * 1024)  // size depends on the call: set every time (per device, checked)
    if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(resample_merge_kernel), smem, "resample_merge")) return rc;
  resample_merge_kernel<<<ceil_div(n_rays, warps), 32 * warps, smem, as_stream(stream)>>>(
      sigma, z_cat, u, seed, step_dev, ray_base, tc, tf, density_scale, order, n_rays, warp_floats);
  return check_launch("resample_merge");
}
