// Dense volume compositing (BASELINE.json configs[0]): every sample of every ray carries sigma, z, rgb and a
// C-class probability vector; weights, w > 1e-4 masks and the three composites follow
// renderer_semantics.py:238-285.  Row a14 (dense form) of SURVEY.md section 8; HBM-bound: 180 B/sample.
//
// One warp per ray.  Pass 1 computes the weights with a multiplicative warp scan and parks the masked
// weights in shared memory; pass 2 streams the probability block of the ray as 16-byte vectors.  Lane l of
// the 32/G*G active lanes (G = C/4) always sees the same four classes (l % G), so its accumulator is one
// float4 in registers for the whole ray; partial sums of the 32/G sample phases are folded at the end.
#include <cstdlib>

#include "mlp_umma.cuh"
#include "weights.cuh"

namespace ucsa {
namespace {

constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// weights of one ray -> wm[] (masked, smem) and optionally global `weights` (unmasked); returns depth sum
__device__ __forceinline__ float ray_weights(const float* __restrict__ zs, const float* __restrict__ sg, uint32_t t,
                                             float density_scale, int lane, float* wm,
                                             float* __restrict__ w_out) {
  float carry = 1.0f, dsum = 0.f;
  for (uint32_t base = 0; base < t; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < t;
    SampleTerms st{1.f, 0.f, 1.f, 0.f};
    if (valid) st = sample_terms(zs, sg, s, t, density_scale);
    const float trans = chunk_transmittance(valid ? st.keep : 1.0f, carry, lane);
    const float w = st.alpha * trans;
    if (valid) {
      const bool keep = w > kMaskThreshold;
      wm[s] = keep ? w : 0.f;
      if (w_out != nullptr) w_out[s] = w;
      if (keep) dsum += w * zs[s];
    }
  }
  return warp_sum(dsum);
}

template <bool VEC4>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
composite_dense_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ z,
                           const float* __restrict__ rgb, const float* __restrict__ prob,
                           const float* __restrict__ dnorm, uint32_t n_rays, uint32_t t, uint32_t c,
                           float density_scale, float* __restrict__ weights, float* __restrict__ depth,
                           float* __restrict__ image, float* __restrict__ semantics) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * kWarpsPerCta + wib;
  if (n >= n_rays) return;
  float* zs = sm + static_cast<size_t>(wib) * 3 * t;
  float* sg = zs + t;
  float* wm = sg + t;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  for (uint32_t s = lane; s < t; s += 32) {
    zs[s] = z[row + s];
    sg[s] = sigma[row + s];
  }
  __syncwarp();
  const float dsum = ray_weights(zs, sg, t, density_scale, lane, wm, weights ? weights + row : nullptr);
  if (lane == 0) depth[n] = dsum / dnorm[n];
  __syncwarp();

  // colour: lanes 0..29 = 10 samples x 3 channels per step
  {
    const int sub = lane / 3, ch = lane % 3;
    float acc = 0.f;
    if (lane < 30) {
      const float* p = rgb + row * 3;
      for (uint32_t s = sub; s < t; s += 10) acc = fmaf(wm[s], p[s * 3 + ch], acc);
    }
    // fold the 10 sample phases: lanes with equal channel
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) tot += __shfl_sync(kFullMask, acc, (3 * k + ch) % 32);
    if (lane < 3) image[static_cast<uint64_t>(n) * 3 + lane] = tot;
  }

  if (VEC4) {
    const uint32_t g = c / 4;       // float4 groups per sample
    const uint32_t spi = 32 / g;    // samples per warp step
    const uint32_t sub = lane / g, grp = lane % g;
    const bool active = static_cast<uint32_t>(lane) < spi * g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const float4* p = reinterpret_cast<const float4*>(prob + row * c);
      uint32_t s = sub;
      // four independent 16-byte loads in flight per lane
      for (; s + 3 * spi < t; s += 4 * spi) {
        const float4 v0 = ldg_stream(p + static_cast<size_t>(s) * g + grp);
        const float4 v1 = ldg_stream(p + static_cast<size_t>(s + spi) * g + grp);
        const float4 v2 = ldg_stream(p + static_cast<size_t>(s + 2 * spi) * g + grp);
        const float4 v3 = ldg_stream(p + static_cast<size_t>(s + 3 * spi) * g + grp);
        const float w0 = wm[s], w1 = wm[s + spi], w2 = wm[s + 2 * spi], w3 = wm[s + 3 * spi];
        acc.x = fmaf(w0, v0.x, acc.x); acc.y = fmaf(w0, v0.y, acc.y); acc.z = fmaf(w0, v0.z, acc.z); acc.w = fmaf(w0, v0.w, acc.w);
        acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y); acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
        acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y); acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
        acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y); acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
      }
      for (; s < t; s += spi) {
        const float4 v = ldg_stream(p + static_cast<size_t>(s) * g + grp);
        const float w = wm[s];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    }
    // fold the 32/G sample phases: lane l < G collects lanes l, l+G, l+2G, ...
    float4 tot = acc;
    for (uint32_t k = 1; k < spi; ++k) {
      const int src = (lane + k * g) % 32;
      tot.x += __shfl_sync(kFullMask, acc.x, src);
      tot.y += __shfl_sync(kFullMask, acc.y, src);
      tot.z += __shfl_sync(kFullMask, acc.z, src);
      tot.w += __shfl_sync(kFullMask, acc.w, src);
    }
    if (static_cast<uint32_t>(lane) < g)
      *reinterpret_cast<float4*>(semantics + static_cast<uint64_t>(n) * c + 4 * lane) = tot;
  } else {
    // generic class count: lanes stride over classes, samples sequential
    const float* p = prob + row * c;
    for (uint32_t cls = lane; cls < c; cls += 32) {
      float acc = 0.f;
      for (uint32_t s = 0; s < t; ++s) acc = fmaf(wm[s], p[static_cast<size_t>(s) * c + cls], acc);
      semantics[static_cast<uint64_t>(n) * c + cls] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged forward (the production path when a ray's inputs fit shared memory).  The kernel above leaves every
// warp with a chain of dependent load rounds (z / sigma -> weights -> rgb -> probabilities, ~1.5 us of HBM latency
// each under load): 30-34 us for a 95 MB problem.  Here ONE CTA of 128 threads owns one ray: one thread requests ALL
// inputs of the ray (z, sigma, rgb and the [T, C] probability block: four bulk copies completing on one mbarrier,
// 23 KB at T = 128, C = 40), nine such CTAs are resident per SM (so ~200 KB are in flight per SM while other CTAs
// compute), and the arithmetic reads shared memory only: thread = sample for the weights (block-wide product scan),
// thread = (class group, sample phase) for the probability sums.  (A first version with one persistent WARP per ray
// and a two-stage ring ran at one warp per scheduler: 7.5 cycles per instruction, 39 us.)
struct DenseStage {
  uint32_t z_off, sigma_off, rgb_off, prob_off, bytes;  // byte offsets inside the stage
};
constexpr int kTmaThreads = 128;

__global__ void __launch_bounds__(kTmaThreads)
composite_dense_fwd_tma_kernel(const float* __restrict__ sigma, const float* __restrict__ z,
                               const float* __restrict__ rgb, const float* __restrict__ prob,
                               const float* __restrict__ dnorm, uint32_t t, uint32_t c, float density_scale,
                               float* __restrict__ weights, float* __restrict__ depth, float* __restrict__ image,
                               float* __restrict__ semantics, DenseStage lay) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [stage | wm[t] | part: float4[subs][g] | warp_prod[4] red[4] | barrier]
  float* wm = reinterpret_cast<float*>(smem_raw + lay.bytes);
  const uint32_t g = c / 4;                 // float4 groups per sample
  const uint32_t subs = kTmaThreads / g;    // sample phases of the probability sums
  float4* part = reinterpret_cast<float4*>(smem_raw + lay.bytes + ((4 * t + 15) & ~15u));
  float* warp_prod = reinterpret_cast<float*>(part + subs * g);
  float* red = warp_prod + 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 4);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t n = blockIdx.x;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    const uint64_t stream = l2_policy_stream();
    const uint32_t dst = umma::smem_u32(smem_raw);
    umma::mbar_expect_tx(bar, lay.bytes);
    umma::bulk_load(dst + lay.z_off, z + row, 4 * t, bar, stream);
    umma::bulk_load(dst + lay.sigma_off, sigma + row, 4 * t, bar, stream);
    umma::bulk_load(dst + lay.rgb_off, rgb + row * 3, 12 * t, bar, stream);
    umma::bulk_load(dst + lay.prob_off, prob + row * c, 4 * t * c, bar, stream);
  }
  __syncthreads();  // the barrier is initialised before anybody polls it
  umma::mbar_wait(bar, 0);
  const float* zs = reinterpret_cast<const float*>(smem_raw + lay.z_off);
  const float* sg = reinterpret_cast<const float*>(smem_raw + lay.sigma_off);
  const float* cs = reinterpret_cast<const float*>(smem_raw + lay.rgb_off);
  const float4* ps = reinterpret_cast<const float4*>(smem_raw + lay.prob_off);

  // weights: thread = sample of a 128-sample chunk; transmittance by a block-wide product scan with the same
  // multiplication order as chunk_transmittance (prefix of earlier chunks, then earlier warps, then the warp scan)
  float carry = 1.0f, dsum = 0.f;
  for (uint32_t base = 0; base < t; base += kTmaThreads) {
    const uint32_t s = base + tid;
    const bool valid = s < t;
    SampleTerms st{1.f, 0.f, 1.f, 0.f};
    if (valid) st = sample_terms(zs, sg, s, t, density_scale);
    const float incl = warp_scan_mul(valid ? st.keep : 1.0f, lane);
    if (lane == 31) warp_prod[wid] = incl;
    __syncthreads();
    float prefix = carry;
    for (int w = 0; w < wid; ++w) prefix *= warp_prod[w];
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w_s = st.alpha * (prefix * excl);
    for (int w = 0; w < 4; ++w) carry *= warp_prod[w];
    if (valid) {
      const bool keep = w_s > kMaskThreshold;
      wm[s] = keep ? w_s : 0.f;
      if (weights != nullptr) weights[row + s] = w_s;
      if (keep) dsum += w_s * zs[s];
    }
    __syncthreads();  // warp_prod is rewritten by the next chunk; wm is complete after the last one
  }
  dsum = warp_sum(dsum);
  if (lane == 0) red[wid] = dsum;

  // colour: warp ch < 3 sums channel ch over all samples
  if (wid < 3) {
    float acc = 0.f;
    for (uint32_t s = lane; s < t; s += 32) acc = fmaf(wm[s], cs[s * 3 + wid], acc);
    acc = warp_sum(acc);
    if (lane == 0) image[static_cast<uint64_t>(n) * 3 + wid] = acc;
  }
  // probabilities: thread = (sample phase, class group); consecutive threads read consecutive 16-byte vectors
  {
    const uint32_t sub = tid / g, grp = tid % g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (static_cast<uint32_t>(tid) < subs * g) {
#pragma unroll 4
      for (uint32_t s = sub; s < t; s += subs) {
        const float4 v = ps[static_cast<size_t>(s) * g + grp];
        const float w = wm[s];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      part[sub * g + grp] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) depth[n] = (red[0] + red[1] + red[2] + red[3]) / dnorm[n];
  if (static_cast<uint32_t>(tid) < c) {  // class tid: fold the sample phases in a fixed order
    const float* pf = reinterpret_cast<const float*>(part);
    float tot = 0.f;
    for (uint32_t k = 0; k < subs; ++k) tot += pf[(k * g) * 4 + tid];
    semantics[static_cast<uint64_t>(n) * c + tid] = tot;
  }
}

template <bool VEC4>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
composite_dense_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ z,
                           const float* __restrict__ rgb, const float* __restrict__ weights,
                           const float* __restrict__ dnorm, const float* __restrict__ g_depth,
                           const float* __restrict__ g_image, const float* __restrict__ g_sem, uint32_t n_rays,
                           uint32_t t, uint32_t c, float density_scale, float* __restrict__ d_sigma,
                           float* __restrict__ d_rgb, float* __restrict__ d_prob) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * kWarpsPerCta + wib;
  if (n >= n_rays) return;
  float* zs = sm + static_cast<size_t>(wib) * 4 * t;
  float* sg = zs + t;
  float* wm = sg + t;  // masked weights
  float* tr = wm + t;  // transmittance
  const uint64_t row = static_cast<uint64_t>(n) * t;
  for (uint32_t s = lane; s < t; s += 32) {
    zs[s] = z[row + s];
    sg[s] = sigma[row + s];
  }
  __syncwarp();
  const float gi0 = g_image[static_cast<uint64_t>(n) * 3 + 0], gi1 = g_image[static_cast<uint64_t>(n) * 3 + 1],
              gi2 = g_image[static_cast<uint64_t>(n) * 3 + 2];
  const float gd = g_depth[n] / dnorm[n];

  float carry = 1.0f;
  for (uint32_t base = 0; base < t; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < t;
    float keep_f = 1.0f;
    if (valid) keep_f = sample_terms(zs, sg, s, t, density_scale).keep;
    const float trans = chunk_transmittance(keep_f, carry, lane);
    if (valid) {
      const float w = weights[row + s];
      tr[s] = trans;
      wm[s] = w > kMaskThreshold ? w : 0.f;
    }
  }
  __syncwarp();
  // d_sigma: back-to-front suffix sums of g_k * w_k over masked-in samples
  float suffix_carry = 0.f;
  const uint32_t n_chunks = (t + 31) / 32;
  for (uint32_t ch = n_chunks; ch-- > 0;) {
    const uint32_t s = ch * 32 + lane;
    const bool valid = s < t;
    float g = 0.f, w = 0.f;
    if (valid) {
      w = weights[row + s];
      if (w > kMaskThreshold) {
        const float* c3 = rgb + (row + s) * 3;
        g = gi0 * c3[0] + gi1 * c3[1] + gi2 * c3[2] + gd * zs[s];
      }
    }
    const float suffix = chunk_suffix(g * w, suffix_carry, lane);
    if (valid) {
      const SampleTerms st = sample_terms(zs, sg, s, t, density_scale);
      d_sigma[row + s] = sigma_grad(st, density_scale, g, tr[s], suffix);
    }
  }
  // d_rgb = masked w * g_image
  if (lane < 30) {
    const int sub = lane / 3, chn = lane % 3;
    const float gi = chn == 0 ? gi0 : (chn == 1 ? gi1 : gi2);
    float* p = d_rgb + row * 3;
    for (uint32_t s = sub; s < t; s += 10) p[s * 3 + chn] = wm[s] * gi;
  }
  // d_prob = masked w * g_semantics (semantic weights are detached: no contribution to d_sigma)
  if (VEC4) {
    const uint32_t g = c / 4, spi = 32 / g;
    const uint32_t sub = lane / g, grp = lane % g;
    if (static_cast<uint32_t>(lane) < spi * g) {
      const float4 gs = *reinterpret_cast<const float4*>(g_sem + static_cast<uint64_t>(n) * c + 4 * grp);
      float4* p = reinterpret_cast<float4*>(d_prob + row * c);
      for (uint32_t s = sub; s < t; s += spi) {
        const float w = wm[s];
        p[static_cast<size_t>(s) * g + grp] = make_float4(w * gs.x, w * gs.y, w * gs.z, w * gs.w);
      }
    }
  } else {
    float* p = d_prob + row * c;
    for (uint32_t cls = lane; cls < c; cls += 32) {
      const float gs = g_sem[static_cast<uint64_t>(n) * c + cls];
      for (uint32_t s = 0; s < t; ++s) p[static_cast<size_t>(s) * c + cls] = wm[s] * gs;
    }
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

static int dense_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) {
      set_error("composite_dense: T too large for shared-memory staging");
      return UCSA_ERR_UNSUPPORTED;
    }
    return set_max_dyn_smem(fn, bytes, "composite_dense");
  }
  return UCSA_OK;
}

extern "C" int ucsa_composite_dense_fwd(const float* sigma, const float* z, const float* rgb, const float* prob,
                                        const float* direction_norms, uint32_t n_rays, uint32_t t,
                                        uint32_t n_classes, float density_scale, float* weights, float* depth,
                                        float* image, float* semantics, void* stream) {
  UCSA_REQUIRE(sigma && z && rgb && prob && direction_norms && depth && image && semantics,
               "composite_dense_fwd: null pointer");
  UCSA_REQUIRE(t >= 1 && n_classes >= 1, "composite_dense_fwd: T and C must be >= 1");
  if (n_rays == 0) return UCSA_OK;
  const size_t smem = static_cast<size_t>(kWarpsPerCta) * 3 * t * sizeof(float);
  const bool vec4 = n_classes % 4 == 0 && n_classes <= 128 && (reinterpret_cast<uintptr_t>(prob) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(semantics) % 16 == 0);
  const dim3 grid(ceil_div(n_rays, kWarpsPerCta)), block(32 * kWarpsPerCta);
  // TMA-staged kernel (one CTA per ray): rows of every input must be 16-byte multiples at 16-byte aligned addresses
  // (bulk copies), and the ray's stage must fit the CTA's shared memory
  {
    DenseStage lay;
    lay.z_off = 0;
    lay.sigma_off = 4 * t;
    lay.rgb_off = 8 * t;
    lay.prob_off = 20 * t;
    lay.bytes = 20 * t + 4 * t * n_classes;
    const uint32_t g4 = n_classes / 4;
    const size_t tma_smem = lay.bytes + ((4 * t + 15) & ~15u) + (g4 ? (kTmaThreads / g4) * g4 * 16 : 0) + 32 + 16;
    const auto aligned = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    static int use_tma = -1;
    if (use_tma < 0) {  // bring-up knob: UCSA_DENSE_TMA=0 forces the load/compute kernel
      const char* e = getenv("UCSA_DENSE_TMA");
      use_tma = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    if (use_tma && vec4 && t % 4 == 0 && g4 >= 1 && g4 <= 32 && n_classes <= kTmaThreads && aligned(sigma) &&
        aligned(z) && aligned(rgb) && tma_smem <= 200 * 1024 && lay.bytes < (1u << 20)) {
      const void* fn = reinterpret_cast<const void*>(composite_dense_fwd_tma_kernel);
      if (int rc = set_max_dyn_smem(fn, tma_smem, "composite_dense_fwd")) return rc;
      composite_dense_fwd_tma_kernel<<<n_rays, kTmaThreads, tma_smem, as_stream(stream)>>>(
          sigma, z, rgb, prob, direction_norms, t, n_classes, density_scale, weights, depth, image, semantics, lay);
      return check_launch("composite_dense_fwd");
    }
  }
  if (vec4) {
    if (int rc = dense_smem(reinterpret_cast<const void*>(composite_dense_fwd_kernel<true>), smem)) return rc;
    composite_dense_fwd_kernel<true><<<grid, block, smem, as_stream(stream)>>>(
        sigma, z, rgb, prob, direction_norms, n_rays, t, n_classes, density_scale, weights, depth, image,
        semantics);
  } else {
    if (int rc = dense_smem(reinterpret_cast<const void*>(composite_dense_fwd_kernel<false>), smem)) return rc;
    composite_dense_fwd_kernel<false><<<grid, block, smem, as_stream(stream)>>>(
        sigma, z, rgb, prob, direction_norms, n_rays, t, n_classes, density_scale, weights, depth, image,
        semantics);
  }
  return check_launch("composite_dense_fwd");
}

extern "C" int ucsa_composite_dense_bwd(const float* sigma, const float* z, const float* rgb, const float* weights,
                                        const float* direction_norms, const float* g_depth, const float* g_image,
                                        const float* g_semantics, uint32_t n_rays, uint32_t t, uint32_t n_classes,
                                        float density_scale, float* d_sigma, float* d_rgb, float* d_prob,
                                        void* stream) {
  UCSA_REQUIRE(sigma && z && rgb && weights && direction_norms && g_depth && g_image && g_semantics && d_sigma &&
                   d_rgb && d_prob,
               "composite_dense_bwd: null pointer");
  if (n_rays == 0) return UCSA_OK;
  const size_t smem = static_cast<size_t>(kWarpsPerCta) * 4 * t * sizeof(float);
  const bool vec4 = n_classes % 4 == 0 && n_classes <= 128 && (reinterpret_cast<uintptr_t>(d_prob) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(g_semantics) % 16 == 0);
  const dim3 grid(ceil_div(n_rays, kWarpsPerCta)), block(32 * kWarpsPerCta);
  if (vec4) {
    if (int rc = dense_smem(reinterpret_cast<const void*>(composite_dense_bwd_kernel<true>), smem)) return rc;
    composite_dense_bwd_kernel<true><<<grid, block, smem, as_stream(stream)>>>(
        sigma, z, rgb, weights, direction_norms, g_depth, g_image, g_semantics, n_rays, t, n_classes,
        density_scale, d_sigma, d_rgb, d_prob);
  } else {
    if (int rc = dense_smem(reinterpret_cast<const void*>(composite_dense_bwd_kernel<false>), smem)) return rc;
    composite_dense_bwd_kernel<false><<<grid, block, smem, as_stream(stream)>>>(
        sigma, z, rgb, weights, direction_norms, g_depth, g_image, g_semantics, n_rays, t, n_classes,
        density_scale, d_sigma, d_rgb, d_prob);
  }
  return check_launch("composite_dense_bwd");
}
