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
// TMA-staged forward (the production path when a ray's inputs fit a stage): the kernel above leaves every warp with a
// handful of dependent load rounds (z / sigma -> weights -> rgb -> probabilities, ~1 us of HBM latency each), which is
// what bounds a 35 us kernel.  Here a PERSISTENT warp walks its rays with a two-stage ring in shared memory: one lane
// requests ALL inputs of the next ray (z, sigma, rgb and the [T, C] probability block, four bulk copies completing on
// one mbarrier) before the warp starts on the current ray, so every byte is in flight one ray ahead and the
// arithmetic reads shared memory only.  23 KB per stage at T = 128, C = 40; 4 warps x 2 stages per SM keep ~92 KB
// in flight per SM, twice what HBM latency x bandwidth needs.  Same arithmetic, in the same order, as the kernel above.
struct DenseStage {
  uint32_t z_off, sigma_off, rgb_off, prob_off, bytes;  // byte offsets inside a stage
};

__global__ void __launch_bounds__(32 * kWarpsPerCta)
composite_dense_fwd_tma_kernel(const float* __restrict__ sigma, const float* __restrict__ z,
                               const float* __restrict__ rgb, const float* __restrict__ prob,
                               const float* __restrict__ dnorm, uint32_t n_rays, uint32_t t, uint32_t c,
                               float density_scale, float* __restrict__ weights, float* __restrict__ depth,
                               float* __restrict__ image, float* __restrict__ semantics, DenseStage lay,
                               uint32_t warp_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* mine = smem_raw + static_cast<size_t>(wib) * warp_bytes;  // [stage 0 | stage 1 | wm | 2 barriers]
  unsigned char* stage_base[2] = {mine, mine + lay.bytes};
  float* wm = reinterpret_cast<float*>(mine + 2 * lay.bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(mine + 2 * lay.bytes + ((4 * t + 15) & ~15u));
  const uint32_t stride = gridDim.x * kWarpsPerCta;
  uint32_t n = blockIdx.x * kWarpsPerCta + wib;
  if (n >= n_rays) return;  // (whole warps only: no CTA-wide barrier below)
  const uint64_t stream = l2_policy_stream();
  if (lane == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
  }
  __syncwarp();

  auto request = [&](uint32_t ray, int st) {  // lane 0 only
    const uint64_t row = static_cast<uint64_t>(ray) * t;
    const uint32_t dst = umma::smem_u32(stage_base[st]);
    umma::mbar_expect_tx(&bars[st], lay.bytes);
    umma::bulk_load(dst + lay.z_off, z + row, 4 * t, &bars[st], stream);
    umma::bulk_load(dst + lay.sigma_off, sigma + row, 4 * t, &bars[st], stream);
    umma::bulk_load(dst + lay.rgb_off, rgb + row * 3, 12 * t, &bars[st], stream);
    umma::bulk_load(dst + lay.prob_off, prob + row * c, 4 * t * c, &bars[st], stream);
  };
  if (lane == 0) request(n, 0);
  uint32_t phase[2] = {0u, 0u};
  const uint32_t g = c / 4;     // float4 groups per sample
  const uint32_t spi = 32 / g;  // samples per warp step
  const uint32_t sub = lane / g, grp = lane % g;
  const bool active = static_cast<uint32_t>(lane) < spi * g;

  for (int st = 0; n < n_rays; n += stride, st ^= 1) {
    if (lane == 0 && n + stride < n_rays) {
      umma::fence_async_smem();  // the other stage was read by this warp (generic proxy) one ray ago
      request(n + stride, st ^ 1);
    }
    umma::mbar_wait(&bars[st], phase[st]);
    phase[st] ^= 1u;
    const float* zs = reinterpret_cast<const float*>(stage_base[st] + lay.z_off);
    const float* sg = reinterpret_cast<const float*>(stage_base[st] + lay.sigma_off);
    const float* cs = reinterpret_cast<const float*>(stage_base[st] + lay.rgb_off);
    const float4* ps = reinterpret_cast<const float4*>(stage_base[st] + lay.prob_off);
    const uint64_t row = static_cast<uint64_t>(n) * t;
    const float dsum = ray_weights(zs, sg, t, density_scale, lane, wm, weights ? weights + row : nullptr);
    if (lane == 0) depth[n] = dsum / dnorm[n];
    __syncwarp();
    {  // colour: lanes 0..29 = 10 samples x 3 channels per step
      const int csub = lane / 3, ch = lane % 3;
      float acc = 0.f;
      if (lane < 30)
        for (uint32_t s = csub; s < t; s += 10) acc = fmaf(wm[s], cs[s * 3 + ch], acc);
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) tot += __shfl_sync(kFullMask, acc, (3 * k + ch) % 32);
      if (lane < 3) image[static_cast<uint64_t>(n) * 3 + lane] = tot;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
#pragma unroll 4
      for (uint32_t s = sub; s < t; s += spi) {
        const float4 v = ps[static_cast<size_t>(s) * g + grp];
        const float w = wm[s];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    }
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
    __syncwarp();  // every lane is done with this stage and with wm before they are refilled
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
  // TMA-staged persistent kernel: rows of every input must be 16-byte multiples at 16-byte aligned addresses (bulk
  // copies), and two stages per warp must fit the CTA's shared memory
  {
    DenseStage lay;
    lay.z_off = 0;
    lay.sigma_off = 4 * t;
    lay.rgb_off = 8 * t;
    lay.prob_off = 20 * t;
    lay.bytes = 20 * t + 4 * t * n_classes;
    const uint32_t warp_bytes = (2 * lay.bytes + ((4 * t + 15) & ~15u) + 16 + 127) & ~127u;
    const size_t tma_smem = static_cast<size_t>(kWarpsPerCta) * warp_bytes;
    const auto aligned = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    static int use_tma = -1;
    if (use_tma < 0) {  // bring-up knob: UCSA_DENSE_TMA=0 forces the load/compute kernel
      const char* e = getenv("UCSA_DENSE_TMA");
      use_tma = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    if (use_tma && vec4 && t % 4 == 0 && 32 / (n_classes / 4) >= 1 && aligned(sigma) && aligned(z) && aligned(rgb) &&
        tma_smem <= 227 * 1024 && lay.bytes < (1u << 20)) {
      const void* fn = reinterpret_cast<const void*>(composite_dense_fwd_tma_kernel);
      if (int rc = set_max_dyn_smem(fn, tma_smem, "composite_dense_fwd")) return rc;
      const uint32_t per_sm = static_cast<uint32_t>((227 * 1024) / (tma_smem + 1024));
      const uint32_t cap = kNumSMs * (per_sm < 1 ? 1 : per_sm);
      const uint32_t want = ceil_div(n_rays, kWarpsPerCta);
      composite_dense_fwd_tma_kernel<<<want < cap ? want : cap, block, tma_smem, as_stream(stream)>>>(
          sigma, z, rgb, prob, direction_norms, n_rays, t, n_classes, density_scale, weights, depth, image,
          semantics, lay, warp_bytes);
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
