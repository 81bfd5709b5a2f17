// Weights, w > 1e-4 masks, depth, and the compaction of the masked-in samples that the colour / semantic
// heads are evaluated on.  Rows a11 (+ the masks of a12/a13) and their backward (a15) of SURVEY.md section 8.
#include "weights.cuh"

namespace ucsa {
namespace {

constexpr int kWarpsPerCta = 4;

// gather a ray's (z, sigma) into sorted order in the warp's shared-memory slice.  The loads of kStageBatch 32-sample
// chunks are issued together (order first, then the two dependent gathers): one chunk at a time the 16 chunks of a
// 512-sample ray were 32 dependent global round trips, most of the run time of the kernels below.
// `slots` (optional): the cat slot of every sorted position, for callers that need it again.
constexpr int kStageBatch = 8;
__device__ __forceinline__ void stage_sorted(const float* __restrict__ z_cat, const float* __restrict__ sigma,
                                             const int32_t* __restrict__ order, uint64_t row, uint32_t t,
                                             int lane, float* zs, float* sg, uint32_t* slots = nullptr) {
  for (uint32_t s0 = lane; s0 < t; s0 += 32 * kStageBatch) {
    uint32_t slot[kStageBatch];
    float zv[kStageBatch], sv[kStageBatch];
#pragma unroll
    for (int i = 0; i < kStageBatch; ++i) {
      const uint32_t s = s0 + 32 * i;
      slot[i] = s < t ? (order != nullptr ? static_cast<uint32_t>(order[row + s]) : s) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kStageBatch; ++i) {
      const bool in = s0 + 32 * i < t;
      zv[i] = in ? z_cat[row + slot[i]] : 0.f;
      sv[i] = in ? sigma[row + slot[i]] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kStageBatch; ++i) {
      const uint32_t s = s0 + 32 * i;
      if (s < t) {
        zs[s] = zv[i];
        sg[s] = sv[i];
        if (slots != nullptr) slots[s] = slot[i];
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
weights_fwd_kernel(const float* __restrict__ z_cat, const float* __restrict__ sigma,
                   const int32_t* __restrict__ order, const float* __restrict__ dnorm, uint32_t n_rays, uint32_t t,
                   float density_scale, float* __restrict__ w_sorted, float* __restrict__ depth,
                   int32_t* __restrict__ ray_count, uint8_t* __restrict__ use_geo) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * kWarpsPerCta + wib;
  if (n >= n_rays) return;
  float* zs = sm + static_cast<size_t>(wib) * 2 * t;
  float* sg = zs + t;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  stage_sorted(z_cat, sigma, order, row, t, lane, zs, sg);

  float carry = 1.0f, dsum = 0.f;
  int count = 0;
  for (uint32_t base = 0; base < t; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < t;
    SampleTerms st{1.f, 0.f, 1.f, 0.f};
    if (valid) st = sample_terms(zs, sg, s, t, density_scale);
    const float trans = chunk_transmittance(valid ? st.keep : 1.0f, carry, lane);
    const float w = st.alpha * trans;
    const bool keep = valid && w > kMaskThreshold;
    if (valid) {
      w_sorted[row + s] = w;
      const uint32_t slot = order != nullptr ? static_cast<uint32_t>(order[row + s]) : s;
      use_geo[row + slot] = keep ? 1 : 0;
    }
    if (keep) dsum += w * zs[s];
    count += __popc(__ballot_sync(kFullMask, keep));
  }
  dsum = warp_sum(dsum);
  if (lane == 0) {
    depth[n] = dsum / dnorm[n];
    ray_count[n] = count;
  }
}

// single-CTA exclusive scan; n_rays <= 2^18 in every configuration, i.e. at most 256 iterations
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t* __restrict__ counts, uint32_t n, int32_t* __restrict__ offsets) {
  __shared__ int warp_total[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    int incl = warp_scan_add_i32(v, lane);
    if (lane == 31) warp_total[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int tot = warp_total[lane];
      const int sc = warp_scan_add_i32(tot, lane);
      warp_total[lane] = sc - tot;  // exclusive offset of each warp
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < n) offsets[i] = carry + warp_total[wid] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_total[wid] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry_s;
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
compact_masked_kernel(const float* __restrict__ w_sorted, const float* __restrict__ z_cat,
                      const int32_t* __restrict__ order, const int32_t* __restrict__ ray_off, uint32_t n_rays,
                      uint32_t t, int32_t* __restrict__ sel, float* __restrict__ w_sel,
                      float* __restrict__ z_sel) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (n >= n_rays) return;
  const uint64_t row = static_cast<uint64_t>(n) * t;
  int out = ray_off[n];
  for (uint32_t base = 0; base < t; base += 32) {
    const uint32_t s = base + lane;
    const float w = s < t ? w_sorted[row + s] : 0.f;
    const bool keep = s < t && w > kMaskThreshold;
    const unsigned ballot = __ballot_sync(kFullMask, keep);
    if (keep) {
      const int r = out + __popc(ballot & ((1u << lane) - 1u));
      const uint32_t slot = order != nullptr ? static_cast<uint32_t>(order[row + s]) : s;
      sel[r] = static_cast<int32_t>(row + slot);
      w_sel[r] = w;
      z_sel[r] = z_cat[row + slot];
    }
    out += __popc(ballot);
  }
}

// weights_fwd + scan_counts + compact_masked in ONE pass: a warp computes its ray's weights (kept in shared memory),
// the per-ray counts become global offsets through a decoupled look-back over the CTAs (single pass, no second read of
// the weights, no single-CTA scan), and the masked-in samples are written at their final positions.  CTAs take their
// position in the scan from a ticket counter, so a CTA only ever waits on CTAs that are already running.
// scratch: [0] ticket, [1] finished CTAs, [2..] one 64-bit status word per CTA (bits 63..62: 0 empty, 1 = this CTA's
// count, 2 = inclusive prefix; low 32 bits the value).  All zero on entry; the last CTA to finish zeroes it again.
constexpr unsigned long long kStatAgg = 1ull << 62, kStatIncl = 2ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
weights_compact_kernel(const float* __restrict__ z_cat, const float* __restrict__ sigma,
                       const int32_t* __restrict__ order, const float* __restrict__ dnorm, uint32_t n_rays, uint32_t t,
                       float density_scale, float* __restrict__ w_sorted, float* __restrict__ depth,
                       int32_t* __restrict__ ray_off, uint8_t* __restrict__ use_geo, int32_t* __restrict__ sel,
                       float* __restrict__ w_sel, float* __restrict__ z_sel, uint32_t* __restrict__ scratch) {
  extern __shared__ float sm[];
  __shared__ uint32_t vid_s;
  __shared__ int counts_s[kWarpsPerCta];
  __shared__ int base_s;
  __shared__ int last_s;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(scratch + 2);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x == 0) vid_s = atomicAdd(scratch, 1u);
  __syncthreads();
  const uint32_t vid = vid_s;
  const uint32_t n = vid * kWarpsPerCta + wib;
  const bool has_ray = n < n_rays;
  float* zs = sm + static_cast<size_t>(wib) * 3 * t;
  float* sg = zs + t;  // sigma, then the weight of the same sample
  uint32_t* slots = reinterpret_cast<uint32_t*>(sg + t);  // cat slot of every sorted position
  const uint64_t row = static_cast<uint64_t>(n) * t;
  int count = 0;
  if (has_ray) {
    stage_sorted(z_cat, sigma, order, row, t, lane, zs, sg, slots);
    float carry = 1.0f, dsum = 0.f;
    for (uint32_t base = 0; base < t; base += 32) {
      const uint32_t s = base + lane;
      const bool valid = s < t;
      SampleTerms st{1.f, 0.f, 1.f, 0.f};
      if (valid) st = sample_terms(zs, sg, s, t, density_scale);
      const float trans = chunk_transmittance(valid ? st.keep : 1.0f, carry, lane);
      const float w = st.alpha * trans;
      const bool keep = valid && w > kMaskThreshold;
      if (valid) {
        w_sorted[row + s] = w;
        sg[s] = w;  // sigma of this sample is not read again (a lane reads sg only at its own s)
        use_geo[row + slots[s]] = keep ? 1 : 0;
      }
      if (keep) dsum += w * zs[s];
      count += __popc(__ballot_sync(kFullMask, keep));
    }
    dsum = warp_sum(dsum);
    if (lane == 0) depth[n] = dsum / dnorm[n];
  }
  if (lane == 0) counts_s[wib] = count;
  __syncthreads();
  if (wib == 0) {
    int agg = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) agg += counts_s[w];
    int excl = 0;
    if (vid > 0) {
      if (lane == 0) st_status(status + vid, kStatAgg | static_cast<uint32_t>(agg));
      int64_t end = static_cast<int64_t>(vid) - 1;  // look back over windows of 32 predecessors
      while (true) {
        const int64_t j = end - lane;
        unsigned long long v = kStatIncl;  // before the first CTA: inclusive prefix 0
        if (j >= 0) {
          do { v = ld_status(status + j); } while ((v >> 62) == 0ull);
        }
        const uint32_t incl = __ballot_sync(kFullMask, (v >> 62) == 2ull);
        const int val = static_cast<int>(static_cast<uint32_t>(v));
        if (incl != 0u) {  // nearest predecessor with a complete prefix: add it and everything in between
          const int stop = __ffs(incl) - 1;
          excl += __reduce_add_sync(kFullMask, lane <= stop ? val : 0);
          break;
        }
        excl += __reduce_add_sync(kFullMask, val);
        end -= 32;
      }
    }
    if (lane == 0) {
      st_status(status + vid, kStatIncl | static_cast<uint32_t>(excl + agg));
      base_s = excl;
    }
  }
  __syncthreads();
  int out = base_s;
  for (int w = 0; w < wib; ++w) out += counts_s[w];
  if (has_ray) {
    if (lane == 0) {
      ray_off[n] = out;
      if (n + 1 == n_rays) ray_off[n_rays] = out + count;
    }
    for (uint32_t base = 0; base < t; base += 32) {
      const uint32_t s = base + lane;
      const float w = s < t ? sg[s] : 0.f;
      const bool keep = s < t && w > kMaskThreshold;
      const unsigned ballot = __ballot_sync(kFullMask, keep);
      if (keep) {
        const int r = out + __popc(ballot & ((1u << lane) - 1u));
        sel[r] = static_cast<int32_t>(row + slots[s]);
        w_sel[r] = w;
        z_sel[r] = zs[s];
      }
      out += __popc(ballot);
    }
  }
  // every CTA is past its look-back once it arrives here; the last one re-arms the scratch block for the next launch
  if (threadIdx.x == 0) last_s = atomicAdd(scratch + 1, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (last_s) {
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) st_status(status + i, 0ull);
    if (threadIdx.x == 0) {
      scratch[0] = 0u;
      scratch[1] = 0u;
    }
  }
}

// d_sigma (cat order) from dL/dw of the masked-in samples (d_w_sel, compact order).
__global__ void __launch_bounds__(32 * kWarpsPerCta)
weights_bwd_kernel(const float* __restrict__ z_cat, const float* __restrict__ sigma,
                   const int32_t* __restrict__ order, const float* __restrict__ w_sorted,
                   const int32_t* __restrict__ ray_off, const float* __restrict__ d_w_sel, uint32_t n_rays,
                   uint32_t t, float density_scale, float* __restrict__ d_sigma) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * kWarpsPerCta + wib;
  if (n >= n_rays) return;
  // Three staged arrays per ray (z, sigma, dL/dw) + one transmittance carry per 32-sample chunk: 6 KB per ray at
  // T = 512, so that all 4096 rays of a training batch are resident in ONE wave (five arrays -- with the weights and
  // the transmittances staged as well -- were 1.4 waves).  The weights are recomputed (alpha * T, the very operations
  // of weights_fwd, hence the same bits) and the transmittance of a chunk is re-scanned from its stored carry.
  const uint32_t n_chunks = (t + 31) / 32;
  float* zs = sm + static_cast<size_t>(wib) * (3 * t + n_chunks);
  float* sg = zs + t;
  float* gw = sg + t;            // dL/dw (0 where masked out)
  float* chunk_carry = gw + t;   // transmittance in front of each chunk
  const uint64_t row = static_cast<uint64_t>(n) * t;
  stage_sorted(z_cat, sigma, order, row, t, lane, zs, sg);

  float carry = 1.0f;
  int out = ray_off[n];
  for (uint32_t base = 0; base < t; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < t;
    SampleTerms st{1.f, 0.f, 1.f, 0.f};
    if (valid) st = sample_terms(zs, sg, s, t, density_scale);
    if (lane == 0) chunk_carry[base >> 5] = carry;
    const float trans = chunk_transmittance(valid ? st.keep : 1.0f, carry, lane);
    const bool keep = valid && st.alpha * trans > kMaskThreshold;  // = the weight weights_fwd stored
    const unsigned ballot = __ballot_sync(kFullMask, keep);
    if (valid) gw[s] = __int_as_float(keep ? out + __popc(ballot & ((1u << lane) - 1u)) : -1);  // compact row, for now
    out += __popc(ballot);
  }
  __syncwarp();
#pragma unroll 4
  for (uint32_t s = lane; s < t; s += 32) {  // independent gathers: several in flight per lane
    const int idx = __float_as_int(gw[s]);
    gw[s] = idx >= 0 ? d_w_sel[idx] : 0.f;
  }
  __syncwarp();
  float suffix_carry = 0.f;
  for (uint32_t c = n_chunks; c-- > 0;) {
    const uint32_t s = c * 32 + lane;
    const bool valid = s < t;
    SampleTerms st{1.f, 0.f, 1.f, 0.f};
    if (valid) st = sample_terms(zs, sg, s, t, density_scale);
    float carry_c = chunk_carry[c];
    const float trans = chunk_transmittance(valid ? st.keep : 1.0f, carry_c, lane);
    const float g = valid ? gw[s] : 0.f;
    const float w = valid ? st.alpha * trans : 0.f;
    const float suffix = chunk_suffix(g * w, suffix_carry, lane);
    if (valid) {
      const uint32_t slot = order != nullptr ? static_cast<uint32_t>(order[row + s]) : s;
      d_sigma[row + slot] = sigma_grad(st, density_scale, g, trans, suffix);
    }
  }
}

int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) {
      set_error("samples per ray too large for the shared-memory staging (%zu bytes)", bytes);
      return UCSA_ERR_UNSUPPORTED;
    }
    return set_max_dyn_smem(fn, bytes, "weights");
  }
  return UCSA_OK;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_weights_fwd(const float* z_cat, const float* sigma, const int32_t* order,
                                const float* direction_norms, uint32_t n_rays, uint32_t t, float density_scale,
                                float* w_sorted, float* depth, int32_t* ray_count, uint8_t* use_geo,
                                void* stream) {
  UCSA_REQUIRE(z_cat && sigma && direction_norms && w_sorted && depth && ray_count && use_geo,
               "weights_fwd: null pointer");
  UCSA_REQUIRE(t >= 1, "weights_fwd: T must be >= 1");
  if (n_rays == 0) return UCSA_OK;
  const size_t smem = static_cast<size_t>(kWarpsPerCta) * 2 * t * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(weights_fwd_kernel), smem)) return rc;
  weights_fwd_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, smem, as_stream(stream)>>>(
      z_cat, sigma, order, direction_norms, n_rays, t, density_scale, w_sorted, depth, ray_count, use_geo);
  return check_launch("weights_fwd");
}

extern "C" int ucsa_scan_counts(const int32_t* ray_count, uint32_t n_rays, int32_t* ray_off, void* stream) {
  UCSA_REQUIRE(ray_count && ray_off, "scan_counts: null pointer");
  scan_counts_kernel<<<1, 1024, 0, as_stream(stream)>>>(ray_count, n_rays, ray_off);
  return check_launch("scan_counts");
}

extern "C" int ucsa_compact_masked(const float* w_sorted, const float* z_cat, const int32_t* order,
                                   const int32_t* ray_off, uint32_t n_rays, uint32_t t, int32_t* sel,
                                   float* w_sel, float* z_sel, void* stream) {
  UCSA_REQUIRE(w_sorted && z_cat && ray_off && sel && w_sel && z_sel, "compact_masked: null pointer");
  UCSA_REQUIRE(static_cast<uint64_t>(n_rays) * t < (1ull << 31), "compact_masked: N*T must fit int32");
  if (n_rays == 0) return UCSA_OK;
  compact_masked_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, 0, as_stream(stream)>>>(
      w_sorted, z_cat, order, ray_off, n_rays, t, sel, w_sel, z_sel);
  return check_launch("compact_masked");
}

extern "C" int ucsa_weights_compact(const float* z_cat, const float* sigma, const int32_t* order,
                                    const float* direction_norms, uint32_t n_rays, uint32_t t, float density_scale,
                                    float* w_sorted, float* depth, int32_t* ray_off, uint8_t* use_geo, int32_t* sel,
                                    float* w_sel, float* z_sel, uint32_t* scratch, void* stream) {
  UCSA_REQUIRE(z_cat && sigma && direction_norms && w_sorted && depth && ray_off && use_geo && sel && w_sel && z_sel &&
                   scratch, "weights_compact: null pointer");
  UCSA_REQUIRE(t >= 1, "weights_compact: T must be >= 1");
  UCSA_REQUIRE(static_cast<uint64_t>(n_rays) * t < (1ull << 31), "weights_compact: N*T must fit int32");
  UCSA_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 8 == 0, "weights_compact: scratch must be 8-byte aligned");
  if (n_rays == 0) return UCSA_OK;
  const size_t smem = static_cast<size_t>(kWarpsPerCta) * 3 * t * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(weights_compact_kernel), smem)) return rc;
  weights_compact_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, smem, as_stream(stream)>>>(
      z_cat, sigma, order, direction_norms, n_rays, t, density_scale, w_sorted, depth, ray_off, use_geo, sel, w_sel,
      z_sel, scratch);
  return check_launch("weights_compact");
}

extern "C" int ucsa_weights_bwd(const float* z_cat, const float* sigma, const int32_t* order,
                                const float* w_sorted, const int32_t* ray_off, const float* d_w_sel,
                                uint32_t n_rays, uint32_t t, float density_scale, float* d_sigma, void* stream) {
  UCSA_REQUIRE(z_cat && sigma && w_sorted && ray_off && d_w_sel && d_sigma, "weights_bwd: null pointer");
  if (n_rays == 0) return UCSA_OK;
  const size_t smem = static_cast<size_t>(kWarpsPerCta) * (3 * t + (t + 31) / 32) * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(weights_bwd_kernel), smem)) return rc;
  weights_bwd_kernel<<<ceil_div(n_rays, kWarpsPerCta), 32 * kWarpsPerCta, smem, as_stream(stream)>>>(
      z_cat, sigma, order, w_sorted, ray_off, d_w_sel, n_rays, t, density_scale, d_sigma);
  return check_launch("weights_bwd");
}
