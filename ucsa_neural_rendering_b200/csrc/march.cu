// Occupancy-grid ray marching, ragged compositing and the inference wavefront (rows a16-a20 of SURVEY.md section 8).
// Replaces the kernels of nr4seg/nerf/raymarching/src/raymarching.cu:138-864 and pcg32.h.
//
// Differences from the reference kernels, all of them within what the reference itself can produce:
//   * sample offsets are handed out in ray order by a count -> scan -> write sequence (warp-shuffle scan), not by
//     atomicAdd, so the packed streams are deterministic (the reference's order depends on atomic arrival);
//   * the occupancy test can read a bitfield (1 bit per cell, 786 KB for 3 x 128^3, L1/L2 resident) produced by
//     ucsa_grid_packbits from the float grid with the very comparison of the reference (density > min(0.01, mean)),
//     so the marched samples are identical to marching the float grid;
//   * live-ray compaction keeps ray order (ballot + prefix), the reference's is atomic-order.
// The per-sample arithmetic (clamps, frexpf mip level, the double-literal index expression, FMA contractions) is
// written as in the reference so that positions and step sizes are bit-identical.
#include "common.cuh"

namespace ucsa {
namespace {

constexpr int kMaxSteps = 1024;
constexpr float kSqrt3 = 1.73205080757f;
constexpr float kDensityThresh = 0.01f;
__host__ __device__ inline float min_stepsize() { return 2 * kSqrt3 / kMaxSteps; }

// ------------------------------------------------------------------ PCG32 (pcg32.h:44-116)
struct Pcg32 {
  uint64_t state, inc;
  __device__ void seed(uint64_t initstate, uint64_t initseq) {
    state = 0u;
    inc = (initseq << 1u) | 1u;
    next_uint();
    state += initstate;
    next_uint();
  }
  __device__ uint32_t next_uint() {
    const uint64_t old = state;
    state = old * 0x5851f42d4c957f2dULL + inc;
    const uint32_t xorshifted = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = static_cast<uint32_t>(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
  }
  __device__ float next_float() { return __uint_as_float((next_uint() >> 9) | 0x3f800000u) - 1.0f; }
};

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float signf(float x) { return copysignf(1.0f, x); }

__device__ __forceinline__ int mip_from_pos(float x, float y, float z, float max_cascade) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int e;
  frexpf(mx, &e);
  return static_cast<int>(fminf(max_cascade - 1, fmaxf(0.f, static_cast<float>(e))));
}

struct Occupancy {
  const float* grid;        // [C,H,H,H] float densities, or
  const uint32_t* bits;     // packed bitfield (bit i of word i/32 = cell i occupied)
  float thresh;
  __device__ __forceinline__ bool occupied(uint32_t index) const {
    if (bits != nullptr) return (__ldg(bits + (index >> 5)) >> (index & 31)) & 1u;
    return __ldg(grid + index) > thresh;
  }
};

struct Ray {
  float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, far;
};

struct March {
  float bound, dt_gamma, dt_min, dt_max;
  uint32_t C, H;
  Occupancy occ;

  // one probe at parameter t (raymarching.cu:187-226): occupied -> one step, else jump to the next voxel
  __device__ __forceinline__ bool probe(const Ray& r, float& t, float& x, float& y, float& z, float& dt) const {
    x = clampf(r.ox + t * r.dx, -bound, bound);
    y = clampf(r.oy + t * r.dy, -bound, bound);
    z = clampf(r.oz + t * r.dz, -bound, bound);
    const int level = mip_from_pos(x, y, z, static_cast<float>(C));
    const float mip_bound = fminf(exp2f(static_cast<float>(level)), bound);
    const float mip_rbound = 1 / mip_bound;
    const int nx = clampf(0.5 * (x * mip_rbound + 1) * H, 0.0f, static_cast<float>(H - 1));
    const int ny = clampf(0.5 * (y * mip_rbound + 1) * H, 0.0f, static_cast<float>(H - 1));
    const int nz = clampf(0.5 * (z * mip_rbound + 1) * H, 0.0f, static_cast<float>(H - 1));
    const uint32_t index = level * H * H * H + nx * H * H + ny * H + nz;
    if (occ.occupied(index)) {
      dt = clampf(t * dt_gamma, dt_min, dt_max);
      t += dt;
      return true;
    }
    const float tx = (((nx + 0.5f + 0.5f * signf(r.dx)) / (H - 1) * 2 - 1) * mip_bound - x) * r.rdx;
    const float ty = (((ny + 0.5f + 0.5f * signf(r.dy)) / (H - 1) * 2 - 1) * mip_bound - y) * r.rdy;
    const float tz = (((nz + 0.5f + 0.5f * signf(r.dz)) / (H - 1) * 2 - 1) * mip_bound - z) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
      const float step = clampf(t * dt_gamma, dt_min, dt_max);
      t += step;
    } while (t < tt);
    return false;
  }
};

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t n,
                                        float far) {
  Ray r;
  r.ox = rays_o[3 * n], r.oy = rays_o[3 * n + 1], r.oz = rays_o[3 * n + 2];
  r.dx = rays_d[3 * n], r.dy = rays_d[3 * n + 1], r.dz = rays_d[3 * n + 2];
  r.rdx = 1 / r.dx, r.rdy = 1 / r.dy, r.rdz = 1 / r.dz;
  r.far = far;
  return r;
}

__device__ __forceinline__ float train_t0(float near, uint32_t n, uint32_t perturb) {
  float t0 = near;
  if (perturb) {
    Pcg32 rng;
    rng.seed(static_cast<uint64_t>(n), 1u);
    t0 += min_stepsize() * rng.next_float();
  }
  return t0;
}

// pass 1: number of occupied steps per ray.  The march is one dependent chain per ray (t += dt, the next probe
// depends on it) and must stay one to remain bit-identical with the reference; what need not be repeated is the
// chain itself: with a staging buffer (t_stage [n_rays, kMaxSteps]) this pass also records the parameter t of every
// occupied step (a store off the dependent chain), and the samples are then written by a pass that is parallel over
// the samples (march_write_staged_kernel) instead of marching every ray a second time.
__global__ void march_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, March m,
                                   uint32_t n_rays, const float* __restrict__ nears, const float* __restrict__ fars,
                                   uint32_t perturb, int32_t* __restrict__ counts, float* __restrict__ t_stage) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rays) return;
  const Ray r = load_ray(rays_o, rays_d, n, fars[n]);
  float t = train_t0(nears[n], n, perturb), x, y, z, dt;
  uint32_t steps = 0;
  if (t_stage != nullptr) {
    float* ts = t_stage + static_cast<uint64_t>(n) * kMaxSteps;
    while (t < r.far && steps < kMaxSteps) {
      const float t_probe = t;
      if (m.probe(r, t, x, y, z, dt)) ts[steps++] = t_probe;
    }
  } else {
    while (t < r.far && steps < kMaxSteps) steps += m.probe(r, t, x, y, z, dt) ? 1u : 0u;
  }
  counts[n] = static_cast<int32_t>(steps);
}

// pass 2 with staged parameters: one warp per ray, lanes = consecutive samples.  Position, step size and the distance
// to the previous sample are the expressions of March::probe / march_write_kernel evaluated at the recorded t.
__global__ void march_write_staged_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, March m,
                                          uint32_t n_rays, uint32_t max_points, const float* __restrict__ nears,
                                          uint32_t perturb, const int32_t* __restrict__ counts,
                                          const int32_t* __restrict__ offsets, const float* __restrict__ t_stage,
                                          float* __restrict__ xyzs, float* __restrict__ dirs,
                                          float* __restrict__ deltas, int32_t* __restrict__ rays) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= n_rays) return;
  const uint32_t num_steps = static_cast<uint32_t>(counts[n]);
  const uint32_t point_index = static_cast<uint32_t>(offsets[n]);
  if (lane == 0) {
    rays[n * 3] = static_cast<int32_t>(n);
    rays[n * 3 + 1] = static_cast<int32_t>(point_index);
    rays[n * 3 + 2] = static_cast<int32_t>(num_steps);
  }
  if (num_steps == 0 || point_index + num_steps >= max_points) return;
  const Ray r = load_ray(rays_o, rays_d, n, 0.f);
  const float* ts = t_stage + static_cast<uint64_t>(n) * kMaxSteps;
  float last_carry = train_t0(nears[n], n, perturb);  // "last_t" in front of the ray's first sample
  for (uint32_t base = 0; base < num_steps; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < num_steps;
    const float t = valid ? ts[s] : 0.f;
    const float dt = clampf(t * m.dt_gamma, m.dt_min, m.dt_max);
    const float t_after = t + dt;
    float last_t = __shfl_up_sync(kFullMask, t_after, 1);
    if (lane == 0) last_t = last_carry;
    last_carry = __shfl_sync(kFullMask, t_after, 31);
    if (valid) {
      const uint64_t i = static_cast<uint64_t>(point_index) + s;
      float* px = xyzs + 3 * i;
      float* pd = dirs + 3 * i;
      px[0] = clampf(r.ox + t * r.dx, -m.bound, m.bound);
      px[1] = clampf(r.oy + t * r.dy, -m.bound, m.bound);
      px[2] = clampf(r.oz + t * r.dz, -m.bound, m.bound);
      pd[0] = r.dx, pd[1] = r.dy, pd[2] = r.dz;
      deltas[2 * i] = dt;
      deltas[2 * i + 1] = t_after - last_t;
    }
  }
}

// pass 2: write the samples of every ray at its scanned offset; rays[n] = (n, offset, count)
__global__ void march_write_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, March m,
                                   uint32_t n_rays, uint32_t max_points, const float* __restrict__ nears,
                                   const float* __restrict__ fars, uint32_t perturb, const int32_t* __restrict__ counts,
                                   const int32_t* __restrict__ offsets, int32_t base_point, int32_t base_ray,
                                   float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                                   int32_t* __restrict__ rays) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rays) return;
  const uint32_t num_steps = static_cast<uint32_t>(counts[n]);
  const uint32_t point_index = static_cast<uint32_t>(base_point + offsets[n]);
  const uint32_t ray_index = static_cast<uint32_t>(base_ray) + n;
  rays[ray_index * 3] = static_cast<int32_t>(n);
  rays[ray_index * 3 + 1] = static_cast<int32_t>(point_index);
  rays[ray_index * 3 + 2] = static_cast<int32_t>(num_steps);
  if (num_steps == 0 || point_index + num_steps >= max_points) return;
  const Ray r = load_ray(rays_o, rays_d, n, fars[n]);
  float* px = xyzs + 3ull * point_index;
  float* pd = dirs + 3ull * point_index;
  float* pl = deltas + 2ull * point_index;
  float t = train_t0(nears[n], n, perturb), x, y, z, dt;
  float last_t = t;
  uint32_t step = 0;
  while (t < r.far && step < num_steps) {
    if (m.probe(r, t, x, y, z, dt)) {
      px[0] = x, px[1] = y, px[2] = z;
      pd[0] = r.dx, pd[1] = r.dy, pd[2] = r.dz;
      pl[0] = dt;
      pl[1] = t - last_t;
      last_t = t;
      px += 3, pd += 3, pl += 2;
      ++step;
    }
  }
}

// adds (total samples, number of rays) to counter[0..1] like the reference's atomicAdds do
__global__ void march_finish_kernel(const int32_t* __restrict__ offsets, uint32_t n_rays, int32_t* __restrict__ counter) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    counter[0] += offsets[n_rays];
    counter[1] += static_cast<int32_t>(n_rays);
  }
}

__device__ __forceinline__ int warp_iscan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

// single-CTA exclusive scan (offsets[n] = total); n_rays <= 2^18 in every configuration
__global__ void __launch_bounds__(1024)
march_scan_kernel(const int32_t* __restrict__ counts, uint32_t n, int32_t* __restrict__ offsets) {
  __shared__ int warp_total[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const int v = i < n ? counts[i] : 0;
    const int incl = warp_iscan(v, lane);
    if (lane == 31) warp_total[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int tot = warp_total[lane];
      warp_total[lane] = warp_iscan(tot, lane) - tot;
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

// Two-level form for many rays (the single CTA above walks the counts 1024 at a time: 256 dependent rounds at 2^18
// rays): every CTA scans its own 1024 counts, the block totals go through march_scan_kernel, a third pass adds the
// block offsets.
__global__ void __launch_bounds__(1024)
march_scan_block_kernel(const int32_t* __restrict__ counts, uint32_t n, int32_t* __restrict__ offsets,
                        int32_t* __restrict__ block_sums) {
  __shared__ int warp_total[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
  const int v = i < n ? counts[i] : 0;
  const int incl = warp_iscan(v, lane);
  if (lane == 31) warp_total[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int tot = warp_total[lane];
    const int scan = warp_iscan(tot, lane);
    warp_total[lane] = scan - tot;
    if (lane == 31) block_sums[blockIdx.x] = scan;
  }
  __syncthreads();
  if (i < n) offsets[i] = warp_total[wid] + incl - v;
}

__global__ void __launch_bounds__(1024)
march_scan_add_kernel(int32_t* __restrict__ offsets, uint32_t n, const int32_t* __restrict__ block_offs) {
  const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
  if (i < n) offsets[i] += block_offs[blockIdx.x];
  if (i == 0) offsets[n] = block_offs[gridDim.x];
}

constexpr uint32_t kScanOneCtaRays = 4096;  // up to here the single CTA is the shortest path (6 us at 4096 rays)

// ------------------------------------------------------------------ inference wavefront (raymarching.cu:528-634)
__global__ void march_rays_kernel(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                                  const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                  const float* __restrict__ rays_d, March m, const float* __restrict__ nears,
                                  const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                                  float* __restrict__ deltas, uint32_t perturb) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_alive) return;
  const int index = rays_alive[n];
  const Ray r = load_ray(rays_o, rays_d, index, fars[index]);
  float t = rays_t[n];
  if (perturb) {
    Pcg32 rng;
    rng.seed(static_cast<uint64_t>(n), static_cast<uint64_t>(perturb));
    t += min_stepsize() * rng.next_float();
  }
  float* px = xyzs + 3ull * n * n_step;
  float* pd = dirs + 3ull * n * n_step;
  float* pl = deltas + 2ull * n * n_step;
  float last_t = t, x, y, z, dt;
  uint32_t step = 0;
  while (t < r.far && step < n_step) {
    if (m.probe(r, t, x, y, z, dt)) {
      px[0] = x, px[1] = y, px[2] = z;
      pd[0] = r.dx, pd[1] = r.dy, pd[2] = r.dz;
      pl[0] = dt;
      pl[1] = t - last_t;
      last_t = t;
      px += 3, pd += 3, pl += 2;
      ++step;
    }
  }
}

// ------------------------------------------------------------------ ragged compositing, training
// (raymarching.cu:318-487 for rgb + depth; the semantic channels are the kernels the reference declares but never
//  implemented, raymarching.h:12-13: semantics_n = sum_s w_s * p_s with the weights detached on that branch like
//  the live path, renderer_semantics.py:270)
// One warp per ray, lanes = 32 consecutive samples of a chunk: every lane loads its own sample, the transmittance in
// front of it comes from a multiplicative warp scan times the product carried over the earlier chunks (the scheme of
// weights.cuh on the live path), running sums from additive scans -- the loads of a chunk are independent, so the
// kernels run at memory speed instead of one dependent load round per sample (the sequential form, all 32 lanes
// repeating the chain of the reference's single thread, took 0.24 + 0.20 ms for 1.6 M samples).  The class channels
// keep lanes = classes: the weight of sample k is broadcast by shuffle and the [C] row is read / written coalesced.
// Same terms as raymarching.cu:318-487, summed in scan order: equal to the sequential result within fp32 rounding.
constexpr int kLogitsLd = UCSA_MAX_CLASSES;  // row stride of the semantic head's logits (heads kernels: [K,48] fp16)
constexpr int kTileStride = kLogitsLd + 1;
static_assert(kLogitsLd % 8 == 0, "logit rows are read as 16-byte vectors");

struct RaggedChunk {
  float alpha, w, t_after;  // of this lane's sample (alpha = w = 0 past the end of the ray)
};

__device__ __forceinline__ RaggedChunk ragged_chunk(float alpha, float& t_carry, int lane) {
  const float incl = warp_scan_mul(1.0f - alpha, lane);
  float excl = __shfl_up_sync(kFullMask, incl, 1);
  if (lane == 0) excl = 1.0f;
  RaggedChunk o;
  o.alpha = alpha;
  o.w = alpha * (t_carry * excl);
  o.t_after = t_carry * incl;
  t_carry *= __shfl_sync(kFullMask, incl, 31);
  return o;
}

// inclusive running sum over the ray: carry + scan inside the chunk; advances carry
__device__ __forceinline__ float ragged_running(float v, float& carry, int lane) {
  const float incl = warp_scan_add(v, lane);
  const float out = carry + incl;
  carry += __shfl_sync(kFullMask, incl, 31);
  return out;
}

__global__ void composite_train_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                           const float* __restrict__ sem, const __half* __restrict__ logits,
                                           uint32_t logits_ld, const float* __restrict__ deltas,
                                           const int32_t* __restrict__ rays, uint32_t M, uint32_t N, uint32_t C,
                                           float* __restrict__ weights_sum, float* __restrict__ depth,
                                           float* __restrict__ image, float* __restrict__ semantics) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
  const bool empty = num_steps == 0 || offset + num_steps >= M;
  float r = 0, g = 0, b = 0, ws = 0, d = 0;  // per-lane partial sums
  float acc0 = 0.f, acc1 = 0.f;              // classes lane, lane + 32
  float t_carry = 1.0f, time_carry = 0.f;
  // the warp's tile for the lanes = samples soft-max: rows of kLogitsLd halves, 16-byte aligned, at most kLogitsLd classes
  __shared__ float tiles[4 * 32 * kTileStride];
  float* tile = (logits != nullptr && logits_ld == kLogitsLd && C <= kLogitsLd &&
                 (reinterpret_cast<uintptr_t>(logits) & 15u) == 0)
                    ? tiles + (threadIdx.x >> 5) * 32 * kTileStride
                    : nullptr;
  if (!empty) {
    for (uint32_t base = 0; base < num_steps; base += 32) {
      const uint32_t s = base + lane;
      const bool valid = s < num_steps;
      const uint32_t i = offset + (valid ? s : 0u);
      float alpha = 0.f, dt_sum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (valid) {
        const float2 dl = *reinterpret_cast<const float2*>(deltas + 2ull * i);
        alpha = 1.0f - __expf(-sigmas[i] * dl.x);
        dt_sum = dl.y;
        c0 = rgbs[3ull * i], c1 = rgbs[3ull * i + 1], c2 = rgbs[3ull * i + 2];
      }
      const RaggedChunk q = ragged_chunk(alpha, t_carry, lane);
      const float t = ragged_running(dt_sum, time_carry, lane);
      r += q.w * c0;
      g += q.w * c1;
      b += q.w * c2;
      d += q.w * t;
      ws += q.w;
      if (sem != nullptr) {
        const uint32_t cnt = num_steps - base < 32u ? num_steps - base : 32u;
        const float* row = sem + static_cast<uint64_t>(offset + base) * C;
#pragma unroll 4
        for (uint32_t k = 0; k < cnt; ++k) {
          const float wk = __shfl_sync(kFullMask, q.w, static_cast<int>(k));
          if (static_cast<uint32_t>(lane) < C) acc0 += wk * row[static_cast<uint64_t>(k) * C + lane];
          if (static_cast<uint32_t>(lane) + 32 < C) acc1 += wk * row[static_cast<uint64_t>(k) * C + lane + 32];
        }
      } else if (logits != nullptr && tile != nullptr) {
        // class probabilities from the fp16 logits of the semantic head; the [M,C] fp32 probability tensor is never
        // formed.  Soft-max with lanes = SAMPLES (a lane reads its own 96-byte row, one max / exp / sum chain per
        // lane, no shuffles), exp values parked in the warp's shared-memory tile [32][49] (odd stride: conflict-free
        // both ways), then lanes = classes add up the chunk's rows.  The soft-max across the lanes (below, kept for
        // other row strides) costs two warp reductions per SAMPLE: ncu showed the kernel issue-bound with it (0.30 ms
        // for 1.6 M samples at 8 % of the DRAM throughput).
        const uint32_t cnt = num_steps - base < 32u ? num_steps - base : 32u;
        float scale = 0.f;
        if (valid) {
          const uint4* lg = reinterpret_cast<const uint4*>(logits + static_cast<uint64_t>(i) * kLogitsLd);
          uint4 v[kLogitsLd / 8];
#pragma unroll
          for (int j = 0; j < kLogitsLd / 8; ++j) v[j] = __ldg(lg + j);
          const __half* hv = reinterpret_cast<const __half*>(v);
          float m = -INFINITY;
#pragma unroll
          for (int c = 0; c < kLogitsLd; ++c)
            if (static_cast<uint32_t>(c) < C) m = fmaxf(m, __half2float(hv[c]));
          float sum = 0.f;
          float* mine = tile + lane * kTileStride;
#pragma unroll
          for (int c = 0; c < kLogitsLd; ++c) {
            if (static_cast<uint32_t>(c) < C) {
              const float e = __expf(__half2float(hv[c]) - m);
              sum += e;
              mine[c] = e;
            }
          }
          scale = q.w / sum;
        }
        __syncwarp();
        for (uint32_t k = 0; k < cnt; ++k) {
          const float sk = __shfl_sync(kFullMask, scale, static_cast<int>(k));
          if (static_cast<uint32_t>(lane) < C) acc0 += sk * tile[k * kTileStride + lane];
          if (static_cast<uint32_t>(lane) + 32 < C) acc1 += sk * tile[k * kTileStride + lane + 32];
        }
        __syncwarp();
      } else if (logits != nullptr) {
        // (any row stride) soft-max across the lanes, as in composite_rays_kernel
        const uint32_t cnt = num_steps - base < 32u ? num_steps - base : 32u;
        const __half* row = logits + static_cast<uint64_t>(offset + base) * logits_ld;
#pragma unroll 4
        for (uint32_t k = 0; k < cnt; ++k) {
          const float wk = __shfl_sync(kFullMask, q.w, static_cast<int>(k));
          const __half* lg = row + static_cast<uint64_t>(k) * logits_ld;
          const float l0 = static_cast<uint32_t>(lane) < C ? __half2float(lg[lane]) : -INFINITY;
          const float l1 = static_cast<uint32_t>(lane) + 32 < C ? __half2float(lg[lane + 32]) : -INFINITY;
          const float m = warp_max(fmaxf(l0, l1));
          const float e0 = static_cast<uint32_t>(lane) < C ? __expf(l0 - m) : 0.f;
          const float e1 = static_cast<uint32_t>(lane) + 32 < C ? __expf(l1 - m) : 0.f;
          const float scale = wk / warp_sum(e0 + e1);
          acc0 += scale * e0;
          acc1 += scale * e1;
        }
      }
    }
  }
  r = warp_sum(r), g = warp_sum(g), b = warp_sum(b), ws = warp_sum(ws), d = warp_sum(d);
  if (lane == 0) {
    weights_sum[index] = ws;
    depth[index] = d;
    image[index * 3] = r, image[index * 3 + 1] = g, image[index * 3 + 2] = b;
  }
  if (semantics != nullptr) {
    if (static_cast<uint32_t>(lane) < C) semantics[static_cast<uint64_t>(index) * C + lane] = acc0;
    if (static_cast<uint32_t>(lane) + 32 < C) semantics[static_cast<uint64_t>(index) * C + lane + 32] = acc1;
  }
}

__global__ void composite_train_bwd_kernel(const float* __restrict__ grad_ws, const float* __restrict__ grad_image,
                                           const float* __restrict__ grad_sem, const float* __restrict__ sigmas,
                                           const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                           const int32_t* __restrict__ rays, const float* __restrict__ weights_sum,
                                           const float* __restrict__ image, uint32_t M, uint32_t N, uint32_t C,
                                           float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs,
                                           float* __restrict__ grad_local_sem) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
  if (num_steps == 0 || offset + num_steps >= M) return;
  const float gi0 = grad_image[3 * index], gi1 = grad_image[3 * index + 1], gi2 = grad_image[3 * index + 2];
  const float gws = grad_ws[index];
  const float rf = image[3 * index], gf = image[3 * index + 1], bf = image[3 * index + 2], wsf = weights_sum[index];
  float gs0 = 0.f, gs1 = 0.f;
  if (grad_sem != nullptr) {
    if (static_cast<uint32_t>(lane) < C) gs0 = grad_sem[static_cast<uint64_t>(index) * C + lane];
    if (static_cast<uint32_t>(lane) + 32 < C) gs1 = grad_sem[static_cast<uint64_t>(index) * C + lane + 32];
  }
  float t_carry = 1.0f, r_carry = 0.f, g_carry = 0.f, b_carry = 0.f, ws_carry = 0.f;
  for (uint32_t base = 0; base < num_steps; base += 32) {
    const uint32_t s = base + lane;
    const bool valid = s < num_steps;
    const uint32_t i = offset + (valid ? s : 0u);
    float alpha = 0.f, dt = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (valid) {
      dt = deltas[2ull * i];
      alpha = 1.0f - __expf(-sigmas[i] * dt);
      c0 = rgbs[3ull * i], c1 = rgbs[3ull * i + 1], c2 = rgbs[3ull * i + 2];
    }
    const RaggedChunk q = ragged_chunk(alpha, t_carry, lane);
    // colour / weight accumulated up to and including this sample, transmittance behind it (raymarching.cu:470-480)
    const float r = ragged_running(q.w * c0, r_carry, lane);
    const float g = ragged_running(q.w * c1, g_carry, lane);
    const float b = ragged_running(q.w * c2, b_carry, lane);
    const float ws = ragged_running(q.w, ws_carry, lane);
    if (valid) {
      grad_rgbs[3ull * i] = gi0 * q.w;
      grad_rgbs[3ull * i + 1] = gi1 * q.w;
      grad_rgbs[3ull * i + 2] = gi2 * q.w;
      const float T = q.t_after;
      grad_sigmas[i] = dt * (gi0 * (T * c0 - (rf - r)) + gi1 * (T * c1 - (gf - g)) + gi2 * (T * c2 - (bf - b)) +
                             gws * (T - (wsf - ws)));
    }
    if (grad_local_sem != nullptr) {
      const uint32_t cnt = num_steps - base < 32u ? num_steps - base : 32u;
      float* row = grad_local_sem + static_cast<uint64_t>(offset + base) * C;
#pragma unroll 4
      for (uint32_t k = 0; k < cnt; ++k) {
        const float wk = __shfl_sync(kFullMask, q.w, static_cast<int>(k));
        if (static_cast<uint32_t>(lane) < C) row[static_cast<uint64_t>(k) * C + lane] = wk * gs0;
        if (static_cast<uint32_t>(lane) + 32 < C) row[static_cast<uint64_t>(k) * C + lane + 32] = wk * gs1;
      }
    }
  }
}

// ------------------------------------------------------------------ inference compositing (raymarching.cu:647-729)
__global__ void composite_rays_kernel(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                                      float* __restrict__ rays_t, const float* __restrict__ sigmas,
                                      const float* __restrict__ rgbs, const float* __restrict__ sem,
                                      const __half* __restrict__ logits, uint32_t logits_ld,
                                      const float* __restrict__ deltas, uint32_t C, float* __restrict__ weights_sum,
                                      float* __restrict__ depth, float* __restrict__ image,
                                      float* __restrict__ semantics) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= n_alive) return;
  const int index = rays_alive[n];
  float t = rays_t[n];
  float ws = weights_sum[index], d = depth[index];
  float r = image[3 * index], g = image[3 * index + 1], b = image[3 * index + 2];
  float acc0 = 0.f, acc1 = 0.f;
  const bool with_sem = sem != nullptr || logits != nullptr;
  if (with_sem) {
    if (static_cast<uint32_t>(lane) < C) acc0 = semantics[static_cast<uint64_t>(index) * C + lane];
    if (static_cast<uint32_t>(lane) + 32 < C) acc1 = semantics[static_cast<uint64_t>(index) * C + lane + 32];
  }
  uint32_t step = 0;
  while (step < n_step) {
    const uint32_t i = n * n_step + step;
    if (deltas[2 * i] == 0) break;
    const float alpha = 1.0f - __expf(-sigmas[i] * deltas[2 * i]);
    const float T = 1 - ws;
    const float w = alpha * T;
    ws += w;
    t += deltas[2 * i + 1];
    d += w * t;
    r += w * rgbs[3 * i];
    g += w * rgbs[3 * i + 1];
    b += w * rgbs[3 * i + 2];
    if (sem != nullptr) {
      if (static_cast<uint32_t>(lane) < C) acc0 += w * sem[static_cast<uint64_t>(i) * C + lane];
      if (static_cast<uint32_t>(lane) + 32 < C) acc1 += w * sem[static_cast<uint64_t>(i) * C + lane + 32];
    } else if (logits != nullptr) {
      // class probabilities from the fp16 logits of the semantic head, soft-max across the warp's lanes (the loop is
      // warp-uniform: every lane follows the same ray): saves the [M,C] probability tensor and two passes over it
      const __half* lg = logits + static_cast<uint64_t>(i) * logits_ld;
      const float l0 = static_cast<uint32_t>(lane) < C ? __half2float(lg[lane]) : -INFINITY;
      const float l1 = static_cast<uint32_t>(lane) + 32 < C ? __half2float(lg[lane + 32]) : -INFINITY;
      const float m = warp_max(fmaxf(l0, l1));
      const float e0 = static_cast<uint32_t>(lane) < C ? __expf(l0 - m) : 0.f;
      const float e1 = static_cast<uint32_t>(lane) + 32 < C ? __expf(l1 - m) : 0.f;
      const float scale = w / warp_sum(e0 + e1);
      acc0 += scale * e0;
      acc1 += scale * e1;
    }
    if (T < 1e-4) break;
    ++step;
  }
  __syncwarp();
  if (lane == 0) {
    rays_t[n] = step < n_step ? -1.0f : t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[3 * index] = r, image[3 * index + 1] = g, image[3 * index + 2] = b;
  }
  if (with_sem) {
    if (static_cast<uint32_t>(lane) < C) semantics[static_cast<uint64_t>(index) * C + lane] = acc0;
    if (static_cast<uint32_t>(lane) + 32 < C) semantics[static_cast<uint64_t>(index) * C + lane + 32] = acc1;
  }
}

// ------------------------------------------------------------------ live-ray compaction, order preserving
// one CTA; warp ballots + prefix inside 1024-ray chunks, running base across chunks
__global__ void __launch_bounds__(1024)
compact_rays_kernel(uint32_t n_alive, int32_t* __restrict__ rays_alive, const int32_t* __restrict__ rays_alive_old,
                    float* __restrict__ rays_t, const float* __restrict__ rays_t_old, int32_t* __restrict__ alive_counter) {
  __shared__ int warp_total[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = alive_counter[0];
  __syncthreads();
  for (uint32_t base = 0; base < n_alive; base += 1024) {
    const uint32_t n = base + threadIdx.x;
    const float t_old = n < n_alive ? rays_t_old[n] : -1.0f;
    const bool keep = n < n_alive && t_old >= 0;
    const unsigned ballot = __ballot_sync(kFullMask, keep);
    if (lane == 0) warp_total[wid] = __popc(ballot);
    __syncthreads();
    if (wid == 0) {
      const int tot = warp_total[lane];
      warp_total[lane] = warp_iscan(tot, lane) - tot;
    }
    __syncthreads();
    const int carry = carry_s;
    if (keep) {
      const int dst = carry + warp_total[wid] + __popc(ballot & ((1u << lane) - 1u));
      rays_alive[dst] = rays_alive_old[n];
      rays_t[dst] = t_old;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_total[wid] + __popc(ballot);
    __syncthreads();
  }
  if (threadIdx.x == 0) alive_counter[0] = carry_s;
}

// The same compaction for many rays, three launches instead of one CTA walking the whole list (a 640x480 wavefront is
// 300 serial 1024-ray chunks: 0.35 ms per round, the largest item of the inference path): per-CTA counts -> scan of
// the <= 1024 counts -> order-preserving writes at the scanned offsets.  `scratch` holds one int per 1024-ray chunk.
__global__ void __launch_bounds__(1024)
compact_count_kernel(uint32_t n_alive, const float* __restrict__ rays_t_old, int32_t* __restrict__ scratch) {
  __shared__ int warp_total[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * 1024u + threadIdx.x;
  const bool keep = n < n_alive && rays_t_old[n] >= 0;
  const unsigned ballot = __ballot_sync(kFullMask, keep);
  if (lane == 0) warp_total[wid] = __popc(ballot);
  __syncthreads();
  if (wid == 0) {
    int tot = warp_total[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
    if (lane == 0) scratch[blockIdx.x] = tot;
  }
}
__global__ void __launch_bounds__(1024)
compact_scan_kernel(uint32_t n_blocks, int32_t* __restrict__ scratch, int32_t* __restrict__ alive_counter) {
  __shared__ int warp_total[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = alive_counter[0];
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const int v = i < n_blocks ? scratch[i] : 0;
    const int incl = warp_iscan(v, lane);
    if (lane == 31) warp_total[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int tot = warp_total[lane];
      warp_total[lane] = warp_iscan(tot, lane) - tot;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < n_blocks) scratch[i] = carry + warp_total[wid] + incl - v;  // first output slot of chunk i
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_total[wid] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) alive_counter[0] = carry_s;
}
__global__ void __launch_bounds__(1024)
compact_write_kernel(uint32_t n_alive, int32_t* __restrict__ rays_alive, const int32_t* __restrict__ rays_alive_old,
                     float* __restrict__ rays_t, const float* __restrict__ rays_t_old,
                     const int32_t* __restrict__ scratch) {
  __shared__ int warp_total[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * 1024u + threadIdx.x;
  const float t_old = n < n_alive ? rays_t_old[n] : -1.0f;
  const bool keep = n < n_alive && t_old >= 0;
  const unsigned ballot = __ballot_sync(kFullMask, keep);
  if (lane == 0) warp_total[wid] = __popc(ballot);
  __syncthreads();
  if (wid == 0) {
    const int tot = warp_total[lane];
    warp_total[lane] = warp_iscan(tot, lane) - tot;
  }
  __syncthreads();
  if (keep) {
    const int dst = scratch[blockIdx.x] + warp_total[wid] + __popc(ballot & ((1u << lane) - 1u));
    rays_alive[dst] = rays_alive_old[n];
    rays_t[dst] = t_old;
  }
}

// ------------------------------------------------------------------ occupancy grid maintenance (row a20)
// density_grid = max(density_grid * decay, fresh) where fresh >= 0 (torch-ngp's update rule; not in the reference)
__global__ void grid_update_kernel(float* __restrict__ grid, const float* __restrict__ fresh, uint64_t n, float decay) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float f = fresh[i];
  if (f >= 0.f) grid[i] = fmaxf(grid[i] * decay, f);
}
// The same update with the bookkeeping that follows it fused in: fresh is scaled by density_scale, and the sum of
// max(grid, 0) over all cells (for mean_density) is accumulated in double (per-CTA partial sums, one atomic each).
__global__ void __launch_bounds__(256)
grid_update_sum_kernel(float* __restrict__ grid, const float* __restrict__ fresh, uint64_t n, float decay,
                       float fresh_scale, double* __restrict__ sum) {
  __shared__ double part[8];
  double acc = 0.0;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    float gval = grid[i];
    const float f = fresh[i] * fresh_scale;
    if (f >= 0.f) {
      gval = fmaxf(gval * decay, f);
      grid[i] = gval;
    }
    acc += static_cast<double>(fmaxf(gval, 0.f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += part[w];
    atomicAdd(sum, tot);
  }
}
// bitfield with the threshold min(0.01, mean) taken from the device-side sum
__global__ void grid_packbits_dev_kernel(const float* __restrict__ grid, uint64_t n, const double* __restrict__ sum,
                                         uint32_t* __restrict__ bits, float* __restrict__ mean_out) {
  const float mean = static_cast<float>(*sum / static_cast<double>(n));
  const float thresh = fminf(kDensityThresh, mean);
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0 && mean_out != nullptr) *mean_out = mean;
  const bool occ = i < n && grid[i] > thresh;
  const unsigned word = __ballot_sync(kFullMask, occ);
  if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = word;
}
// bit i = grid[i] > thresh, 32 cells per thread-word via ballot
__global__ void grid_packbits_kernel(const float* __restrict__ grid, uint64_t n, float thresh, uint32_t* __restrict__ bits) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool occ = i < n && grid[i] > thresh;
  const unsigned word = __ballot_sync(kFullMask, occ);
  if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = word;
}

int make_march(March& m, const float* grid, const uint32_t* bits, float mean_density, float bound, float dt_gamma,
               uint32_t C, uint32_t H) {
  UCSA_REQUIRE(grid != nullptr || bits != nullptr, "march: need a density grid or a bitfield");
  UCSA_REQUIRE(C >= 1 && H >= 2 && bound > 0.f, "march: bad grid geometry");
  m.bound = bound;
  m.dt_gamma = dt_gamma;
  m.dt_min = min_stepsize();
  m.dt_max = 2 * bound / H;
  m.C = C;
  m.H = H;
  m.occ.grid = grid;
  m.occ.bits = bits;
  m.occ.thresh = fminf(kDensityThresh, mean_density);
  return UCSA_OK;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_march_rays_train(const float* rays_o, const float* rays_d, const float* grid,
                                     const uint32_t* bitfield, float mean_density, float bound, float dt_gamma,
                                     uint32_t n_rays, uint32_t C, uint32_t H, uint32_t max_points, const float* nears,
                                     const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays,
                                     int32_t* counter, uint32_t perturb, int32_t* scratch, float* t_stage,
                                     void* stream) {
  UCSA_REQUIRE(rays_o && rays_d && nears && fars && xyzs && dirs && deltas && rays && counter && scratch,
               "march_rays_train: null pointer");
  March m;
  if (int rc = make_march(m, grid, bitfield, mean_density, bound, dt_gamma, C, H)) return rc;
  if (n_rays == 0) return UCSA_OK;
  cudaStream_t st = as_stream(stream);
  int32_t* counts = scratch;            // [n_rays]
  int32_t* offsets = scratch + n_rays;  // [n_rays + 1]
  march_count_kernel<<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, m, n_rays, nears, fars, perturb, counts,
                                                            t_stage);
  if (n_rays <= kScanOneCtaRays) {
    march_scan_kernel<<<1, 1024, 0, st>>>(counts, n_rays, offsets);
  } else {
    const uint32_t n_blocks = ceil_div(n_rays, 1024u);
    int32_t* block_sums = offsets + n_rays + 1;     // [n_blocks]
    int32_t* block_offs = block_sums + n_blocks;    // [n_blocks + 1]
    march_scan_block_kernel<<<n_blocks, 1024, 0, st>>>(counts, n_rays, offsets, block_sums);
    march_scan_kernel<<<1, 1024, 0, st>>>(block_sums, n_blocks, block_offs);
    march_scan_add_kernel<<<n_blocks, 1024, 0, st>>>(offsets, n_rays, block_offs);
  }
  // the reference accumulates into `counter`; offsets start at its current value (0 for a fresh counter)
  if (t_stage != nullptr)
    march_write_staged_kernel<<<ceil_div(static_cast<uint64_t>(n_rays) * 32, 128), 128, 0, st>>>(
        rays_o, rays_d, m, n_rays, max_points, nears, perturb, counts, offsets, t_stage, xyzs, dirs, deltas, rays);
  else
    march_write_kernel<<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, m, n_rays, max_points, nears, fars,
                                                              perturb, counts, offsets, 0, 0, xyzs, dirs, deltas, rays);
  march_finish_kernel<<<1, 32, 0, st>>>(offsets, n_rays, counter);
  return check_launch("march_rays_train");
}

extern "C" int ucsa_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                               const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t C,
                               uint32_t H, const float* grid, const uint32_t* bitfield, float mean_density,
                               const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                               uint32_t perturb, void* stream) {
  UCSA_REQUIRE(rays_alive && rays_t && rays_o && rays_d && nears && fars && xyzs && dirs && deltas,
               "march_rays: null pointer");
  March m;
  if (int rc = make_march(m, grid, bitfield, mean_density, bound, dt_gamma, C, H)) return rc;
  if (n_alive == 0) return UCSA_OK;
  march_rays_kernel<<<ceil_div(n_alive, 128), 128, 0, as_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o,
                                                                           rays_d, m, nears, fars, xyzs, dirs, deltas,
                                                                           perturb);
  return check_launch("march_rays");
}

extern "C" int ucsa_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* local_semantics,
                                                 const void* logits_h, uint32_t logits_ld, const float* deltas,
                                                 const int32_t* rays, uint32_t M, uint32_t N, uint32_t n_classes,
                                                 float* weights_sum, float* depth, float* image, float* semantics,
                                                 void* stream) {
  UCSA_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image, "composite_rays_train_forward: null pointer");
  UCSA_REQUIRE(!(local_semantics && logits_h), "composite_rays_train_forward: give probabilities or logits, not both");
  UCSA_REQUIRE((local_semantics == nullptr && logits_h == nullptr) == (semantics == nullptr),
               "composite_rays_train_forward: semantics in/out mismatch");
  UCSA_REQUIRE(logits_h == nullptr || logits_ld >= n_classes, "composite_rays_train_forward: logits row stride < classes");
  UCSA_REQUIRE(n_classes <= 64, "composite_rays_train_forward: at most 64 classes");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7u) == 0, "composite_rays_train_forward: deltas [M,2] must be 8-byte aligned");
  if (N == 0) return UCSA_OK;
  composite_train_fwd_kernel<<<ceil_div(static_cast<uint64_t>(N) * 32, 128), 128, 0, as_stream(stream)>>>(
      sigmas, rgbs, local_semantics, static_cast<const __half*>(logits_h), logits_ld, deltas, rays, M, N, n_classes,
      weights_sum, depth, image, semantics);
  return check_launch("composite_rays_train_forward");
}

extern "C" int ucsa_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                                  const float* grad_semantics, const float* sigmas, const float* rgbs,
                                                  const float* deltas, const int32_t* rays, const float* weights_sum,
                                                  const float* image, uint32_t M, uint32_t N, uint32_t n_classes,
                                                  float* grad_sigmas, float* grad_rgbs, float* grad_local_semantics,
                                                  void* stream) {
  UCSA_REQUIRE(grad_weights_sum && grad_image && sigmas && rgbs && deltas && rays && weights_sum && image &&
                   grad_sigmas && grad_rgbs,
               "composite_rays_train_backward: null pointer");
  UCSA_REQUIRE(n_classes <= 64, "composite_rays_train_backward: at most 64 classes");
  if (N == 0) return UCSA_OK;
  composite_train_bwd_kernel<<<ceil_div(static_cast<uint64_t>(N) * 32, 128), 128, 0, as_stream(stream)>>>(
      grad_weights_sum, grad_image, grad_semantics, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, n_classes,
      grad_sigmas, grad_rgbs, grad_local_semantics);
  return check_launch("composite_rays_train_backward");
}

extern "C" int ucsa_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                                   const float* sigmas, const float* rgbs, const float* local_semantics,
                                   const void* logits_h, uint32_t logits_ld, const float* deltas, uint32_t n_classes,
                                   float* weights_sum, float* depth, float* image, float* semantics, void* stream) {
  UCSA_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, "composite_rays: null pointer");
  UCSA_REQUIRE(!(local_semantics && logits_h), "composite_rays: give probabilities or logits, not both");
  UCSA_REQUIRE((local_semantics == nullptr && logits_h == nullptr) == (semantics == nullptr),
               "composite_rays: semantics in/out mismatch");
  UCSA_REQUIRE(logits_h == nullptr || logits_ld >= n_classes, "composite_rays: logits row stride < classes");
  UCSA_REQUIRE(n_classes <= 64, "composite_rays: at most 64 classes");
  if (n_alive == 0) return UCSA_OK;
  composite_rays_kernel<<<ceil_div(static_cast<uint64_t>(n_alive) * 32, 128), 128, 0, as_stream(stream)>>>(
      n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, local_semantics, static_cast<const __half*>(logits_h), logits_ld,
      deltas, n_classes, weights_sum, depth, image, semantics);
  return check_launch("composite_rays");
}

extern "C" int ucsa_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                                 const float* rays_t_old, int32_t* alive_counter, int32_t* scratch, void* stream) {
  UCSA_REQUIRE(rays_alive && rays_alive_old && rays_t && rays_t_old && alive_counter, "compact_rays: null pointer");
  if (scratch == nullptr || n_alive <= 4096) {  // a few chunks: one CTA is the shortest path
    compact_rays_kernel<<<1, 1024, 0, as_stream(stream)>>>(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old,
                                                           alive_counter);
    return check_launch("compact_rays");
  }
  const uint32_t n_blocks = ceil_div(n_alive, 1024);
  compact_count_kernel<<<n_blocks, 1024, 0, as_stream(stream)>>>(n_alive, rays_t_old, scratch);
  compact_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(n_blocks, scratch, alive_counter);
  compact_write_kernel<<<n_blocks, 1024, 0, as_stream(stream)>>>(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old,
                                                                 scratch);
  return check_launch("compact_rays");
}

extern "C" int ucsa_grid_update(float* density_grid, const float* fresh, uint64_t n_cells, float decay, void* stream) {
  UCSA_REQUIRE(density_grid && fresh, "grid_update: null pointer");
  if (n_cells == 0) return UCSA_OK;
  grid_update_kernel<<<ceil_div(n_cells, 256), 256, 0, as_stream(stream)>>>(density_grid, fresh, n_cells, decay);
  return check_launch("grid_update");
}

extern "C" int ucsa_grid_update_pack(float* density_grid, const float* fresh, uint64_t n_cells, float decay,
                                     float fresh_scale, double* sum_scratch, float* mean_density_dev,
                                     uint32_t* bitfield, void* stream) {
  UCSA_REQUIRE(density_grid && fresh && sum_scratch && mean_density_dev && bitfield, "grid_update_pack: null pointer");
  UCSA_REQUIRE(n_cells % 32 == 0 && n_cells > 0, "grid_update_pack: cell count must be a positive multiple of 32");
  if (cudaMemsetAsync(sum_scratch, 0, sizeof(double), as_stream(stream)) != cudaSuccess) return check_launch("grid_update_pack");
  grid_update_sum_kernel<<<kNumSMs * 8, 256, 0, as_stream(stream)>>>(density_grid, fresh, n_cells, decay, fresh_scale,
                                                                     sum_scratch);
  grid_packbits_dev_kernel<<<ceil_div(n_cells, 256), 256, 0, as_stream(stream)>>>(density_grid, n_cells, sum_scratch,
                                                                                  bitfield, mean_density_dev);
  return check_launch("grid_update_pack");
}

extern "C" int ucsa_grid_packbits(const float* density_grid, uint64_t n_cells, float mean_density, uint32_t* bitfield,
                                  void* stream) {
  UCSA_REQUIRE(density_grid && bitfield, "grid_packbits: null pointer");
  UCSA_REQUIRE(n_cells % 32 == 0, "grid_packbits: cell count must be a multiple of 32");
  if (n_cells == 0) return UCSA_OK;
  grid_packbits_kernel<<<ceil_div(n_cells, 256), 256, 0, as_stream(stream)>>>(density_grid, n_cells,
                                                                              fminf(kDensityThresh, mean_density), bitfield);
  return check_launch("grid_packbits");
}
