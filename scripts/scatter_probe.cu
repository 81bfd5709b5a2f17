// Bring-up probe (not part of the library): how fast can the hash-grid gradient scatter go on a B200?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/scatter_probe scripts/scatter_probe.cu
// Variants of one thread = one (sample, level) scatter over ray-ordered samples (4096 rays x 512 samples).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ucsa_neural_rendering_b200/csrc/grid.cuh"

using namespace ucsa;

__device__ __forceinline__ void red_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_h2(__half2* addr, __half2 v) {
  asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(*reinterpret_cast<uint32_t*>(&v)) : "memory");
}

// mode 0: v2.f32 per corner; 1: two scalar f32; 2: f16x2; 3: v4 merge of aligned x pairs; 4: warp-aggregated (match)
template <int MODE>
__global__ void scatter_kernel(const float* __restrict__ x01, uint32_t n, ucsa_grid_desc grid,
                               const __half* __restrict__ d_enc, float* __restrict__ gt, __half2* __restrict__ gth,
                               int l_lo, int l_hi, int level_major) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_lv = l_hi - l_lo;
  if (i >= (uint64_t)n * n_lv) return;
  uint32_t s, l;
  if (level_major) {  // consecutive threads = consecutive samples of one level
    l = l_lo + (uint32_t)(i / n);
    s = (uint32_t)(i % n);
  } else {
    s = (uint32_t)(i / n_lv);
    l = l_lo + (uint32_t)(i % n_lv);
  }
  const float x[3] = {x01[3ull * s], x01[3ull * s + 1], x01[3ull * s + 2]};
  const float2 g = __half22float2(reinterpret_cast<const __half2*>(d_enc)[(uint64_t)s * 16 + l]);
  const LevelGeom lv = level_geom(grid, l);
  const Cell cell = locate(lv, x);
  if (MODE == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = corner_weight(cell, c);
      red_add_f32x2(gt + 2ull * corner_entry(lv, cell, c), w * g.x, w * g.y);
    }
  } else if (MODE == 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = corner_weight(cell, c);
      float* p = gt + 2ull * corner_entry(lv, cell, c);
      atomicAdd(p, w * g.x);
      atomicAdd(p + 1, w * g.y);
    }
  } else if (MODE == 2) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = corner_weight(cell, c);
      red_h2(gth + corner_entry(lv, cell, c), __floats2half2_rn(w * g.x, w * g.y));
    }
  } else if (MODE == 3) {
#pragma unroll
    for (int c = 0; c < 8; c += 2) {  // corners c (x) and c+1 (x+1)
      const uint32_t e0 = corner_entry(lv, cell, c), e1 = corner_entry(lv, cell, c + 1);
      const float w0 = corner_weight(cell, c), w1 = corner_weight(cell, c + 1);
      if ((e0 ^ e1) == 1u) {
        float* p = gt + 2ull * (e0 & ~1u);
        if (e0 & 1u) red_v4(p, w1 * g.x, w1 * g.y, w0 * g.x, w0 * g.y);
        else red_v4(p, w0 * g.x, w0 * g.y, w1 * g.x, w1 * g.y);
      } else {
        red_add_f32x2(gt + 2ull * e0, w0 * g.x, w0 * g.y);
        red_add_f32x2(gt + 2ull * e1, w1 * g.x, w1 * g.y);
      }
    }
  } else if (MODE == 5 || MODE == 6) {
    // replicated coarse levels: CTA b adds into replica b % R of the dense levels, cutting same-address contention
    const int R = MODE == 5 ? 16 : 64;
    float* base = gt;
    if (!lv.hashed) base = gt + 2ull * grid.total_entries + (size_t)(blockIdx.x % R) * 2ull * grid.offset[4];
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const uint32_t e0 = corner_entry(lv, cell, c), e1 = corner_entry(lv, cell, c + 1);
      const float w0 = corner_weight(cell, c), w1 = corner_weight(cell, c + 1);
      if ((e0 ^ e1) == 1u) {
        float* p = base + 2ull * (e0 & ~1u);
        if (e0 & 1u) red_v4(p, w1 * g.x, w1 * g.y, w0 * g.x, w0 * g.y);
        else red_v4(p, w0 * g.x, w0 * g.y, w1 * g.x, w1 * g.y);
      } else {
        red_add_f32x2(base + 2ull * e0, w0 * g.x, w0 * g.y);
        red_add_f32x2(base + 2ull * e1, w1 * g.x, w1 * g.y);
      }
    }
  } else if (MODE == 4) {
    // lanes of a warp that sit in the same cell of the same level: the lowest lane of each group adds for all
    const uint32_t key = (cell.c[0] * 73856093u) ^ (cell.c[1] * 19349663u) ^ (cell.c[2] * 83492791u) ^ (l << 28);
    const unsigned peers = __match_any_sync(__activemask(), key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = corner_weight(cell, c);
      float a = w * g.x, b = w * g.y;
      // sum over the peer group with a loop over set bits (groups are small or the whole warp)
      float sa = 0.f, sb = 0.f;
      unsigned m = peers;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        sa += __shfl_sync(peers, a, src);
        sb += __shfl_sync(peers, b, src);
      }
      if (lane == leader) red_add_f32x2(gt + 2ull * corner_entry(lv, cell, c), sa, sb);
    }
  }
}

int main(int argc, char** argv) {
  const uint32_t n_rays = 4096, t = 512, n = n_rays * t;
  std::vector<float> x(3ull * n);
  srand(1);
  auto rnd = [] { return rand() / (float)RAND_MAX; };
  const bool same_origin = argc > 1;
  const float o_shared[3] = {1.1f, 0.05f, 0.03f};
  for (uint32_t r = 0; r < n_rays; ++r) {
    float o[3] = {(rnd() - .5f) * 2, (rnd() - .5f) * 2, (rnd() - .5f) * .5f}, d[3], nn = 0;
    if (same_origin) { o[0] = o_shared[0]; o[1] = o_shared[1]; o[2] = o_shared[2]; }
    for (int k = 0; k < 3; ++k) { d[k] = rnd() - .5f; nn += d[k] * d[k]; }
    if (same_origin) { d[0] = -0.6f + 0.5f * (rnd() - .5f); d[1] = 0.7f + 0.5f * (rnd() - .5f); d[2] = -0.3f + 0.4f * (rnd() - .5f); nn = d[0]*d[0]+d[1]*d[1]+d[2]*d[2]; }
    for (int k = 0; k < 3; ++k) d[k] /= sqrtf(nn);
    float far = 1e9f;
    for (int k = 0; k < 3; ++k) { float tt = ((d[k] > 0 ? 4.f : -4.f) - o[k]) / d[k]; if (tt < far) far = tt; }
    for (uint32_t j = 0; j < t; ++j) {
      // half the samples uniform, half clustered around a "surface" at 60 % of the ray
      float z = j < 256 ? 0.2f + (far - 0.2f) * (j + rnd()) / 256.f : 0.6f * far + (rnd() - .5f) * 0.3f;
      for (int k = 0; k < 3; ++k) {
        float p = fminf(fmaxf(o[k] + d[k] * z, -4.f), 4.f);
        x[3ull * (r * t + j) + k] = (p + 4.f) / 8.f;
      }
    }
  }
  ucsa_grid_desc grid;
  ucsa_grid_desc_init(1.5157166f, 16, 19, &grid);
  float *dx, *gt;
  __half* denc;
  __half2* gth;
  cudaMalloc(&dx, x.size() * 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&denc, (size_t)n * 32 * 2);
  cudaMemset(denc, 0x3c, (size_t)n * 32 * 2);
  cudaMalloc(&gt, (size_t)grid.total_entries * 8 + 64ull * grid.offset[4] * 8);
  cudaMalloc(&gth, (size_t)grid.total_entries * 4);
  float* flush;
  cudaMalloc(&flush, 256u << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](const char* name, int mode, int lo, int hi, int lm) {
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
      cudaMemset(flush, 0, 256u << 20);
      cudaMemset(gt, 0, (size_t)grid.total_entries * 8);
      cudaMemset(gth, 0, (size_t)grid.total_entries * 4);
      const uint64_t total = (uint64_t)n * (hi - lo);
      const uint32_t blocks = (uint32_t)((total + 255) / 256);
      cudaEventRecord(e0);
      switch (mode) {
        case 0: scatter_kernel<0><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 1: scatter_kernel<1><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 2: scatter_kernel<2><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 3: scatter_kernel<3><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 4: scatter_kernel<4><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 5: scatter_kernel<5><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
        case 6: scatter_kernel<6><<<blocks, 256>>>(dx, n, grid, denc, gt, gth, lo, hi, lm); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double atoms = (double)n * (hi - lo) * 8;
    printf("%-44s levels [%2d,%2d) %s  %7.3f ms  %6.1f G corner-adds/s  (%s)\n", name, lo, hi,
           lm ? "level-major " : "sample-major", best, atoms / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int lm = 0; lm < 2; ++lm) {
    run("v2.f32 per corner", 0, 0, 16, lm);
    run("2 x scalar f32", 1, 0, 16, lm);
    run("f16x2", 2, 0, 16, lm);
    run("v4 merge of aligned x pairs", 3, 0, 16, lm);
    run("warp-aggregated (match_any)", 4, 0, 16, lm);
    run("v4 + 16 replicas of dense levels", 5, 0, 16, lm);
    run("v4 + 64 replicas of dense levels", 6, 0, 16, lm);
  }
  for (int l = 0; l < 5; l += 1) {
    run("v2.f32", 0, l, l + 1, 0);
    run("warp-aggregated", 4, l, l + 1, 0);
    run("16 replicas", 5, l, l + 1, 0);
    run("64 replicas", 6, l, l + 1, 0);
  }
  return 0;
}
