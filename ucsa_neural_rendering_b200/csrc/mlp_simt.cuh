// Thread-per-row MLP building blocks on CUDA cores (fp16 operands, fp32 accumulate).
//
// Arithmetic contract (rows a6 of SURVEY.md section 8; frozen in oracle/tcnn_spec.py): weights and
// activations are fp16 values, products are accumulated in fp32, ReLU is applied to the fp32 sum and the
// result is rounded to fp16.  These blocks are the plain-CUDA realisation of that contract: they are what
// the tcgen05 tile path (mlp_umma.cuh) is checked against on the GPU, and they serve the row counts too
// ragged for a 128-row tensor-core tile.
//
// Layout: a CTA owns a tile of 128 rows, one thread per row.  Activations of a row live in that thread's
// row of a shared-memory tile [128][W+8] of halves (the +8 keeps 16-byte row accesses bank-conflict free);
// weights live in shared memory as fp32 [out][in] (holding fp16-representable values).
#pragma once

#include "common.cuh"

namespace ucsa {

constexpr int kTileRows = 128;
__host__ __device__ constexpr int tile_ld(int width) { return width + 8; }

// fp16 global weights [n] -> fp32 shared copy, whole CTA cooperates.
__device__ __forceinline__ void load_weights_f32(float* __restrict__ dst, const __half* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __half2float(src[i]);
}

// y[OUT] = act(W[OUT][IN] . x[IN]); x and y are this thread's rows of two tiles.
template <int IN, int OUT, bool RELU>
__device__ __forceinline__ void dense_row_fwd(const float* __restrict__ w, const __half* __restrict__ xrow,
                                              __half* __restrict__ yrow) {
  float x[IN];
#pragma unroll
  for (int k = 0; k < IN; k += 8) {
    H8 v;
    v.v = *reinterpret_cast<const uint4*>(xrow + k);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[k + i] = __half2float(v.h[i]);
  }
#pragma unroll 1
  for (int j0 = 0; j0 < OUT; j0 += 8) {
    float acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.f;
#pragma unroll
    for (int k = 0; k < IN; k += 4) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 wv = *reinterpret_cast<const float4*>(w + (j0 + jj) * IN + k);  // warp-wide broadcast
        acc[jj] = fmaf(x[k + 0], wv.x, acc[jj]);
        acc[jj] = fmaf(x[k + 1], wv.y, acc[jj]);
        acc[jj] = fmaf(x[k + 2], wv.z, acc[jj]);
        acc[jj] = fmaf(x[k + 3], wv.w, acc[jj]);
      }
    }
    H8 o;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) o.h[jj] = __float2half_rn(RELU ? fmaxf(acc[jj], 0.f) : acc[jj]);
    *reinterpret_cast<uint4*>(yrow + j0) = o.v;
  }
}

// dx[IN] = W^T . dy[OUT]  (dy is this thread's row of a tile).  Result stays in fp32 registers.
template <int IN, int OUT>
__device__ __forceinline__ void dense_row_bwd(const float* __restrict__ w, const __half* __restrict__ dyrow,
                                              float (&dx)[IN]) {
#pragma unroll
  for (int k = 0; k < IN; ++k) dx[k] = 0.f;
#pragma unroll 1
  for (int j0 = 0; j0 < OUT; j0 += 8) {
    H8 v;
    v.v = *reinterpret_cast<const uint4*>(dyrow + j0);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float g = __half2float(v.h[jj]);
#pragma unroll
      for (int k = 0; k < IN; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + (j0 + jj) * IN + k);
        dx[k + 0] = fmaf(g, wv.x, dx[k + 0]);
        dx[k + 1] = fmaf(g, wv.y, dx[k + 1]);
        dx[k + 2] = fmaf(g, wv.z, dx[k + 2]);
        dx[k + 3] = fmaf(g, wv.w, dx[k + 3]);
      }
    }
  }
}

// Store dx (fp32 regs) into this thread's row as fp16, zeroing where the forward activation was <= 0.
template <int N, bool RELU_MASK>
__device__ __forceinline__ void store_row_masked(const float (&dx)[N], const __half* __restrict__ actrow,
                                                 __half* __restrict__ outrow) {
#pragma unroll
  for (int k = 0; k < N; k += 8) {
    H8 a, o;
    if (RELU_MASK) a.v = *reinterpret_cast<const uint4*>(actrow + k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = dx[k + i];
      if (RELU_MASK && !(__half2float(a.h[i]) > 0.f)) v = 0.f;
      o.h[i] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(outrow + k) = o.v;
  }
}

// Weight-gradient contribution of one 128-row tile: acc[j][k] += sum_r dy[r][j] * x[r][k].
// 128 threads share the OUT*IN entries: thread t owns column k = t % IN and the J = OUT*IN/128 outputs
// j = (t / IN) * J + jj.  Requires IN in {16,32,64} and J a multiple of 8.
template <int IN, int OUT>
struct WGrad {
  static constexpr int kGroups = kTileRows / IN;
  static constexpr int J = OUT / kGroups;
  static_assert(J % 8 == 0, "weight-gradient slice must be a multiple of 8 outputs");
  float acc[J];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < J; ++i) acc[i] = 0.f;
  }
  __device__ __forceinline__ void add_tile(const __half* __restrict__ dy_tile, const __half* __restrict__ x_tile) {
    const int k = threadIdx.x % IN;
    const int j0 = (threadIdx.x / IN) * J;
#pragma unroll 2
    for (int r = 0; r < kTileRows; ++r) {
      const float xv = __half2float(x_tile[r * tile_ld(IN) + k]);
#pragma unroll
      for (int jj = 0; jj < J; jj += 8) {
        H8 v;
        v.v = *reinterpret_cast<const uint4*>(dy_tile + r * tile_ld(OUT) + j0 + jj);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[jj + i] = fmaf(__half2float(v.h[i]), xv, acc[jj + i]);
      }
    }
  }
  // grad_w is the fp32 [OUT][IN] block of this layer
  __device__ __forceinline__ void flush(float* __restrict__ grad_w, float scale) const {
    const int k = threadIdx.x % IN;
    const int j0 = (threadIdx.x / IN) * J;
#pragma unroll
    for (int jj = 0; jj < J; ++jj) atomicAdd(grad_w + (j0 + jj) * IN + k, acc[jj] * scale);
  }
};

}  // namespace ucsa
