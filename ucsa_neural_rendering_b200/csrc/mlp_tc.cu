// Stand-alone fused MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM): forward and
// backward of the three networks of the reference (sigma 32-64-16, colour 32-64-64-16, semantics 16-64-48;
// network_tcnn_semantics.py:48-58,74-84,90-100 = tcnn FullyFusedMLP).  Row a6 of SURVEY.md section 8.
//
// Persistent CTAs of 128 threads; one 128-row tile per iteration; weights resident in shared memory for the
// CTA's lifetime; activations go TMEM -> registers (ReLU, fp16) -> shared-memory tile of the next layer without
// touching HBM; weight gradients accumulate in TMEM across all tiles of the CTA and are flushed once.
#include "mlp_umma.cuh"

namespace ucsa {
namespace {

using umma::Tile;

template <int W>
__device__ __forceinline__ void row_g2t(unsigned char* tile, const __half* __restrict__ src, bool valid) {
#pragma unroll
  for (int c = 0; c < W / 8; ++c)
    *Tile<W>::chunk(tile, threadIdx.x, c) =
        valid ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0, 0, 0, 0);
}
template <int W>
__device__ __forceinline__ void row_t2g(__half* __restrict__ dst, const unsigned char* tile) {
#pragma unroll
  for (int c = 0; c < W / 8; ++c)
    reinterpret_cast<uint4*>(dst)[c] = *Tile<W>::chunk(const_cast<unsigned char*>(tile), threadIdx.x, c);
}
// accumulator columns [0,W) of this thread's row -> fp16 row in global memory.  tcgen05.ld is warp-collective
// (.sync.aligned): every thread of the warp executes the load, only the store is predicated on `valid`.
template <int W>
__device__ __forceinline__ void acc_to_global(const umma::Ctx& ctx, uint32_t col0, __half* __restrict__ dst,
                                              bool valid) {
#pragma unroll
  for (int c0 = 0; c0 < W; c0 += 16) {
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(col0 + c0), v);
    H8 a, b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a.h[i] = __float2half_rn(v[i]);
      b.h[i] = __float2half_rn(v[8 + i]);
    }
    if (valid) {
      reinterpret_cast<uint4*>(dst + c0)[0] = a.v;
      reinterpret_cast<uint4*>(dst + c0)[1] = b.v;
    }
  }
}

template <int D0, int D1, int D2, int D3>
struct Shape {
  static constexpr bool kThree = D3 > 0;
  static constexpr int kOut = kThree ? D3 : D2;
  static constexpr int kActs = kThree ? D1 + D2 : D1;
  static constexpr uint32_t kW0 = D1 / 8 * Tile<D0>::kGroupBytes;
  static constexpr uint32_t kW1 = D2 / 8 * Tile<D1>::kGroupBytes;
  static constexpr uint32_t kW2 = kThree ? D3 / 8 * Tile<D2>::kGroupBytes : 0;
  static constexpr uint32_t kWeights = kW0 + kW1 + kW2;
  static constexpr uint32_t kFwdTiles = Tile<D0>::kBytes + Tile<D1>::kBytes + (kThree ? Tile<D2>::kBytes : 0);
  static constexpr uint32_t kFwdSmem = kWeights + kFwdTiles + 64;
  static constexpr uint32_t kFwdCols = kThree ? 256 : 128;
  // backward: x, a1, [a2], dy, d1, [d2]
  static constexpr uint32_t kBwdTiles = Tile<D0>::kBytes + 2 * Tile<D1>::kBytes + Tile<kOut>::kBytes +
                                        (kThree ? 2 * Tile<D2>::kBytes : 0);
  static constexpr uint32_t kBwdSmem = kWeights + kBwdTiles + 64;
  static constexpr uint32_t kBwdCols = kThree ? 256 : 128;
};

template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(128)
mlp_fwd_tc_kernel(const __half* __restrict__ x, uint32_t n, const __half* __restrict__ w, __half* __restrict__ y,
                  __half* __restrict__ acts) {
  using S = Shape<D0, D1, D2, D3>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* w0 = smem;
  unsigned char* w1 = w0 + S::kW0;
  unsigned char* w2 = w1 + S::kW1;
  unsigned char* t0 = w2 + S::kW2;
  unsigned char* t1 = t0 + Tile<D0>::kBytes;
  unsigned char* t2 = t1 + Tile<D1>::kBytes;
  unsigned char* tail = t2 + (S::kThree ? Tile<D2>::kBytes : 0);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 8);

  umma::load_weight_tile<D0>(w0, w, D1);
  umma::load_weight_tile<D1>(w1, w + D0 * D1, D2);
  if (S::kThree) umma::load_weight_tile<D2>(w2, w + D0 * D1 + D1 * D2, D3);
  umma::Ctx ctx = umma::ctx_init(slot, bar, S::kFwdCols);
  const uint32_t a0 = umma::smem_u32(t0), a1 = umma::smem_u32(t1), a2 = umma::smem_u32(t2);
  const uint32_t b0 = umma::smem_u32(w0), b1 = umma::smem_u32(w1), b2 = umma::smem_u32(w2);
  constexpr uint32_t kAcc0 = 0, kAcc1 = D1, kAcc2 = D1 + D2;

  const uint32_t n_tiles = (n + 127) / 128;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < n;
    row_g2t<D0>(t0, x + static_cast<uint64_t>(r) * D0, valid);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<D0, D1>(ctx.tmem + kAcc0, a0, b0);
      umma::commit(ctx.bar);
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < D1; c0 += 16) umma::acc_to_tile16<D1, true>(ctx, kAcc0 + c0, t1, c0);
    if (acts != nullptr && valid) row_t2g<D1>(acts + static_cast<uint64_t>(r) * S::kActs, t1);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<D1, D2>(ctx.tmem + kAcc1, a1, b1);
      umma::commit(ctx.bar);
    }
    ctx.wait();
    if (S::kThree) {
#pragma unroll
      for (int c0 = 0; c0 < D2; c0 += 16) umma::acc_to_tile16<D2, true>(ctx, kAcc1 + c0, t2, c0);
      if (acts != nullptr && valid) row_t2g<D2>(acts + static_cast<uint64_t>(r) * S::kActs + D1, t2);
      ctx.publish();
      if (threadIdx.x == 0) {
        umma::tc_fence_after();
        umma::issue_fwd<D2, (D3 > 0 ? D3 : 16)>(ctx.tmem + kAcc2, a2, b2);
        umma::commit(ctx.bar);
      }
      ctx.wait();
      acc_to_global<(D3 > 0 ? D3 : 16)>(ctx, kAcc2, y + static_cast<uint64_t>(r) * S::kOut, valid);
    } else {
      acc_to_global<D2>(ctx, kAcc1, y + static_cast<uint64_t>(r) * S::kOut, valid);
    }
  }
  umma::ctx_free(ctx, S::kFwdCols);
}

// flush a [64 x N] weight-gradient accumulator (UMMA M = 64: row m in TMEM lane (m%16) + 32*(m/16)).
// natural: acc[m][n] is d(W[m][n]) with leading dimension `ld`; transposed: acc[m][n] is d(W[n][m]).
template <int N, bool TRANSPOSED>
__device__ __forceinline__ void flush_wgrad(const umma::Ctx& ctx, uint32_t col0, float* __restrict__ grad, int ld,
                                            float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(col0 + c0), v);  // warp-collective: every lane takes part
    if (lane < 16) {
      const int m = warp * 16 + lane;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int nn = c0 + i;
        atomicAdd(grad + (TRANSPOSED ? nn * ld + m : m * ld + nn), v[i] * scale);
      }
    }
  }
}

template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(128)
mlp_bwd_tc_kernel(const __half* __restrict__ x, uint32_t n, const __half* __restrict__ w,
                  const __half* __restrict__ acts, const __half* __restrict__ dy, float inv_scale,
                  __half* __restrict__ dx, float* __restrict__ grad_w) {
  using S = Shape<D0, D1, D2, D3>;
  constexpr int DL = S::kOut;
  static_assert(D1 == 64 && (!S::kThree || D2 == 64), "hidden layers are 64 wide");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* w0 = smem;
  unsigned char* w1 = w0 + S::kW0;
  unsigned char* w2 = w1 + S::kW1;
  unsigned char* t_x = w2 + S::kW2;
  unsigned char* t_a1 = t_x + Tile<D0>::kBytes;
  unsigned char* t_d1 = t_a1 + Tile<D1>::kBytes;
  unsigned char* t_dy = t_d1 + Tile<D1>::kBytes;
  unsigned char* t_a2 = t_dy + Tile<DL>::kBytes;
  unsigned char* t_d2 = t_a2 + (S::kThree ? Tile<D2>::kBytes : 0);
  unsigned char* tail = t_d2 + (S::kThree ? Tile<D2>::kBytes : 0);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 8);

  umma::load_weight_tile<D0>(w0, w, D1);
  umma::load_weight_tile<D1>(w1, w + D0 * D1, D2);
  if (S::kThree) umma::load_weight_tile<D2>(w2, w + D0 * D1 + D1 * D2, D3);
  umma::Ctx ctx = umma::ctx_init(slot, bar, S::kBwdCols);
  const uint32_t s_x = umma::smem_u32(t_x), s_a1 = umma::smem_u32(t_a1), s_d1 = umma::smem_u32(t_d1),
                 s_dy = umma::smem_u32(t_dy), s_a2 = umma::smem_u32(t_a2), s_d2 = umma::smem_u32(t_d2);
  const uint32_t b0 = umma::smem_u32(w0), b1 = umma::smem_u32(w1), b2 = umma::smem_u32(w2);
  // TMEM columns: [0,64) data-gradient scratch, then the weight-gradient accumulators
  constexpr uint32_t kAcc = 0, kG0 = 64, kG1 = kG0 + D0, kG2 = kG1 + (S::kThree ? D1 : DL);

  const uint32_t n_tiles = (n + 127) / 128;
  bool first = true;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < n;
    if (!first) ctx.wait();  // weight-gradient MMAs of the previous tile have consumed the tiles
    row_g2t<D0>(t_x, x + static_cast<uint64_t>(r) * D0, valid);
    row_g2t<D1>(t_a1, acts + static_cast<uint64_t>(r) * S::kActs, valid);
    row_g2t<DL>(t_dy, dy + static_cast<uint64_t>(r) * DL, valid);
    if (S::kThree) row_g2t<D2>(t_a2, acts + static_cast<uint64_t>(r) * S::kActs + D1, valid);
    ctx.publish();
    if (S::kThree) {
      if (threadIdx.x == 0) {
        umma::tc_fence_after();
        umma::issue_dgrad<DL, D2>(ctx.tmem + kAcc, s_dy, b2);
        umma::commit(ctx.bar);
        umma::issue_wgrad<DL>(ctx.tmem + kG2, s_a2, s_dy, first);  // d(W2)^T = a2^T . dy
      }
      ctx.wait();
#pragma unroll
      for (int c0 = 0; c0 < D2; c0 += 16) umma::acc_to_tile16<D2, false>(ctx, kAcc + c0, t_d2, c0, t_a2);
      ctx.publish();
      if (threadIdx.x == 0) {
        umma::tc_fence_after();
        umma::issue_dgrad<D2, D1>(ctx.tmem + kAcc, s_d2, b1);
        umma::commit(ctx.bar);
        umma::issue_wgrad<D1>(ctx.tmem + kG1, s_d2, s_a1, first);  // d(W1) = d2^T . a1
      }
    } else {
      if (threadIdx.x == 0) {
        umma::tc_fence_after();
        umma::issue_dgrad<DL, D1>(ctx.tmem + kAcc, s_dy, b1);
        umma::commit(ctx.bar);
        umma::issue_wgrad<DL>(ctx.tmem + kG1, s_a1, s_dy, first);  // d(W1)^T = a1^T . dy
      }
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < D1; c0 += 16) umma::acc_to_tile16<D1, false>(ctx, kAcc + c0, t_d1, c0, t_a1);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      if (dx != nullptr) {
        umma::issue_dgrad<D1, D0>(ctx.tmem + kAcc, s_d1, b0);
        umma::commit(ctx.bar);
      }
      umma::issue_wgrad<D0>(ctx.tmem + kG0, s_d1, s_x, first);  // d(W0) = d1^T . x
      if (dx == nullptr) umma::commit(ctx.bar);
    }
    if (dx != nullptr) {
      ctx.wait();
      acc_to_global<D0>(ctx, kAcc, dx + static_cast<uint64_t>(r) * D0, valid);
      umma::tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        umma::tc_fence_after();
        umma::commit(ctx.bar);  // covers the weight-gradient MMAs issued above
      }
    }
    first = false;
  }
  ctx.wait();  // all weight-gradient MMAs done
  flush_wgrad<D0, false>(ctx, kG0, grad_w, D0, inv_scale);
  if (S::kThree) {
    flush_wgrad<D1, false>(ctx, kG1, grad_w + D0 * D1, D1, inv_scale);
    flush_wgrad<DL, true>(ctx, kG2, grad_w + D0 * D1 + D1 * D2, D2, inv_scale);
  } else {
    flush_wgrad<DL, true>(ctx, kG1, grad_w + D0 * D1, D1, inv_scale);
  }
  umma::ctx_free(ctx, S::kBwdCols);
}

uint32_t tc_grid(uint32_t n, int per_sm) {
  const uint32_t tiles = (n + 127) / 128;
  const uint32_t cap = kNumSMs * per_sm;
  return tiles < cap ? tiles : cap;
}

template <int D0, int D1, int D2, int D3>
int launch_fwd(const void* x, uint32_t n, const void* w, void* y, void* acts, cudaStream_t st) {
  using S = Shape<D0, D1, D2, D3>;
  if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(mlp_fwd_tc_kernel<D0, D1, D2, D3>), S::kFwdSmem, "mlp_fwd_tc_kernel")) return rc;
  mlp_fwd_tc_kernel<D0, D1, D2, D3><<<tc_grid(n, 512 / S::kFwdCols), 128, S::kFwdSmem, st>>>(
      static_cast<const __half*>(x), n, static_cast<const __half*>(w), static_cast<__half*>(y),
      static_cast<__half*>(acts));
  return check_launch("mlp_fwd");
}
template <int D0, int D1, int D2, int D3>
int launch_bwd(const void* x, uint32_t n, const void* w, const void* acts, const void* dy, float inv_scale, void* dx,
               float* grad_w, cudaStream_t st) {
  using S = Shape<D0, D1, D2, D3>;
  if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(mlp_bwd_tc_kernel<D0, D1, D2, D3>), S::kBwdSmem, "mlp_bwd_tc_kernel")) return rc;
  mlp_bwd_tc_kernel<D0, D1, D2, D3><<<tc_grid(n, 512 / S::kBwdCols), 128, S::kBwdSmem, st>>>(
      static_cast<const __half*>(x), n, static_cast<const __half*>(w), static_cast<const __half*>(acts),
      static_cast<const __half*>(dy), inv_scale, static_cast<__half*>(dx), grad_w);
  return check_launch("mlp_bwd");
}

int shape_id(const uint32_t* d, uint32_t n_layers) {
  if (n_layers == 2 && d[0] == 32 && d[1] == 64 && d[2] == 16) return 0;                 // sigma net
  if (n_layers == 3 && d[0] == 32 && d[1] == 64 && d[2] == 64 && d[3] == 16) return 1;   // colour net
  if (n_layers == 2 && d[0] == 16 && d[1] == 64 && d[2] == 48) return 2;                 // semantic net
  return -1;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_mlp_fwd(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims, uint32_t n_layers,
                            void* y_h, void* acts_h, void* stream) {
  UCSA_REQUIRE(x_h && w_h && dims && y_h, "mlp_fwd: null pointer");
  if (n == 0) return UCSA_OK;
  switch (shape_id(dims, n_layers)) {
    case 0: return launch_fwd<32, 64, 16, 0>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    case 1: return launch_fwd<32, 64, 64, 16>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    case 2: return launch_fwd<16, 64, 48, 0>(x_h, n, w_h, y_h, acts_h, as_stream(stream));
    default:
      set_error("mlp_fwd: unsupported layer widths (supported: 32-64-16, 32-64-64-16, 16-64-48)");
      return UCSA_ERR_UNSUPPORTED;
  }
}

extern "C" int ucsa_mlp_bwd(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims, uint32_t n_layers,
                            const void* acts_h, const void* dy_h, float inv_loss_scale, void* dx_h, float* grad_w,
                            void* stream) {
  UCSA_REQUIRE(x_h && w_h && dims && acts_h && dy_h && grad_w, "mlp_bwd: null pointer");
  if (n == 0) return UCSA_OK;
  switch (shape_id(dims, n_layers)) {
    case 0: return launch_bwd<32, 64, 16, 0>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    case 1: return launch_bwd<32, 64, 64, 16>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    case 2: return launch_bwd<16, 64, 48, 0>(x_h, n, w_h, acts_h, dy_h, inv_loss_scale, dx_h, grad_w, as_stream(stream));
    default:
      set_error("mlp_bwd: unsupported layer widths (supported: 32-64-16, 32-64-64-16, 16-64-48)");
      return UCSA_ERR_UNSUPPORTED;
  }
}
