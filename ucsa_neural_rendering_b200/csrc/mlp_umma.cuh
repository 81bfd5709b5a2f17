// tcgen05 / TMEM building blocks for the 64-wide MLPs (row a6 of SURVEY.md section 8) on sm_100a.
//
// One CTA = 128 threads = one 128-row tile; thread t owns row t of every activation tile and TMEM lane t of
// every accumulator (UMMA M = 128, cta_group::1).  All shared-memory operands use ONE layout, the un-swizzled
// "core matrix" form of the UMMA descriptors: an [R x C] half tile is stored as
//
//        [R/8 row groups][C/8 column chunks][8 rows][8 halves]          (a core matrix = 8 x 16 B = 128 B)
//
// Because a core matrix is 8 rows x 8 columns in either reading, the very same bytes are a valid
//   * K-major  operand (rows = M or N, columns = K):  LBO = 128 B (next K chunk), SBO = C*16 B (next row group)
//   * MN-major operand (rows = K, columns = M or N):  SBO = 128 B (next MN chunk), LBO = C*16 B (next K group)
// so an activation tile written once by the epilogue serves as A of the next layer (K-major), as B of the
// weight-gradient product (MN-major), and a weight tile serves forward (K-major B) and data-gradient (MN-major B).
//
// Weight gradients accumulate in TMEM over all tiles a persistent CTA processes (UMMA M = 64: row m of the
// accumulator lives in TMEM lane (m % 16) + 32 * (m / 16)).
#pragma once

#include "common.cuh"

namespace ucsa {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done) {  // a lost arrival becomes a launch failure after ~2 s instead of a hang
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  } while (!done);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier over a subset of the CTA's warps (id 1..15; `threads` = a multiple of 32): the worker warps of a
// warp-specialised kernel meet here without the driver warp
__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- bulk (TMA) tile copies, issued by ONE thread
// The activations saved for the backward pass live in HBM in the core-matrix tile layout described above, one
// contiguous Tile<C>::kBytes block per 128-row tile, so a tile moves as a single bulk copy instead of 128 threads
// each storing / loading their own 2C-byte row (which costs one L1 wavefront per row and 16-byte chunk).
// `policy` = l2_policy_stream(): written once, read once.
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(ssrc), "r"(bytes), "l"(policy)
               : "memory");
}
// closes the group of bulk stores issued so far and waits until they have READ their shared-memory source
__device__ __forceinline__ void bulk_store_fence_reads() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// global -> shared; completion is signalled on `bar` (arm it with mbar_expect_tx first)
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(sdst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
template <int C>
__device__ __forceinline__ unsigned char* tile_block(void* base, uint64_t tile) {
  return static_cast<unsigned char*>(base) + tile * (128 * C * 2);
}
template <int C>
__device__ __forceinline__ const unsigned char* tile_block(const void* base, uint64_t tile) {
  return static_cast<const unsigned char*>(base) + tile * (128 * C * 2);
}

// ---------------------------------------------------------------- fences
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM allocation (one full warp calls these)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// ---------------------------------------------------------------- descriptors
// shared-memory matrix descriptor, no swizzle, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// instruction descriptor for kind::f16: fp16 A/B, fp32 accumulate
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers (32 lanes x 32 bit, N columns)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- the core-matrix tile
template <int C>
struct Tile {
  static constexpr int kCols = C;
  static constexpr uint32_t kGroupBytes = C * 16;  // one 8-row group
  static constexpr uint32_t kBytes = 128 / 8 * kGroupBytes;
  __device__ static __forceinline__ uint32_t offset(int row, int col) {  // bytes
    return static_cast<uint32_t>((row >> 3) * kGroupBytes + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2);
  }
  // 8 halves (one chunk) of a row
  __device__ static __forceinline__ uint4* chunk(unsigned char* base, int row, int chunk_idx) {
    return reinterpret_cast<uint4*>(base + (row >> 3) * kGroupBytes + chunk_idx * 128 + (row & 7) * 16);
  }
  // operand descriptors; `k16` selects the 16-wide K step
  __device__ static __forceinline__ uint64_t k_major(uint32_t base, int k16) {  // rows = M/N, cols = K
    return smem_desc(base + k16 * 256, 128, kGroupBytes);
  }
  __device__ static __forceinline__ uint64_t mn_major(uint32_t base, int k16) {  // rows = K, cols = M/N
    return smem_desc(base + k16 * 2 * kGroupBytes, kGroupBytes, 128);
  }
};

// fp16 row-major weights [OUT][IN] in global memory -> core-matrix tile (rows = OUT, cols = IN); whole CTA
template <int IN>
__device__ __forceinline__ void load_weight_tile(unsigned char* dst, const __half* __restrict__ w, int out_rows) {
  const int chunks = IN / 8;
  for (int i = threadIdx.x; i < out_rows * chunks; i += blockDim.x) {
    const int o = i / chunks, c = i % chunks;
    *Tile<IN>::chunk(dst, o, c) = __ldg(reinterpret_cast<const uint4*>(w + o * IN + c * 8));
  }
}

// D[128 x N] = A[128 x K] . W^T   (forward; A tile K-major, W tile [N][K] K-major)
template <int K, int N>
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile) {
  constexpr uint32_t idesc = instr_desc(128, N, false, false);
#pragma unroll
  for (int k = 0; k < K / 16; ++k)
    mma_f16(tmem_d, Tile<K>::k_major(a_tile, k), Tile<K>::k_major(w_tile, k), idesc, k > 0);
}
// D[128 x IN] = dY[128 x OUT] . W   (data gradient; dY tile K-major, W tile [OUT][IN] read MN-major)
template <int OUT, int IN>
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t dy_tile, uint32_t w_tile) {
  constexpr uint32_t idesc = instr_desc(128, IN, false, true);
#pragma unroll
  for (int k = 0; k < OUT / 16; ++k)
    mma_f16(tmem_d, Tile<OUT>::k_major(dy_tile, k), Tile<IN>::mn_major(w_tile, k), idesc, k > 0);
}
// D[64 x N] (+)= P^T[64 x 128] . Q[128 x N]   (weight gradient; both tiles read MN-major, K = the 128 rows)
template <int N>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t p_tile, uint32_t q_tile, bool first) {
  constexpr uint32_t idesc = instr_desc(64, N, true, true);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    mma_f16(tmem_d, Tile<64>::mn_major(p_tile, k), Tile<N>::mn_major(q_tile, k), idesc, !(first && k == 0));
}

// Per-CTA tensor-core context: TMEM base + one mbarrier with its phase.
struct Ctx {
  uint32_t tmem;
  uint64_t* bar;
  uint32_t phase;

  // all threads: make tile writes visible, then thread 0 may issue
  __device__ __forceinline__ void publish() {
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
  }
  __device__ __forceinline__ void wait() {
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
  }
  // TMEM address of column `col` for the calling warp's 32 lanes
  __device__ __forceinline__ uint32_t lane_addr(uint32_t col) const {
    return tmem + (((threadIdx.x >> 5) * 32u) << 16) + col;
  }
};

// set-up / tear-down (whole CTA).  `cols` must be a power of two >= 32.
__device__ __forceinline__ Ctx ctx_init(uint32_t* tmem_slot, uint64_t* bar, uint32_t cols) {
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, cols);
  if (threadIdx.x == 32) mbar_init(bar, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return Ctx{*tmem_slot, bar, 0u};
}
__device__ __forceinline__ void ctx_free(const Ctx& c, uint32_t cols) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(c.tmem, cols);
}

// epilogue helper: accumulator columns [col0, col0+16) of this thread's row -> 16 halves (two chunks) of a tile,
// with optional ReLU (forward) or ReLU mask taken from an activation tile (backward)
template <int C, bool RELU>
__device__ __forceinline__ void acc_to_tile16(const Ctx& c, uint32_t col0, unsigned char* tile, int tile_col0,
                                              const unsigned char* mask_tile = nullptr) {
  float v[16];
  tmem_ld16(c.lane_addr(col0), v);
  const int row = threadIdx.x;
#pragma unroll
  for (int half_idx = 0; half_idx < 2; ++half_idx) {
    H8 o, m;
    if (mask_tile != nullptr)
      m.v = *reinterpret_cast<const uint4*>(mask_tile + Tile<C>::offset(row, tile_col0 + half_idx * 8));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = v[half_idx * 8 + i];
      if (RELU) x = fmaxf(x, 0.f);
      if (mask_tile != nullptr && !(__half2float(m.h[i]) > 0.f)) x = 0.f;
      o.h[i] = __float2half_rn(x);
    }
    *Tile<C>::chunk(tile, row, (tile_col0 >> 3) + half_idx) = o.v;
  }
}

}  // namespace umma
}  // namespace ucsa
