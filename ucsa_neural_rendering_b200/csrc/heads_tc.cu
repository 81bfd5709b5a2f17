// Colour and semantic heads on the masked-in samples, on the tcgen05 tensor cores: SH degree-4 direction
// encoding, colour MLP (SH16 + geo15 + 1 -> 64 -> 64 -> 3, sigmoid) and semantic MLP (geo15 + 1 -> 64 -> C).
// Rows a7, a12, a13 and their backward of SURVEY.md section 8 (network_tcnn_semantics.py:147-207).
//
// The reference evaluates the heads only where w > 1e-4 (boolean-mask gather / scatter, :159-161,:174); here the
// K surviving rows arrive as the compact list `sel` (row -> n*T+slot) built by compact_masked.  Colour and semantic
// networks run in separate kernels, forward and backward, so that each CTA carries only its own weights, tiles and
// TMEM columns and four to five CTAs are resident per SM; the weight gradients accumulate in TMEM across all tiles
// of a persistent CTA.
//
// The hidden activations saved for the backward pass (hc1, hc2, hs) are tile-layout buffers (mlp_umma.cuh): tile i
// of the compact row list is one contiguous 16 KB block, written and read back with single bulk copies.  The
// logits are not saved at all: the semantic backward kernel recomputes them from hs with one extra product.
#include <cstdlib>
#include <mutex>

#include "mlp_umma.cuh"
#include "sh4.cuh"

namespace ucsa {
namespace {

using umma::Tile;

__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kSemOut = UCSA_MAX_CLASSES;  // semantic output layer is always 48 wide (pad16 of 33..48 classes)
constexpr int kColorW1 = 0, kColorW2 = 64 * 32, kColorW3 = kColorW2 + 64 * 64;  // offsets in w_color
constexpr int kSemW1 = 0, kSemW2 = 64 * 16;                                      // offsets in w_sem

// shared-memory weight tiles
constexpr uint32_t kWc1 = 64 / 8 * Tile<32>::kGroupBytes;
constexpr uint32_t kWc2 = 64 / 8 * Tile<64>::kGroupBytes;
constexpr uint32_t kWc3 = 16 / 8 * Tile<64>::kGroupBytes;
constexpr uint32_t kWs1 = 64 / 8 * Tile<16>::kGroupBytes;
constexpr uint32_t kWs2 = kSemOut / 8 * Tile<64>::kGroupBytes;
constexpr uint32_t kColorWeightBytesFwd = kWc1 + kWc2 + kWc3, kSemWeightBytesFwd = kWs1 + kWs2;

// Raw per-row inputs of the heads, loaded one tile AHEAD into registers: sel -> (h row, ray direction) is a chain of
// dependent global loads, and with only a few CTAs per SM nothing else hides it (it was 45 % of the stall samples of
// the colour backward kernel).  The loop holds the inputs of the current tile, issues the loads of the next one and
// reads sel[] two tiles ahead.
struct RowInputs {
  uint4 h_lo, h_hi;  // the fp16 h row [log-density | geo_feat(15)]
  float dir[3];
};

__device__ __forceinline__ void load_row_inputs(RowInputs& x, const float* __restrict__ rays_d,
                                                const __half* __restrict__ h, uint32_t flat, uint32_t t, bool valid,
                                                bool want_dir) {
  x.h_lo = x.h_hi = make_uint4(0, 0, 0, 0);
  x.dir[0] = x.dir[1] = x.dir[2] = 0.f;
  if (!valid) return;
  x.h_lo = __ldg(reinterpret_cast<const uint4*>(h + static_cast<uint64_t>(flat) * 16));
  x.h_hi = __ldg(reinterpret_cast<const uint4*>(h + static_cast<uint64_t>(flat) * 16 + 8));
  if (want_dir) {
    const uint32_t n = flat / t;
#pragma unroll
    for (int d = 0; d < 3; ++d) x.dir[d] = __ldg(rays_d + 3 * n + d);
  }
}

// [geo_feat(15) | 1] as two 16-byte chunks from a loaded h row
__device__ __forceinline__ void geo_from_row(const RowInputs& x, bool valid, H8& g0, H8& g1) {
  if (!valid) {
    g0.v = g1.v = make_uint4(0, 0, 0, 0);
    return;
  }
  H8 lo, hi;
  lo.v = x.h_lo;
  hi.v = x.h_hi;
#pragma unroll
  for (int i = 0; i < 7; ++i) g0.h[i] = lo.h[i + 1];
  g0.h[7] = hi.h[0];
#pragma unroll
  for (int i = 0; i < 7; ++i) g1.h[i] = hi.h[i + 1];
  g1.h[7] = __float2half_rn(1.0f);
}

// colour input row [SH(16) | geo_feat(15) | 1] and semantic input row [geo_feat(15) | 1] of this thread's row
__device__ __forceinline__ void write_inputs(const RowInputs& x, bool valid, unsigned char* t_in_c,
                                             unsigned char* t_in_s) {
  const int row = threadIdx.x;
  H8 sh_lo, sh_hi, g0, g1;
  geo_from_row(x, valid, g0, g1);
  if (valid && t_in_c != nullptr) {
    float sh[16];
    sh4_eval(x.dir[0], x.dir[1], x.dir[2], sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sh_lo.h[i] = __float2half_rn(sh[i]);
      sh_hi.h[i] = __float2half_rn(sh[8 + i]);
    }
  } else {
    sh_lo.v = sh_hi.v = make_uint4(0, 0, 0, 0);
  }
  if (t_in_c != nullptr) {
    *Tile<32>::chunk(t_in_c, row, 0) = sh_lo.v;
    *Tile<32>::chunk(t_in_c, row, 1) = sh_hi.v;
    *Tile<32>::chunk(t_in_c, row, 2) = g0.v;
    *Tile<32>::chunk(t_in_c, row, 3) = g1.v;
  }
  if (t_in_s != nullptr) {
    *Tile<16>::chunk(t_in_s, row, 0) = g0.v;
    *Tile<16>::chunk(t_in_s, row, 1) = g1.v;
  }
}

// compact row r of tile `tile` -> its flat sample index (0 past the end)
__device__ __forceinline__ uint32_t flat_of(const int32_t* __restrict__ sel, uint32_t tile, uint32_t n_tiles,
                                            uint32_t k_rows) {
  const uint32_t r = tile * 128 + threadIdx.x;
  return (tile < n_tiles && r < k_rows) ? static_cast<uint32_t>(__ldg(sel + r)) : 0u;
}

// ---------------------------------------------------------------------------------------------- forward
// Two kernels (colour, semantics), like the backward pass: each keeps only its own weights and tiles, so that five /
// six CTAs are resident per SM instead of three and the dependent stages of more tiles overlap.
//   colour   : weights 14 KB, tiles in_c 8 KB + h 16 KB (h1, later h2);            TMEM 64 columns; 5 CTAs / SM
//   semantics: weights  8 KB, tiles in_s 4 KB + hs 16 KB, re-used (plus 7 KB) as the fp32 staging `part` of the
//              fused compositing;                                                   TMEM 64 columns; 5 CTAs / SM
// Fused compositing (renderer_semantics.py:279-285): every thread stages w * value of its row, then one thread per
// (row segment, column) walks the rows - the rows of a ray are contiguous - and flushes one atomicAdd per
// (ray, column) run.
constexpr uint32_t kFwdTmemCols = 64;
constexpr int kFwdColorCtas = 5, kFwdSemCtas = 5;
constexpr int kPartLd = kSemOut + 1;  // odd stride: conflict-free row-wise writes and column-wise reads
constexpr uint32_t kPartBytes = 128 * kPartLd * sizeof(float);
constexpr uint32_t kFwdColorSmem = kColorWeightBytesFwd + Tile<32>::kBytes + Tile<64>::kBytes + 128 * 4 * sizeof(float) +
                                   128 * sizeof(int) + 64;
constexpr uint32_t kFwdSemSmem = kSemWeightBytesFwd + kPartBytes + 128 * sizeof(int) + 64;
static_assert(Tile<16>::kBytes + Tile<64>::kBytes <= kPartBytes, "the semantic tiles live inside the staging region");
static_assert(kFwdColorCtas * (kFwdColorSmem + 1024) <= 227 * 1024 && kFwdSemCtas * (kFwdSemSmem + 1024) <= 227 * 1024,
              "forward heads: shared memory of the resident CTAs");

// Run structure of a tile for the fused compositing.  Rows of a ray are contiguous in the compact list, so a tile is
// a few runs of equal ray id.  Each thread publishes its row's ray id; bit `row` of heads[row / 32] says "this row
// starts a new run" (the id of the previous row comes from the neighbouring lane, or from sel[] for a warp's first
// lane).  Rows past K carry id -1 and form a last run that is skipped.
__device__ __forceinline__ void publish_runs(const int32_t* __restrict__ sel, uint32_t r, uint32_t k_rows, uint32_t t,
                                             int id, int* ray_of_row, uint32_t* heads) {
  const int lane = threadIdx.x & 31;
  int prev = __shfl_up_sync(kFullMask, id, 1);
  if (lane == 0) {
    if (threadIdx.x == 0) prev = -2;  // the first row of a tile always starts a run
    else prev = (r - 1 < k_rows) ? static_cast<int>(static_cast<uint32_t>(sel[r - 1]) / t) : -1;
  }
  const uint32_t mask = __ballot_sync(kFullMask, id != prev);
  ray_of_row[threadIdx.x] = id;
  if (lane == 0) heads[threadIdx.x >> 5] = mask;
}

// first row in (a, r1) that starts a run, or r1
__device__ __forceinline__ int next_run(const uint32_t* heads, int a, int r1) {
  int row = a + 1;
  while (row < r1) {
    const uint32_t w = heads[row >> 5] >> (row & 31);
    if (w != 0u) {
      row += __ffs(w) - 1;
      return row < r1 ? row : r1;
    }
    row = (row | 31) + 1;
  }
  return r1;
}

// dst[ray][col] += sum over the ray's rows of part[row][col]: one thread per (row segment, column); inside a run the
// sum is branch-free with four independent accumulators, one atomicAdd per (run, column, segment)
__device__ __forceinline__ void composite_flush(const float* part, int ld, const int* ray_of_row, const uint32_t* heads,
                                                int n_col, float* __restrict__ dst, int dst_ld) {
  const int n_seg = 128 / n_col > 4 ? 4 : (128 / n_col > 0 ? 128 / n_col : 1);  // 40 classes: 3 segments
  const int rows = (128 + n_seg - 1) / n_seg;
  const int seg = threadIdx.x / n_col, col = threadIdx.x % n_col;
  if (seg >= n_seg) return;
  const int r0 = seg * rows, r1 = r0 + rows < 128 ? r0 + rows : 128;
  int a = r0;
  while (a < r1) {
    const int b = next_run(heads, a, r1);
    const int id = ray_of_row[a];
    if (id >= 0) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      const float* p = part + a * ld + col;
      int rr = a;
      for (; rr + 3 < b; rr += 4, p += 4 * ld) {
        acc0 += p[0];
        acc1 += p[ld];
        acc2 += p[2 * ld];
        acc3 += p[3 * ld];
      }
      for (; rr < b; ++rr, p += ld) acc0 += p[0];
      atomicAdd(dst + static_cast<uint64_t>(id) * dst_ld + col, (acc0 + acc1) + (acc2 + acc3));
    }
    a = b;
  }
}

__global__ void __launch_bounds__(128, kFwdColorCtas)
heads_fwd_color_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                       const float* __restrict__ rays_d, const __half* __restrict__ h,
                       const __half* __restrict__ w_color, const float* __restrict__ w_sel, float* __restrict__ rgb,
                       __half* __restrict__ hc1, __half* __restrict__ hc2, float* __restrict__ image) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wc1 = smem;
  unsigned char* wc2 = wc1 + kWc1;
  unsigned char* wc3 = wc2 + kWc2;
  unsigned char* t_in_c = wc3 + kWc3;
  unsigned char* t_h = t_in_c + Tile<32>::kBytes;  // h1, then h2 (layer 2 has consumed h1 by then)
  float* part = reinterpret_cast<float*>(t_h + Tile<64>::kBytes);  // [128][4]: w * rgb
  int* ray_of_row = reinterpret_cast<int*>(part + 128 * 4);
  unsigned char* tail = reinterpret_cast<unsigned char*>(ray_of_row + 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 8);
  uint32_t* heads = reinterpret_cast<uint32_t*>(tail + 16);  // [4] run-start bits of the tile's rows
  umma::load_weight_tile<32>(wc1, w_color + kColorW1, 64);
  umma::load_weight_tile<64>(wc2, w_color + kColorW2, 64);
  umma::load_weight_tile<64>(wc3, w_color + kColorW3, 16);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kFwdTmemCols);
  const uint32_t s_in_c = umma::smem_u32(t_in_c), s_h = umma::smem_u32(t_h), b1 = umma::smem_u32(wc1),
                 b2 = umma::smem_u32(wc2), b3 = umma::smem_u32(wc3);
  const bool fuse_composite = image != nullptr;
  const uint64_t stream = l2_policy_stream();

  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + 127) / 128;
  // row inputs one tile ahead, sel[] two tiles ahead (see RowInputs)
  uint32_t flat = flat_of(sel, blockIdx.x, n_tiles, k_rows);
  uint32_t flat_next = flat_of(sel, blockIdx.x + gridDim.x, n_tiles, k_rows);
  RowInputs in_cur, in_next;
  load_row_inputs(in_cur, rays_d, h, flat, t, blockIdx.x * 128 + threadIdx.x < k_rows, true);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < k_rows;
    const float w_row = (valid && fuse_composite) ? w_sel[r] : 0.f;
    write_inputs(in_cur, valid, t_in_c, nullptr);
    {  // only now the next tile's loads: issued before the consumer above they would share its scoreboard and be
       // waited for right here
      const uint32_t tile_next = tile + gridDim.x;
      load_row_inputs(in_next, rays_d, h, flat_next, t, tile_next < n_tiles && tile_next * 128 + threadIdx.x < k_rows,
                      true);
    }
    const uint32_t flat_next2 = flat_of(sel, tile + 2 * gridDim.x, n_tiles, k_rows);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<32, 64>(ctx.tmem, s_in_c, b1);
      umma::commit(ctx.bar);
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, true>(ctx, c0, t_h, c0);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<64, 64>(ctx.tmem, s_h, b2);
      if (hc1 != nullptr) {  // the commit waits for the copy's reads: the next epilogue overwrites the tile
        umma::bulk_store(umma::tile_block<64>(hc1, tile), s_h, Tile<64>::kBytes, stream);
        umma::bulk_store_fence_reads();
      }
      umma::commit(ctx.bar);
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, true>(ctx, c0, t_h, c0);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<64, 16>(ctx.tmem, s_h, b3);
      if (hc2 != nullptr) {
        umma::bulk_store(umma::tile_block<64>(hc2, tile), s_h, Tile<64>::kBytes, stream);
        umma::bulk_store_fence_reads();
      }
      umma::commit(ctx.bar);
    }
    ctx.wait();
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(0), v);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = round_h(v[c]);
      const float col = round_h(1.0f / (1.0f + expf(-x)));  // fp16 sigmoid under autocast
      if (valid) rgb[static_cast<uint64_t>(r) * 3 + c] = col;
      if (fuse_composite) part[threadIdx.x * 4 + c] = w_row * col;
    }
    if (fuse_composite) {  // image_n = sum_rows w * rgb
      publish_runs(sel, r, k_rows, t, valid ? static_cast<int>(flat / t) : -1, ray_of_row, heads);
      __syncthreads();
      composite_flush(part, 4, ray_of_row, heads, 3, image, 3);
      // the next tile reaches two CTA barriers before it writes part / ray_of_row again
    }
    in_cur = in_next;
    flat = flat_next;
    flat_next = flat_next2;
  }
  umma::ctx_free(ctx, kFwdTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised form of the colour forward kernel: 4 worker warps (thread = row = TMEM lane) + 1 driver warp, TWO
// tiles in flight per CTA (slots A / B with their own tiles and 64 TMEM columns each).  The workers alternate between
// the slots in program order -- inputs A, inputs B, epilogue-1 A, epilogue-1 B, ... -- and one lane of the driver warp
// issues every MMA, commit and bulk store, so while the tensor core (and the commit -> mbarrier round trip, and the
// bulk store's read of a tile) work on one slot the workers compute on the other, with no CTA-wide barrier in the
// loop.  Hand-offs are mbarriers: in_full[s] (128 worker arrivals: "the operand tile of slot s is written"),
// acc_full[s] (tcgen05.commit: "the accumulator of slot s is complete"), st_done[s] (driver: "the bulk store has read
// the tile").  The single-tile kernel above overlaps stages only across CTAs (5 per SM) and stalls all 128 threads on
// thread 0's MMA issue / store wait; here 3 CTAs = 6 tiles are in flight per SM.
constexpr int kWsThreads = 160, kWsCtas = 3;
constexpr uint32_t kWsSlotBytes = Tile<32>::kBytes + Tile<64>::kBytes + 128 * 4 * sizeof(float) + 128 * sizeof(int) + 128;
constexpr uint32_t kFwdColorWsSmem = kColorWeightBytesFwd + 2 * kWsSlotBytes + 6 * 8 + 16;
static_assert(kWsCtas * (kFwdColorWsSmem + 1024) <= 227 * 1024, "colour forward (warp-specialised): shared memory");

__global__ void __launch_bounds__(kWsThreads, kWsCtas)
heads_fwd_color_ws_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                          const float* __restrict__ rays_d, const __half* __restrict__ h,
                          const __half* __restrict__ w_color, const float* __restrict__ w_sel,
                          float* __restrict__ rgb, __half* __restrict__ hc1, __half* __restrict__ hc2,
                          float* __restrict__ image) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wc1 = smem;
  unsigned char* wc2 = wc1 + kWc1;
  unsigned char* wc3 = wc2 + kWc2;
  unsigned char* slots = wc3 + kWc3;
  unsigned char* tail = slots + 2 * kWsSlotBytes;
  uint64_t* in_full = reinterpret_cast<uint64_t*>(tail);  // [2]
  uint64_t* acc_full = in_full + 2;                        // [2]
  uint64_t* st_done = acc_full + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  umma::load_weight_tile<32>(wc1, w_color + kColorW1, 64);
  umma::load_weight_tile<64>(wc2, w_color + kColorW2, 64);
  umma::load_weight_tile<64>(wc3, w_color + kColorW3, 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&in_full[s], 128);
      umma::mbar_init(&acc_full[s], 1);
      umma::mbar_init(&st_done[s], 1);
    }
  }
  if (warp == 4) umma::tmem_alloc(tmem_slot, 128);
  umma::fence_async_smem();  // the weight tiles are read by the tensor core (async proxy)
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool save = hc1 != nullptr;
  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + 127) / 128;
  const uint32_t n_pairs = (n_tiles + 1) / 2;

  if (warp == 4) {
    // ------------------------------------------------------------ driver: MMA issue, commits, bulk stores
    if (lane == 0) {
      const uint32_t b1 = umma::smem_u32(wc1), b2 = umma::smem_u32(wc2), b3 = umma::smem_u32(wc3);
      const uint64_t stream = l2_policy_stream();
      uint32_t ph_in[2] = {0u, 0u};
      for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll
        for (int stage = 0; stage < 3; ++stage) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint32_t tile = 2 * pair + s;
            unsigned char* base = slots + s * kWsSlotBytes;
            const uint32_t s_in_c = umma::smem_u32(base), s_h = umma::smem_u32(base + Tile<32>::kBytes);
            umma::mbar_wait(&in_full[s], ph_in[s]);
            ph_in[s] ^= 1u;
            umma::tc_fence_after();
            if (stage == 0) umma::issue_fwd<32, 64>(tmem + 64 * s, s_in_c, b1);
            else if (stage == 1) umma::issue_fwd<64, 64>(tmem + 64 * s, s_h, b2);
            else umma::issue_fwd<64, 16>(tmem + 64 * s, s_h, b3);
            if (save && stage > 0) {  // one bulk group per slot (empty for a tile past the end)
              if (tile < n_tiles)
                umma::bulk_store(umma::tile_block<64>(stage == 1 ? hc1 : hc2, tile), s_h, Tile<64>::kBytes, stream);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            umma::commit(&acc_full[s]);
          }
          if (save && stage > 0) {  // both slots' products are issued: now wait for the stores to have read their tiles
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            umma::mbar_arrive(&st_done[0]);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            umma::mbar_arrive(&st_done[1]);
          }
        }
      }
      if (save) umma::bulk_store_fence_reads();
    }
  } else {
    // ------------------------------------------------------------ workers: rows of both slots
    const uint64_t stream = l2_policy_stream();
    (void)stream;
    const bool fuse_composite = image != nullptr;
    uint32_t ph_acc[2] = {0u, 0u}, ph_st[2] = {0u, 0u};
    bool st_pending[2] = {false, false};  // the hc2 store of the previous pair may still be reading t_h
    // tile index of (pair, slot) -> flat sample index of this thread's row; inputs one pair ahead, sel two pairs ahead
    auto flat_at = [&](uint32_t pair_, int s_) { return flat_of(sel, 2 * pair_ + s_, n_tiles, k_rows); };
    uint32_t flat[2], flat_next[2];
    RowInputs in_cur[2], in_next[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      flat[s] = flat_at(blockIdx.x, s);
      flat_next[s] = flat_at(blockIdx.x + gridDim.x, s);
      const uint32_t tile = 2 * blockIdx.x + s;
      load_row_inputs(in_cur[s], rays_d, h, flat[s], t, tile < n_tiles && tile * 128 + threadIdx.x < k_rows, true);
    }
    for (uint32_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      bool valid[2];
      float w_row[2];
      uint32_t r[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // stage 0 operands
        const uint32_t tile = 2 * pair + s;
        r[s] = tile * 128 + threadIdx.x;
        valid[s] = tile < n_tiles && r[s] < k_rows;
        w_row[s] = (valid[s] && fuse_composite) ? w_sel[r[s]] : 0.f;
        write_inputs(in_cur[s], valid[s], slots + s * kWsSlotBytes, nullptr);
        umma::fence_async_smem();
        umma::tc_fence_before();
        umma::mbar_arrive(&in_full[s]);
      }
      uint32_t flat_next2[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // the next pair's loads, after this pair's values were consumed
        const uint32_t tile_next = 2 * (pair + gridDim.x) + s;
        load_row_inputs(in_next[s], rays_d, h, flat_next[s], t,
                        tile_next < n_tiles && tile_next * 128 + threadIdx.x < k_rows, true);
        flat_next2[s] = flat_at(pair + 2 * gridDim.x, s);
      }
#pragma unroll
      for (int stage = 0; stage < 2; ++stage) {  // hidden layers: accumulator -> ReLU -> fp16 tile of the next product
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          unsigned char* t_h = slots + s * kWsSlotBytes + Tile<32>::kBytes;
          umma::mbar_wait(&acc_full[s], ph_acc[s]);
          ph_acc[s] ^= 1u;
          if (st_pending[s]) {  // t_h is rewritten below: the store that was issued from it has to have read it
            umma::mbar_wait(&st_done[s], ph_st[s]);
            ph_st[s] ^= 1u;
          }
          st_pending[s] = save;  // this stage's product is followed by a store of the tile written now
          umma::tc_fence_after();
          umma::Ctx ctx{tmem + 64u * s, nullptr, 0u};
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, true>(ctx, c0, t_h, c0);
          umma::fence_async_smem();
          umma::tc_fence_before();
          umma::mbar_arrive(&in_full[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // output layer + compositing
        unsigned char* base = slots + s * kWsSlotBytes;
        float* part = reinterpret_cast<float*>(base + Tile<32>::kBytes + Tile<64>::kBytes);
        int* ray_of_row = reinterpret_cast<int*>(part + 128 * 4);
        uint32_t* heads = reinterpret_cast<uint32_t*>(ray_of_row + 128);
        umma::mbar_wait(&acc_full[s], ph_acc[s]);
        ph_acc[s] ^= 1u;
        umma::tc_fence_after();
        umma::Ctx ctx{tmem + 64u * s, nullptr, 0u};
        float v[16];
        umma::tmem_ld16(ctx.lane_addr(0), v);
        umma::tc_fence_before();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float x = round_h(v[c]);
          const float col = round_h(1.0f / (1.0f + expf(-x)));  // fp16 sigmoid under autocast
          if (valid[s]) rgb[static_cast<uint64_t>(r[s]) * 3 + c] = col;
          if (fuse_composite) part[threadIdx.x * 4 + c] = w_row[s] * col;
        }
        if (fuse_composite) {
          publish_runs(sel, r[s], k_rows, t, valid[s] ? static_cast<int>(flat[s] / t) : -1, ray_of_row, heads);
          umma::named_barrier(1, 128);
          composite_flush(part, 4, ray_of_row, heads, 3, image, 3);
          // part / ray_of_row of this slot are rewritten one pair later, after three hand-offs that need all workers
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        in_cur[s] = in_next[s];
        flat[s] = flat_next[s];
        flat_next[s] = flat_next2[s];
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 4) umma::tmem_dealloc(tmem, 128);
}

// NC: the class count as a compile-time constant (40 = the reference's configuration: the soft-max / staging loops
// lose their per-column predicates and the 8 padded columns altogether), 0 = read n_classes_rt at run time
template <int NC>
__global__ void __launch_bounds__(128, kFwdSemCtas)
heads_fwd_sem_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                     const __half* __restrict__ h, const __half* __restrict__ w_sem, int n_classes_rt,
                     const float* __restrict__ w_sel, __half* __restrict__ logits, __half* __restrict__ hs,
                     float* __restrict__ semantics) {
  const int n_classes = NC > 0 ? NC : n_classes_rt;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ws1 = smem;
  unsigned char* ws2 = ws1 + kWs1;
  unsigned char* t_in_s = ws2 + kWs2;                     // staging region: [in_s 4 KB | hs 16 KB | ...]
  unsigned char* t_hs = t_in_s + Tile<16>::kBytes;
  float* part = reinterpret_cast<float*>(t_in_s);         // ... re-used as w * softmax once both tiles are consumed
  int* ray_of_row = reinterpret_cast<int*>(t_in_s + kPartBytes);
  unsigned char* tail = reinterpret_cast<unsigned char*>(ray_of_row + 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 8);
  uint32_t* heads = reinterpret_cast<uint32_t*>(tail + 16);
  umma::load_weight_tile<16>(ws1, w_sem + kSemW1, 64);
  umma::load_weight_tile<64>(ws2, w_sem + kSemW2, kSemOut);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kFwdTmemCols);
  const uint32_t s_in_s = umma::smem_u32(t_in_s), s_hs = umma::smem_u32(t_hs), b1 = umma::smem_u32(ws1),
                 b2 = umma::smem_u32(ws2);
  const bool fuse_composite = semantics != nullptr;
  const uint64_t stream = l2_policy_stream();

  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + 127) / 128;
  uint32_t flat = flat_of(sel, blockIdx.x, n_tiles, k_rows);
  uint32_t flat_next = flat_of(sel, blockIdx.x + gridDim.x, n_tiles, k_rows);
  RowInputs in_cur, in_next;
  load_row_inputs(in_cur, nullptr, h, flat, t, blockIdx.x * 128 + threadIdx.x < k_rows, false);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < k_rows;
    const float w_row = (valid && fuse_composite) ? w_sel[r] : 0.f;
    write_inputs(in_cur, valid, nullptr, t_in_s);
    {  // after the consumer (scoreboards, see the colour kernel)
      const uint32_t tile_next = tile + gridDim.x;
      load_row_inputs(in_next, nullptr, h, flat_next, t, tile_next < n_tiles && tile_next * 128 + threadIdx.x < k_rows,
                      false);
    }
    const uint32_t flat_next2 = flat_of(sel, tile + 2 * gridDim.x, n_tiles, k_rows);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<16, 64>(ctx.tmem, s_in_s, b1);
      umma::commit(ctx.bar);
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, true>(ctx, c0, t_hs, c0);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_fwd<64, kSemOut>(ctx.tmem, s_hs, b2);
      if (hs != nullptr) {  // the commit waits for the copy's reads: `part` overwrites the tile below
        umma::bulk_store(umma::tile_block<64>(hs, tile), s_hs, Tile<64>::kBytes, stream);
        umma::bulk_store_fence_reads();
      }
      umma::commit(ctx.bar);
    }
    ctx.wait();
    {
      // the row's logits: fp16 like the reference's network output, kept in registers for the soft-max
      float lg[kSemOut];
#pragma unroll
      for (int c0 = 0; c0 < kSemOut; c0 += 16) {
        float v[16];
        umma::tmem_ld16(ctx.lane_addr(c0), v);
        H8 a, b;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a.h[i] = __float2half_rn(v[i]);
          b.h[i] = __float2half_rn(v[8 + i]);
          lg[c0 + i] = __half2float(a.h[i]);
          lg[c0 + 8 + i] = __half2float(b.h[i]);
        }
        if (valid && logits != nullptr) {  // optional output: the backward pass recomputes the logits
          st_stream(logits + static_cast<uint64_t>(r) * kSemOut + c0, a.v, stream);
          st_stream(logits + static_cast<uint64_t>(r) * kSemOut + c0 + 8, b.v, stream);
        }
      }
      if (fuse_composite) {
        // semantics_n = sum_rows w * softmax(logits)   (network_tcnn_semantics.py:202-203, renderer_semantics.py:284)
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < kSemOut; ++c)
          if (c < n_classes) m = fmaxf(m, lg[c]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kSemOut; ++c) {
          if (c < n_classes) {  // (uniform) columns past the class count are never staged nor read
            lg[c] = exp2f_approx((lg[c] - m) * 1.4426950408889634f);  // = __expf(lg - m): one multiply + ex2.approx
            sum += lg[c];
          }
        }
        const float scale = w_row / sum;
        float* my_part = part + threadIdx.x * kPartLd;
#pragma unroll
        for (int c = 0; c < kSemOut; ++c)
          if (c < n_classes) my_part[c] = lg[c] * scale;
      }
    }
    if (fuse_composite) {
      publish_runs(sel, r, k_rows, t, valid ? static_cast<int>(flat / t) : -1, ray_of_row, heads);
      __syncthreads();
      composite_flush(part, kPartLd, ray_of_row, heads, n_classes, semantics, n_classes);
      __syncthreads();  // part[] lives in the input tiles: the next tile may only rebuild them after the reads
    }
    in_cur = in_next;
    flat = flat_next;
    flat_next = flat_next2;
  }
  umma::ctx_free(ctx, kFwdTmemCols);
}

// HI: the accumulator was placed at TMEM lane offset 16 (rows m at lanes 16 + m % 16 + 32 * (m / 16)): an M = 64
// product accepts that address, so two weight-gradient accumulators share the same columns (scripts/dev/
// tmem_lane_probe.cu is the experiment that established it on B200).
template <int N, bool TRANSPOSED, bool HI = false>
__device__ __forceinline__ void flush_wgrad(const umma::Ctx& ctx, uint32_t col0, float* __restrict__ grad, int ld,
                                            float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(ctx.lane_addr(col0 + c0), v);
    if (HI ? lane >= 16 : lane < 16) {
      const int m = warp * 16 + (lane & 15);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        atomicAdd(grad + (TRANSPOSED ? (c0 + i) * ld + m : m * ld + c0 + i), v[i] * scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward
// Two kernels, colour first, then semantics, so that each fits several CTAs per SM:
//   colour   : tiles in_c, h1, h2 (dh2 in place), dh1, dpre (60 KB) + colour weights; TMEM 128 columns (the three
//              weight-gradient accumulators share 64 columns: dWc2 on the low, dWc1 | dWc3 on the high 16 lanes of
//              every 32-lane quadrant);                                                                3 CTAs / SM
//   semantics: tiles hs, dhs, dlog (in_s re-uses dlog) (44 KB) + weights;   TMEM 128 columns;        4 CTAs / SM
// The colour kernel writes its share of dL/dgeo_feat into dh and the semantic kernel adds its own.
constexpr uint32_t kBwdColorCols = 128, kBwdSemCols = 128;
constexpr int kBwdColorCtas = 3, kBwdSemCtas = 4;
constexpr uint32_t kColorWeightBytes = kWc1 + kWc2 + kWc3;
constexpr uint32_t kSemWeightBytes = kWs1 + kWs2;
constexpr uint32_t kBwdColorSmem = kColorWeightBytes + Tile<32>::kBytes + 3 * Tile<64>::kBytes + Tile<16>::kBytes + 64;
static_assert(kBwdColorCtas * (kBwdColorSmem + 1024) <= 227 * 1024, "colour backward: shared memory of the resident CTAs");
constexpr uint32_t kBwdSemSmem = kSemWeightBytes + 2 * Tile<64>::kBytes + Tile<kSemOut>::kBytes + 64;
static_assert(kBwdSemCtas * (kBwdSemSmem + 1024) <= 227 * 1024, "semantic backward: shared memory of the resident CTAs");

__global__ void __launch_bounds__(128, kBwdColorCtas)
heads_bwd_color_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                       const float* __restrict__ rays_d, const __half* __restrict__ h,
                       const __half* __restrict__ w_color, const float* __restrict__ rgb,
                       const __half* __restrict__ hc1, const __half* __restrict__ hc2,
                       const float* __restrict__ w_sel, const float* __restrict__ z_sel,
                       const float* __restrict__ g_image, const float* __restrict__ g_depth,
                       const float* __restrict__ dnorm, float loss_scale, __half* __restrict__ dh,
                       float* __restrict__ d_w_sel, float* __restrict__ grad_w_color) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wc1 = smem;
  unsigned char* wc2 = wc1 + kWc1;
  unsigned char* wc3 = wc2 + kWc2;
  unsigned char* t_in_c = wc3 + kWc3;
  unsigned char* t_h1 = t_in_c + Tile<32>::kBytes;
  unsigned char* t_h2 = t_h1 + Tile<64>::kBytes;
  unsigned char* t_dh1 = t_h2 + Tile<64>::kBytes;
  unsigned char* t_dh2 = t_h2;  // in place: every thread turns its own h2 chunks into dh2 chunks once the products
                                // reading h2 (the Wc3 weight gradient) are covered by the commit it waited for
  unsigned char* t_dpre = t_dh1 + Tile<64>::kBytes;
  unsigned char* tail = t_dpre + Tile<16>::kBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* ld_bar = reinterpret_cast<uint64_t*>(tail + 8);  // completion of the bulk loads of the saved tiles
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 16);
  if (threadIdx.x == 64) umma::mbar_init(ld_bar, 1);
  umma::load_weight_tile<32>(wc1, w_color + kColorW1, 64);
  umma::load_weight_tile<64>(wc2, w_color + kColorW2, 64);
  umma::load_weight_tile<64>(wc3, w_color + kColorW3, 16);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kBwdColorCols);
  const uint32_t s_in_c = umma::smem_u32(t_in_c), s_h1 = umma::smem_u32(t_h1), s_h2 = umma::smem_u32(t_h2),
                 s_dh1 = umma::smem_u32(t_dh1), s_dh2 = umma::smem_u32(t_dh2), s_dpre = umma::smem_u32(t_dpre),
                 b1 = umma::smem_u32(wc1), b2 = umma::smem_u32(wc2), b3 = umma::smem_u32(wc3);
  // TMEM columns: scratch 0..63 | dWc2 64..127 on lanes 0-15 (+32 i) | dWc1 64..95 and dWc3 96..111 on lanes 16-31 (+32 i)
  constexpr uint32_t kHi = 16u << 16;  // TMEM address: lane offset 16
  constexpr uint32_t kAcc = 0, kGc2 = 64, kGc1 = 64, kGc3 = 96;
  const uint64_t stream = l2_policy_stream();
  uint32_t ld_phase = 0;

  const float inv_scale = 1.0f / loss_scale;
  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + 127) / 128;
  const int row = threadIdx.x;
  bool first = true;
  // per-row scalars of the compositing backward, also one tile ahead
  struct Scalars {
    float w, z, gd_over_dn, rgb[3], gi[3];
  };
  auto load_scalars = [&](Scalars& q, uint32_t tile_, uint32_t flat_) {
    const uint32_t r_ = tile_ * 128 + threadIdx.x;
    q = Scalars{};
    if (tile_ < n_tiles && r_ < k_rows) {
      const uint32_t n = flat_ / t;
      q.w = __ldg(w_sel + r_);
      q.z = __ldg(z_sel + r_);
      q.gd_over_dn = __ldg(g_depth + n) / __ldg(dnorm + n);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        q.rgb[c] = __ldg(rgb + static_cast<uint64_t>(r_) * 3 + c);
        q.gi[c] = __ldg(g_image + static_cast<uint64_t>(n) * 3 + c);
      }
    }
  };
  uint32_t flat = flat_of(sel, blockIdx.x, n_tiles, k_rows);
  uint32_t flat_next = flat_of(sel, blockIdx.x + gridDim.x, n_tiles, k_rows);
  RowInputs in_cur, in_next;
  Scalars sc_cur, sc_next;
  load_row_inputs(in_cur, rays_d, h, flat, t, blockIdx.x * 128 + threadIdx.x < k_rows, true);
  load_scalars(sc_cur, blockIdx.x, flat);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < k_rows;
    if (!first) ctx.wait();  // weight-gradient MMAs of the previous tile are done with the tiles
    if (threadIdx.x == 0) {
      umma::mbar_expect_tx(ld_bar, 2 * Tile<64>::kBytes);
      umma::bulk_load(s_h1, umma::tile_block<64>(hc1, tile), Tile<64>::kBytes, ld_bar, stream);
      umma::bulk_load(s_h2, umma::tile_block<64>(hc2, tile), Tile<64>::kBytes, ld_bar, stream);
    }
    write_inputs(in_cur, valid, t_in_c, nullptr);
    {
      // backward of image_n = sum w rgb and depth_n = sum w z / dn (renderer_semantics.py:276-282):
      //   d rgb = w g_image ;  d w = g_image . rgb + g_depth z / dn ; then through the sigmoid: s (1 - s)
      H8 lo, hi;
      lo.v = make_uint4(0, 0, 0, 0);
      hi.v = make_uint4(0, 0, 0, 0);
      if (valid) {
        float dw = sc_cur.gd_over_dn * sc_cur.z;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sg = sc_cur.rgb[c];
          const float gi = sc_cur.gi[c];
          dw = fmaf(gi, sg, dw);
          lo.h[c] = __float2half_rn(sc_cur.w * gi * sg * (1.0f - sg) * loss_scale);
        }
        d_w_sel[r] = dw;
      }
      *Tile<16>::chunk(t_dpre, row, 0) = lo.v;
      *Tile<16>::chunk(t_dpre, row, 1) = hi.v;
    }
    {  // the next tile's loads, issued AFTER this tile's values are consumed (before them they would share the
       // consumers' scoreboards and be waited for at once); they land while this tile is processed
      const uint32_t tile_next = tile + gridDim.x;
      const bool valid_next = tile_next < n_tiles && tile_next * 128 + threadIdx.x < k_rows;
      load_row_inputs(in_next, rays_d, h, flat_next, t, valid_next, true);
      load_scalars(sc_next, tile_next, flat_next);
    }
    const uint32_t flat_next2 = flat_of(sel, tile + 2 * gridDim.x, n_tiles, k_rows);
    ctx.publish();
    umma::mbar_wait(ld_bar, ld_phase);  // h1 and h2 have landed (operands of the products, ReLU masks below)
    ld_phase ^= 1u;
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<16, 64>(ctx.tmem + kAcc, s_dpre, b3);
      umma::issue_wgrad<16>(ctx.tmem + kGc3 + kHi, s_h2, s_dpre, first);  // d(Wc3)^T = h2^T . dpre
      umma::commit(ctx.bar);  // covers the weight gradient too: the epilogue below overwrites h2 in place
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, false>(ctx, kAcc + c0, t_dh2, c0, t_h2);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<64, 64>(ctx.tmem + kAcc, s_dh2, b2);
      umma::commit(ctx.bar);
      umma::issue_wgrad<64>(ctx.tmem + kGc2, s_dh2, s_h1, first);  // d(Wc2) = dh2^T . h1
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, false>(ctx, kAcc + c0, t_dh1, c0, t_h1);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<64, 32>(ctx.tmem + kAcc, s_dh1, b1);
      umma::commit(ctx.bar);
      umma::issue_wgrad<32>(ctx.tmem + kGc1 + kHi, s_dh1, s_in_c, first);  // d(Wc1) = dh1^T . in_c
    }
    ctx.wait();
    float d_in_c[16];
    umma::tmem_ld16(ctx.lane_addr(kAcc + 16), d_in_c);  // columns 16..31 = the geo_feat (+1) inputs
    if (valid) {  // colour share of dL/dgeo_feat (fp16, still scaled); the semantic kernel adds its own
      H8 lo, hi;
      lo.h[0] = __float2half_rn(0.f);
#pragma unroll
      for (int i = 0; i < 7; ++i) lo.h[i + 1] = __float2half_rn(d_in_c[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) hi.h[i] = __float2half_rn(d_in_c[7 + i]);
      reinterpret_cast<uint4*>(dh + static_cast<uint64_t>(flat) * 16)[0] = lo.v;
      reinterpret_cast<uint4*>(dh + static_cast<uint64_t>(flat) * 16)[1] = hi.v;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::commit(ctx.bar);  // covers the weight-gradient products of this tile
    }
    first = false;
    in_cur = in_next;
    sc_cur = sc_next;
    flat = flat_next;
    flat_next = flat_next2;
  }
  if (!first) {
    ctx.wait();
    flush_wgrad<32, false, true>(ctx, kGc1, grad_w_color + kColorW1, 32, inv_scale);
    flush_wgrad<64, false>(ctx, kGc2, grad_w_color + kColorW2, 64, inv_scale);
    flush_wgrad<16, true, true>(ctx, kGc3, grad_w_color + kColorW3, 64, inv_scale);
  }
  umma::ctx_free(ctx, kBwdColorCols);
}

template <int NC>
__global__ void __launch_bounds__(128, kBwdSemCtas)
heads_bwd_sem_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ k_ptr, uint32_t t,
                     const float* __restrict__ rays_d, const __half* __restrict__ h, const __half* __restrict__ w_sem,
                     int n_classes_rt, const __half* __restrict__ hs,
                     const float* __restrict__ w_sel, const float* __restrict__ g_sem, float loss_scale,
                     __half* __restrict__ dh, __half* __restrict__ dh_sem, float* __restrict__ grad_w_sem) {
  const int n_classes = NC > 0 ? NC : n_classes_rt;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ws1 = smem;
  unsigned char* ws2 = ws1 + kWs1;
  unsigned char* t_hs = ws2 + kWs2;
  unsigned char* t_dhs = t_hs + Tile<64>::kBytes;
  unsigned char* t_dlog = t_dhs + Tile<64>::kBytes;
  unsigned char* t_in_s = t_dlog;  // built once every product reading dlog has completed (four CTAs per SM fit)
  unsigned char* tail = t_dlog + Tile<kSemOut>::kBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* ld_bar = reinterpret_cast<uint64_t*>(tail + 8);
  uint32_t* slot = reinterpret_cast<uint32_t*>(tail + 16);
  if (threadIdx.x == 64) umma::mbar_init(ld_bar, 1);
  umma::load_weight_tile<16>(ws1, w_sem + kSemW1, 64);
  umma::load_weight_tile<64>(ws2, w_sem + kSemW2, kSemOut);
  umma::Ctx ctx = umma::ctx_init(slot, bar, kBwdSemCols);
  const uint32_t s_in_s = umma::smem_u32(t_in_s), s_hs = umma::smem_u32(t_hs), s_dhs = umma::smem_u32(t_dhs),
                 s_dlog = umma::smem_u32(t_dlog), b1 = umma::smem_u32(ws1), b2 = umma::smem_u32(ws2);
  constexpr uint32_t kAcc = 0, kGs1 = 64, kGs2 = 80;
  const uint64_t stream = l2_policy_stream();
  uint32_t ld_phase = 0;

  const float inv_scale = 1.0f / loss_scale;
  const uint32_t k_rows = static_cast<uint32_t>(*k_ptr);
  const uint32_t n_tiles = (k_rows + 127) / 128;
  const int row = threadIdx.x;
  bool first = true;
  // Latency plan: sel[] is read two tiles ahead and the hs tile of the next tile is requested as soon as this tile's
  // products have released the buffer, so neither the sel -> h chain nor the bulk load is waited for at the top of a
  // tile; the row's own h / dh values and the ray's output gradients are requested early and used late.
  uint32_t flat = flat_of(sel, blockIdx.x, n_tiles, k_rows);
  uint32_t flat_next = flat_of(sel, blockIdx.x + gridDim.x, n_tiles, k_rows);
  if (threadIdx.x == 0 && blockIdx.x < n_tiles) {
    umma::mbar_expect_tx(ld_bar, Tile<64>::kBytes);
    umma::bulk_load(s_hs, umma::tile_block<64>(hs, blockIdx.x), Tile<64>::kBytes, ld_bar, stream);
  }
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r = tile * 128 + threadIdx.x;
    const bool valid = r < k_rows;
    const uint32_t flat_next2 = flat_of(sel, tile + 2 * gridDim.x, n_tiles, k_rows);
    if (valid) {  // used in the soft-max block below: pull the ray's row into L1 now (no registers held)
      const float* g_row = g_sem + static_cast<uint64_t>(flat / t) * n_classes;  // the ray's dL/dsemantics
      prefetch_l1(g_row);
      prefetch_l1(g_row + n_classes - 1);
    }
    const float w_row = valid ? __ldg(w_sel + r) : 0.f;
    if (!first) ctx.wait();
    // CTA barrier: two waits on `bar` in a row (above and below) would let a warp that is late for the first one
    // miss its phase - mbarrier parity cannot tell phase k from k+2.  First tile: also publishes the weight tiles.
    ctx.publish();
    umma::mbar_wait(ld_bar, ld_phase);
    ld_phase ^= 1u;
    if (threadIdx.x == 0) {  // the logits again: hs . Ws2^T, the very product of the forward pass
      umma::tc_fence_after();
      umma::issue_fwd<64, kSemOut>(ctx.tmem + kAcc, s_hs, b2);
      umma::commit(ctx.bar);
    }
    ctx.wait();
    {
      // backward of semantics_n = sum w softmax(l) with detached weights (renderer_semantics.py:270,284):
      //   d l = p (w g_sem - <w g_sem, p>)
      float p[kSemOut];
#pragma unroll
      for (int c0 = 0; c0 < kSemOut; c0 += 16) {
        float v[16];
        umma::tmem_ld16(ctx.lane_addr(kAcc + c0), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) p[c0 + i] = round_h(v[i]);  // fp16 network output, as in the forward pass
      }
      if (valid) {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < kSemOut; ++c)
          if (c < n_classes) m = fmaxf(m, p[c]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kSemOut; ++c) {
          // same soft-max arithmetic as the forward kernel
          p[c] = c < n_classes ? exp2f_approx((p[c] - m) * 1.4426950408889634f) : 0.f;
          sum += p[c];
        }
        const float inv = 1.0f / sum;
        float dot = 0.f;
        float q[kSemOut];
        const float* g_row = g_sem + static_cast<uint64_t>(flat / t) * n_classes;
        if (NC > 0 && NC % 4 == 0) {  // the ray's dL/dsemantics row as 16-byte loads (rows are 16-byte aligned)
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(g_row + c));
            q[c] = w_row * g4.x, q[c + 1] = w_row * g4.y, q[c + 2] = w_row * g4.z, q[c + 3] = w_row * g4.w;
          }
#pragma unroll
          for (int c = NC; c < kSemOut; ++c) q[c] = 0.f;
        } else {
#pragma unroll
          for (int c = 0; c < kSemOut; ++c) q[c] = c < n_classes ? w_row * __ldg(g_row + c) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < kSemOut; ++c) {
          p[c] *= inv;
          dot = fmaf(q[c], p[c], dot);
        }
#pragma unroll
        for (int c = 0; c < kSemOut; ++c) p[c] = p[c] * (q[c] - dot) * loss_scale;
      } else {
#pragma unroll
        for (int c = 0; c < kSemOut; ++c) p[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < kSemOut / 8; ++c) {
        H8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.h[i] = __float2half_rn(p[c * 8 + i]);
        *Tile<kSemOut>::chunk(t_dlog, row, c) = o.v;
      }
    }
    // the row's h and dh values are needed after the two products below: request them now, while few registers
    // are live (plain loads for dh: the colour kernel wrote it)
    RowInputs in_row;
    load_row_inputs(in_row, nullptr, h, flat, t, valid, false);
    uint4 dh_lo = make_uint4(0, 0, 0, 0), dh_hi = make_uint4(0, 0, 0, 0);
    if (valid && dh_sem == nullptr) {  // (with dh_sem the semantic share goes to its own buffer: nothing to read)
      dh_lo = reinterpret_cast<const uint4*>(dh + static_cast<uint64_t>(flat) * 16)[0];
      dh_hi = reinterpret_cast<const uint4*>(dh + static_cast<uint64_t>(flat) * 16)[1];
    }
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<kSemOut, 64>(ctx.tmem + kAcc, s_dlog, b2);
      umma::commit(ctx.bar);
      umma::issue_wgrad<kSemOut>(ctx.tmem + kGs2, s_hs, s_dlog, first);  // d(Ws2)^T = hs^T . dlogits
    }
    ctx.wait();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) umma::acc_to_tile16<64, false>(ctx, kAcc + c0, t_dhs, c0, t_hs);
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_dgrad<64, 16>(ctx.tmem + kAcc, s_dhs, b1);
      umma::commit(ctx.bar);
    }
    ctx.wait();  // every product issued so far is complete: the dlog tile may be re-used for the layer-1 input
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) {  // ... and the hs buffer for the next tile's activations
      umma::mbar_expect_tx(ld_bar, Tile<64>::kBytes);
      umma::bulk_load(s_hs, umma::tile_block<64>(hs, tile + gridDim.x), Tile<64>::kBytes, ld_bar, stream);
    }
    {
      H8 g0, g1;
      geo_from_row(in_row, valid, g0, g1);
      *Tile<16>::chunk(t_in_s, row, 0) = g0.v;
      *Tile<16>::chunk(t_in_s, row, 1) = g1.v;
    }
    float d_in_s[16];
    umma::tmem_ld16(ctx.lane_addr(kAcc), d_in_s);
    if (valid) {  // dL/dgeo_feat = colour share (already in dh) + semantic share, handed to density_bwd as fp16
      H8 lo, hi;
      lo.v = dh_lo;
      hi.v = dh_hi;
#pragma unroll
      for (int i = 0; i < 7; ++i) lo.h[i + 1] = __float2half_rn(__half2float(lo.h[i + 1]) + round_h(d_in_s[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) hi.h[i] = __float2half_rn(__half2float(hi.h[i]) + round_h(d_in_s[7 + i]));
      // (dh_sem: 0 + round_h(x) is exact, so density_bwd's half(colour + semantic) reproduces the in-place sum bit for bit)
      __half* out = dh_sem != nullptr ? dh_sem : dh;
      reinterpret_cast<uint4*>(out + static_cast<uint64_t>(flat) * 16)[0] = lo.v;
      reinterpret_cast<uint4*>(out + static_cast<uint64_t>(flat) * 16)[1] = hi.v;
    }
    ctx.publish();
    if (threadIdx.x == 0) {
      umma::tc_fence_after();
      umma::issue_wgrad<16>(ctx.tmem + kGs1, s_dhs, s_in_s, first);  // d(Ws1) = dhs^T . in_s
      umma::commit(ctx.bar);  // waited on at the top of the next tile
    }
    first = false;
    flat = flat_next;
    flat_next = flat_next2;
  }
  if (!first) {
    ctx.wait();
    flush_wgrad<16, false>(ctx, kGs1, grad_w_sem + kSemW1, 16, inv_scale);
    flush_wgrad<kSemOut, true>(ctx, kGs2, grad_w_sem + kSemW2, 64, inv_scale);
  }
  umma::ctx_free(ctx, kBwdSemCols);
}

// Library-owned helper stream per device for the kernel pairs that run concurrently: forked from and joined back to the
// caller's stream with events, so everything stays ordered on the caller's stream and is capturable into a CUDA graph.
std::mutex g_side_mutex;
cudaStream_t g_side[64] = {};
cudaEvent_t g_ev_fork[64] = {}, g_ev_join[64] = {};

int fork_side_stream(cudaStream_t st, cudaStream_t* side_out, int* dev_out) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (g_side[dev] == nullptr) {  // first call on this device (never inside a stream capture: engines warm up first)
    std::lock_guard<std::mutex> lock(g_side_mutex);
    if (g_side[dev] == nullptr) {
      cudaStream_t s_new;
      if (cudaStreamCreateWithFlags(&s_new, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&g_ev_fork[dev], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&g_ev_join[dev], cudaEventDisableTiming) != cudaSuccess)
        return check_launch("helper stream");
      g_side[dev] = s_new;
    }
  }
  cudaEventRecord(g_ev_fork[dev], st);
  cudaStreamWaitEvent(g_side[dev], g_ev_fork[dev], 0);
  *side_out = g_side[dev];
  *dev_out = dev;
  return UCSA_OK;
}
void join_side_stream(cudaStream_t st, int dev) {
  cudaEventRecord(g_ev_join[dev], g_side[dev]);
  cudaStreamWaitEvent(st, g_ev_join[dev], 0);
}

uint32_t heads_grid(uint32_t k_max, int ctas_per_sm) {
  const uint32_t tiles = (k_max + 127) / 128;
  const uint32_t cap = kNumSMs * ctas_per_sm;
  return tiles < cap ? tiles : cap;
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_heads_fwd(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t,
                              uint32_t k_max, const float* rays_d, const void* h, const void* w_color_h,
                              const void* w_sem_h, uint32_t n_classes, const float* w_sel, float* rgb, void* logits,
                              void* hc1, void* hc2, void* hs, float* image, float* semantics, void* stream) {
  UCSA_REQUIRE(sel && ray_off && rays_d && h && w_color_h && w_sem_h && rgb, "heads_fwd: null pointer");
  UCSA_REQUIRE((hc1 != nullptr) == (hc2 != nullptr) && (hc1 != nullptr) == (hs != nullptr),
               "heads_fwd: pass hc1, hc2 and hs together");
  UCSA_REQUIRE((image == nullptr) == (semantics == nullptr) && (image == nullptr || w_sel != nullptr),
               "heads_fwd: fused compositing needs w_sel, image and semantics together");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "heads_fwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  if (k_max == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_fwd_color_kernel), kFwdColorSmem, "heads_fwd_color_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_fwd_sem_kernel<0>), kFwdSemSmem, "heads_fwd_sem_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_fwd_sem_kernel<40>), kFwdSemSmem, "heads_fwd_sem_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  cudaStream_t st = as_stream(stream);
  static int use_ws = -1;
  if (use_ws < 0) {  // bring-up knob: UCSA_HEADS_WS=1 selects the warp-specialised two-slot colour kernel
    const char* e = getenv("UCSA_HEADS_WS");
    use_ws = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (use_ws) {
    if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_fwd_color_ws_kernel), kFwdColorWsSmem, "heads_fwd_color_ws_kernel")) return rc;
    const uint32_t pairs = ((k_max + 127) / 128 + 1) / 2;
    const uint32_t cap = kNumSMs * kWsCtas;
    heads_fwd_color_ws_kernel<<<pairs < cap ? pairs : cap, kWsThreads, kFwdColorWsSmem, st>>>(
        sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_color_h), w_sel,
        rgb, static_cast<__half*>(hc1), static_cast<__half*>(hc2), image);
    auto sem_kernel = n_classes == 40 ? heads_fwd_sem_kernel<40> : heads_fwd_sem_kernel<0>;
    sem_kernel<<<heads_grid(k_max, kFwdSemCtas), 128, kFwdSemSmem, st>>>(
        sel, ray_off + n_rays, t, static_cast<const __half*>(h), static_cast<const __half*>(w_sem_h),
        static_cast<int>(n_classes), w_sel, static_cast<__half*>(logits), static_cast<__half*>(hs), semantics);
    return check_launch("heads_fwd");
  } else {
    // The two kernels run CONCURRENTLY: the semantic kernel goes to a helper stream forked from / joined to `st` with
    // events (capturable), and each kernel's persistent grid is sized so that both fit an SM at once (2 colour + 3
    // semantic CTAs: 187 KB of shared memory, 320 TMEM columns).  In training the colour kernel is bound by the HBM
    // writes of its saved activations (64 % of peak) and the semantic kernel by instruction issue (58 %), so together
    // they take 0.275 ms instead of 0.316 ms one after the other (bench.py, config 2; (2,4) 0.279, (3,3) 0.283,
    // (2,2) 0.348).  UCSA_FWD_OVERLAP=0 restores the sequential launch; UCSA_FWD_COLOR_CTAS / UCSA_FWD_SEM_CTAS
    // override the grid sizes (bring-up knobs).
    // Without saved activations (inference) the colour kernel is latency-bound like the semantic one and wants a third
    // CTA: full-frame rendering 50.6 ms per 640x480 view with (3,3) against 56.0 with (2,3) and 54.5 one after the
    // other (scripts/render_probe.py).
    static int overlap = -1, color_ctas_env = 0, sem_ctas = 0;
    if (overlap < 0) {
      std::lock_guard<std::mutex> lock(g_side_mutex);
      const char* e = getenv("UCSA_FWD_OVERLAP");
      const int ov = (e != nullptr && e[0] == '0') ? 0 : 1;
      sem_ctas = ov ? 3 : kFwdSemCtas;
      if (const char* c = getenv("UCSA_FWD_COLOR_CTAS")) color_ctas_env = atoi(c) >= 1 && atoi(c) <= 5 ? atoi(c) : 0;
      if (const char* c = getenv("UCSA_FWD_SEM_CTAS")) sem_ctas = atoi(c) >= 1 && atoi(c) <= 5 ? atoi(c) : sem_ctas;
      overlap = ov;
    }
    const int color_ctas = color_ctas_env ? color_ctas_env : (!overlap ? kFwdColorCtas : (hc1 != nullptr ? 2 : 3));
    cudaStream_t st_sem = st;
    int dev = 0;
    if (overlap) {
      if (int rc = fork_side_stream(st, &st_sem, &dev)) return rc;
    }
    heads_fwd_color_kernel<<<heads_grid(k_max, color_ctas), 128, kFwdColorSmem, st>>>(
        sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_color_h), w_sel,
        rgb, static_cast<__half*>(hc1), static_cast<__half*>(hc2), image);
    auto sem_kernel = n_classes == 40 ? heads_fwd_sem_kernel<40> : heads_fwd_sem_kernel<0>;
    sem_kernel<<<heads_grid(k_max, sem_ctas), 128, kFwdSemSmem, st_sem>>>(
        sel, ray_off + n_rays, t, static_cast<const __half*>(h), static_cast<const __half*>(w_sem_h),
        static_cast<int>(n_classes), w_sel, static_cast<__half*>(logits), static_cast<__half*>(hs), semantics);
    if (overlap) join_side_stream(st, dev);
    return check_launch("heads_fwd");
  }
  return check_launch("heads_fwd");
}

extern "C" int ucsa_heads_bwd(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t,
                              uint32_t k_max, const float* rays_d, const void* h, const void* w_color_h,
                              const void* w_sem_h, uint32_t n_classes, const float* rgb,
                              const void* hc1, const void* hc2, const void* hs, const float* w_sel,
                              const float* z_sel, const float* g_image, const float* g_depth,
                              const float* g_semantics, const float* direction_norms, float loss_scale, void* dh,
                              void* dh_sem, float* d_w_sel, float* grad_w_color, float* grad_w_sem, void* stream) {
  UCSA_REQUIRE(sel && ray_off && rays_d && h && w_color_h && w_sem_h && rgb && hc1 && hc2 && hs && w_sel &&
                   z_sel && g_image && g_depth && g_semantics && direction_norms && dh && d_w_sel && grad_w_color &&
                   grad_w_sem,
               "heads_bwd: null pointer");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= UCSA_MAX_CLASSES, "heads_bwd: 1 <= classes <= %d", UCSA_MAX_CLASSES);
  UCSA_REQUIRE(loss_scale > 0.f, "heads_bwd: loss_scale must be positive");
  if (k_max == 0) return UCSA_OK;
  {
    static std::atomic<uint64_t> smem_devices{0};  // per device: the attribute belongs to the context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(smem_devices.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_bwd_color_kernel), kBwdColorSmem, "heads_bwd_color_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_bwd_sem_kernel<0>), kBwdSemSmem, "heads_bwd_sem_kernel")) return rc;
      if (int rc = set_max_dyn_smem(reinterpret_cast<const void*>(heads_bwd_sem_kernel<40>), kBwdSemSmem, "heads_bwd_sem_kernel")) return rc;
      smem_devices.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
  }
  cudaStream_t st = as_stream(stream);
  // With dh_sem the two kernels are independent (each writes its own share of dL/dgeo_feat) and CAN run concurrently,
  // the semantic one on the helper stream (see ucsa_heads_fwd).  Measured (bench.py, config 2, ms for both kernels;
  // colour:semantic CTAs per SM): sequential 0.450 | 1:2 0.445 | 1:4 0.484 | 1:3 0.515 | 2:2 0.619 | 2:1 0.707 --
  // unlike the forward pair, both backward kernels are bound by the same thing (latency at low occupancy), so
  // sharing the SM buys nothing: the default stays sequential at full occupancy, UCSA_BWD_OVERLAP=1 opts in.
  static int color_ctas = 0, sem_ctas = 0, overlap = -1;
  if (overlap < 0) {
    std::lock_guard<std::mutex> lock(g_side_mutex);
    const char* e = getenv("UCSA_BWD_OVERLAP");
    const int ov = (e != nullptr && e[0] == '1') ? 1 : 0;
    color_ctas = ov ? 1 : kBwdColorCtas;
    sem_ctas = ov ? 2 : kBwdSemCtas;
    if (const char* c = getenv("UCSA_BWD_COLOR_CTAS")) color_ctas = atoi(c) >= 1 && atoi(c) <= kBwdColorCtas ? atoi(c) : color_ctas;
    if (const char* c = getenv("UCSA_BWD_SEM_CTAS")) sem_ctas = atoi(c) >= 1 && atoi(c) <= kBwdSemCtas ? atoi(c) : sem_ctas;
    overlap = ov;
  }
  const bool concurrent = overlap && dh_sem != nullptr;
  cudaStream_t st_sem = st;
  int dev = 0;
  if (concurrent) {
    if (int rc = fork_side_stream(st, &st_sem, &dev)) return rc;
  }
  heads_bwd_color_kernel<<<heads_grid(k_max, concurrent ? color_ctas : kBwdColorCtas), 128, kBwdColorSmem, st>>>(
      sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_color_h), rgb,
      static_cast<const __half*>(hc1), static_cast<const __half*>(hc2), w_sel, z_sel, g_image, g_depth,
      direction_norms, loss_scale, static_cast<__half*>(dh), d_w_sel, grad_w_color);
  auto sem_kernel = (n_classes == 40 && reinterpret_cast<uintptr_t>(g_semantics) % 16 == 0) ? heads_bwd_sem_kernel<40>
                                                                                          : heads_bwd_sem_kernel<0>;
  sem_kernel<<<heads_grid(k_max, concurrent ? sem_ctas : kBwdSemCtas), 128, kBwdSemSmem, st_sem>>>(
      sel, ray_off + n_rays, t, rays_d, static_cast<const __half*>(h), static_cast<const __half*>(w_sem_h),
      static_cast<int>(n_classes), static_cast<const __half*>(hs), w_sel,
      g_semantics, loss_scale, static_cast<__half*>(dh), static_cast<__half*>(dh_sem), grad_w_sem);
  if (concurrent) join_side_stream(st, dev);
  return check_launch("heads_bwd");
}
