// Probe: does an M = 64 tcgen05.mma accept an accumulator address with lane offset 16, so that two [64 x N] weight-
// gradient accumulators can share the same TMEM columns (lanes 0-15 / 16-31 of every 32-lane quadrant)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../ucsa_neural_rendering_b200/csrc \
//        tmem_lane_probe.cu -o ../_build/tmem_lane_probe
#include <cstdio>
#include <vector>

#include "mlp_umma.cuh"

namespace ucsa {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
int set_max_dyn_smem(const void*, size_t, const char*) { return 0; }
}  // namespace ucsa
using namespace ucsa;
using umma::Tile;

__global__ void __launch_bounds__(128) probe(float* out, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* t_p = smem;                        // [128 x 64] halves: P[row][m] = (m + 1) * 0.01 if row == m else 0
  unsigned char* t_q = t_p + Tile<64>::kBytes;      // [128 x 16] halves: Q[row][n] = n + 1 (rows < 64), 0 otherwise
  unsigned char* t_q2 = t_q + Tile<16>::kBytes;     // second product: Q2[row][n] = 100 + n
  uint64_t* bar = reinterpret_cast<uint64_t*>(t_q2 + Tile<16>::kBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int row = threadIdx.x;
  for (int c = 0; c < 8; ++c) {
    H8 v;
    for (int i = 0; i < 8; ++i) {
      const int m = c * 8 + i;
      v.h[i] = __float2half_rn(row == m ? 1.0f : 0.f);  // P = identity on the first 64 rows -> D = Q[0:64]
    }
    *Tile<64>::chunk(t_p, row, c) = v.v;
  }
  for (int c = 0; c < 2; ++c) {
    H8 v, w;
    for (int i = 0; i < 8; ++i) {
      v.h[i] = __float2half_rn(row < 64 ? static_cast<float>(row * 16 + c * 8 + i) : 0.f);
      w.h[i] = __float2half_rn(row < 64 ? static_cast<float>(1024 + row * 16 + c * 8 + i) : 0.f);
    }
    *Tile<16>::chunk(t_q, row, c) = v.v;
    *Tile<16>::chunk(t_q2, row, c) = w.v;
  }
  umma::Ctx ctx = umma::ctx_init(slot, bar, 32);
  // clear the 32 columns on all lanes with an M = 128 product of zeros?  simpler: two M = 64 products at lane offsets
  ctx.publish();
  if (threadIdx.x == 0) {
    umma::tc_fence_after();
    umma::issue_wgrad<16>(ctx.tmem, umma::smem_u32(t_p), umma::smem_u32(t_q), true);  // lanes 0-15 (+32 i)
    if (mode >= 1)
      umma::issue_wgrad<16>(ctx.tmem + (16u << 16), umma::smem_u32(t_p), umma::smem_u32(t_q2), true);  // lane offset 16
    umma::commit(ctx.bar);
  }
  ctx.wait();
  float v[16];
  umma::tmem_ld16(ctx.lane_addr(0), v);
  for (int i = 0; i < 16; ++i) out[row * 16 + i] = v[i];
  umma::ctx_free(ctx, 32);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * sizeof(float));
  const size_t smem = Tile<64>::kBytes + 2 * Tile<16>::kBytes + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(d_out, 0xff, 128 * 16 * sizeof(float));
    probe<<<1, 128, smem>>>(d_out, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> h(128 * 16);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    // expected: accumulator row m (0..63) = Q[m][:] = m*16 + n at TMEM lane m%16 + 32*(m/16); second product at +16
    int ok1 = 0, ok2 = 0;
    for (int m = 0; m < 64; ++m) {
      const int lane = m % 16 + 32 * (m / 16);
      bool a = true, b = true;
      for (int n = 0; n < 16; ++n) {
        a &= h[lane * 16 + n] == static_cast<float>(m * 16 + n);
        b &= h[(lane + 16) * 16 + n] == static_cast<float>(1024 + m * 16 + n);
      }
      ok1 += a;
      ok2 += b;
    }
    printf("  rows correct at lanes m%%16+32(m/16): %d / 64; second product at +16: %d / 64\n", ok1, ok2);
    printf("  lane 0: %.0f %.0f | lane 16: %.0f %.0f | lane 32: %.0f | lane 48: %.0f\n", h[0], h[1], h[16 * 16], h[16 * 16 + 1],
           h[32 * 16], h[48 * 16]);
  }
  return 0;
}
