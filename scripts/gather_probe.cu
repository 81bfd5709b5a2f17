// Bring-up probe (not part of the library): what bounds random 16-byte gathers from an L2-resident 25 MB table?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/gather_probe scripts/gather_probe.cu
// Variants: 0 = ld.global.nc.v4 (the kernel's path), 1 = ld.global.nc.L1::no_allocate.v4,
//           2 = cp.async.cg 16 B into shared memory (bypasses L1), 3 = ld.global.nc.b32 (4-byte gathers, 2 per pair)
// for `depth` loads in flight per thread and a given CTA count per SM.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE, int DEPTH>
__global__ void __launch_bounds__(128) gather_kernel(const uint4* __restrict__ table, uint32_t n_groups, int rounds,
                                                     uint32_t* __restrict__ out) {
  extern __shared__ uint4 stage[];  // MODE 2: [DEPTH][128]
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (int r = 0; r < rounds; ++r) {
    uint4 v[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const uint32_t g = mix(tid * 977u + r * 131071u + d * 7919u) % n_groups;
      const uint4* p = table + g;
      if (MODE == 0) {
        asm("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[d].x), "=r"(v[d].y), "=r"(v[d].z), "=r"(v[d].w) : "l"(p));
      } else if (MODE == 1) {
        asm("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
            : "=r"(v[d].x), "=r"(v[d].y), "=r"(v[d].z), "=r"(v[d].w) : "l"(p));
      } else if (MODE == 2) {
        const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(&stage[d * 128 + threadIdx.x]));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(p) : "memory");
      } else {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        asm("ld.global.nc.b32 %0, [%1];" : "=r"(v[d].x) : "l"(q + (g & 3)));
        asm("ld.global.nc.b32 %0, [%1];" : "=r"(v[d].y) : "l"(q + ((g >> 2) & 3)));
        v[d].z = v[d].w = 0;
      }
    }
    if (MODE == 2) {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) v[d] = stage[d * 128 + threadIdx.x];
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc += v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
  }
  out[tid] = acc;
}

// random reductions into a 52 MB fp32 table (the gradient scatter of the fine hashed levels)
template <int VEC>
__global__ void __launch_bounds__(128) red_kernel(float* __restrict__ table, uint32_t n_groups, int rounds) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int r = 0; r < rounds; ++r) {
    const uint32_t g = mix(tid * 977u + r * 131071u) % n_groups;
    float* p = table + 4ull * g;
    if (VEC == 4) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    else if (VEC == 2) asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p + 2 * (g & 1)), "f"(1.0f) : "memory");
    else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + (g & 3)), "f"(1.0f) : "memory");
  }
}

template <int VEC>
void run_red(float* table, uint32_t n_groups, int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm, rounds = 1024;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(a);
    red_kernel<VEC><<<grid, 128>>>(table, n_groups, rounds);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  const double ops = (double)grid * 128 * rounds;
  printf("red.global.add.v%d.f32 random   ctas/sm %2d : %7.3f ms  %7.1f G reductions/s  (%s)\n", VEC, ctas_per_sm, best,
         ops / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int DEPTH>
void run(const uint4* table, uint32_t n_groups, uint32_t* out, int ctas_per_sm, const char* name) {
  const int grid = 148 * ctas_per_sm, rounds = 256 / DEPTH * 4;
  const size_t smem = MODE == 2 ? sizeof(uint4) * DEPTH * 128 : 0;
  cudaFuncSetAttribute(gather_kernel<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(a);
    gather_kernel<MODE, DEPTH><<<grid, 128, smem>>>(table, n_groups, rounds, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  const double loads = (double)grid * 128 * rounds * DEPTH;
  printf("%-28s depth %2d  ctas/sm %d : %7.3f ms  %7.1f G gathers/s  (%s)\n", name, DEPTH, ctas_per_sm, best,
         loads / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const uint32_t n_groups = 6537456 / 4;  // 16-byte groups of the 25 MB fp16 table
  uint4* table;
  uint32_t* out;
  cudaMalloc(&table, sizeof(uint4) * n_groups);
  cudaMemset(table, 1, sizeof(uint4) * n_groups);
  cudaMalloc(&out, sizeof(uint32_t) * 148 * 16 * 128);
  for (int c : {4, 8, 16}) {
    run<0, 16>(table, n_groups, out, c, "ld.nc.v4");
    run<1, 16>(table, n_groups, out, c, "ld.nc.L1::no_allocate.v4");
    run<2, 16>(table, n_groups, out, c, "cp.async.cg 16B");
    run<3, 16>(table, n_groups, out, c, "2 x ld.nc.b32 (same group)");
  }
  run<0, 4>(table, n_groups, out, 4, "ld.nc.v4");
  run<0, 32>(table, n_groups, out, 4, "ld.nc.v4");
  run<2, 32>(table, n_groups, out, 4, "cp.async.cg 16B");
  float* grad;
  const uint32_t grad_groups = 13074912 / 4;  // 16-byte groups of the 52 MB fp32 gradient table
  cudaMalloc(&grad, sizeof(float) * 4ull * grad_groups);
  cudaMemset(grad, 0, sizeof(float) * 4ull * grad_groups);
  for (int c : {4, 16}) {
    run_red<4>(grad, grad_groups, c);
    run_red<2>(grad, grad_groups, c);
    run_red<1>(grad, grad_groups, c);
  }
  return 0;
}
