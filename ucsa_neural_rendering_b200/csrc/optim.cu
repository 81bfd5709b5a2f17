// Parameter plumbing: fp32 master -> fp16 working copy, and the fused Adam step of row f1 of SURVEY.md
// section 8 (joint_train_lightning_net.py:897-919: Adam, betas (0.9,0.99), eps 1e-15, weight decay on the
// MLP group; GradScaler unscale / inf-skip semantics of :46,509-513 folded into the same pass).
#include "common.cuh"

namespace ucsa {
namespace {

__global__ void cast_kernel(const float* __restrict__ src, uint64_t n, __half* __restrict__ dst) {
  const uint64_t i = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (uint64_t k = i; k < n; ++k) dst[k] = __float2half_rn(src[k]);
  }
}

// torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected moments), operation by operation
// as torch/optim/adam.py writes them: exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad,
// value = 1 - beta2); denom = exp_avg_sq.sqrt() / sqrt(bias_correction2) + eps; param.addcdiv_(exp_avg, denom,
// value = -lr / bias_correction1).  The scalars (1 - beta, bias corrections, step size) are formed in double like
// Python does and rounded once to float.
struct AdamCoef {
  float step_size, b1, one_minus_b1, b2, one_minus_b2, eps, wd, ginv, bc2_sqrt;
};

// the per-launch scalars, from the (device-side) step count; thread 0 of a CTA
__device__ __forceinline__ void adam_scalars(double lr, double b1, double b2, int step, float* out_step_size,
                                             float* out_bc2_sqrt) {
  const double bc1 = 1.0 - pow(b1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(b2, static_cast<double>(step));
  *out_step_size = static_cast<float>(lr / bc1);
  *out_bc2_sqrt = static_cast<float>(sqrt(bc2));
}

// Explicit roundings (no compiler-chosen contraction): both kernels below produce identical bits, and the sequence is
// the one torch's foreach kernels compile to (lerp -> fma(w, g - m, m); addcmul -> fma(value * g, g, v); addcdiv ->
// fma(value, m / denom, p); add(param, alpha = wd) -> fma(wd, p, g)).
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCoef& c) {
  float grad = __fmul_rn(g, c.ginv);
  if (c.wd != 0.f) grad = __fmaf_rn(c.wd, p, grad);
  m = __fmaf_rn(c.one_minus_b1, __fsub_rn(grad, m), m);
  v = __fmul_rn(v, c.b2);
  v = __fmaf_rn(__fmul_rn(c.one_minus_b2, grad), grad, v);
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
  p = __fmaf_rn(-c.step_size, __fdiv_rn(m, denom), p);
}

// Four parameters per thread (16-byte accesses); the bias corrections are formed once per CTA, because with the
// step count on the device (CUDA-graph replay) they cost two pow per evaluation.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __half* __restrict__ p_h, uint64_t n, double lr, double b1, double b2, float eps, float wd,
            uint64_t wd_begin, float ginv, const float* __restrict__ grad_scale_dev,
            const float* __restrict__ found_inf, int step, const int32_t* __restrict__ step_dev,
            const int32_t* __restrict__ skipped_dev) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: skip the step on overflow
  __shared__ float bc[3];
  if (threadIdx.x == 0) {
    // step count on the device (CUDA-graph replay); skipped steps do not advance it (GradScaler.step + Adam)
    if (step_dev != nullptr) step = *step_dev - (skipped_dev != nullptr ? *skipped_dev : 0);
    adam_scalars(lr, b1, b2, step, &bc[0], &bc[1]);
    // GradScaler hands its scale over as a device tensor: unscale by its reciprocal, formed like torch does (in double)
    bc[2] = grad_scale_dev != nullptr ? ginv * static_cast<float>(1.0 / static_cast<double>(*grad_scale_dev)) : ginv;
  }
  __syncthreads();
  AdamCoef c{bc[0], static_cast<float>(b1), static_cast<float>(1.0 - b1), static_cast<float>(b2),
             static_cast<float>(1.0 - b2), eps, wd, bc[2], bc[1]};
  // grid-stride over float4 groups: a bounded number of CTAs, so the double-precision scalars above are formed a few
  // thousand times per launch instead of once per 1024 parameters
  const uint64_t n4 = n / 4;
  for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < n4;
       q += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t i = q * 4;
    c.wd = i >= wd_begin ? wd : 0.f;  // weight decay applies from parameter wd_begin on (a multiple of 4)
    float4 pp = *reinterpret_cast<const float4*>(p + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, c);
    adam_one(pp.y, gg.y, mm.y, vv.y, c);
    adam_one(pp.z, gg.z, mm.z, vv.z, c);
    adam_one(pp.w, gg.w, mm.w, vv.w, c);
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (p_h != nullptr) {
      const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p_h + i) = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) {  // scalar tail
    const uint64_t k = n4 * 4 + threadIdx.x;
    c.wd = k >= wd_begin ? wd : 0.f;
    float pk = p[k], mk = m[k], vk = v[k];
    adam_one(pk, g[k], mk, vk, c);
    p[k] = pk;
    m[k] = mk;
    v[k] = vk;
    if (p_h != nullptr) p_h[k] = __float2half_rn(pk);
  }
}


// GradScaler's inf check (joint_train_lightning_net.py:46,509-513 -> torch.amp.GradScaler.unscale_/step): one pass
// over the flat gradient buffer; found_inf = 1 when any entry is inf / NaN.  The last CTA to finish publishes the
// flag, bumps the skipped-step counter and re-arms the scratch words, so the launch is graph-replayable.
__global__ void __launch_bounds__(256)
grad_check_kernel(const float* __restrict__ g, uint64_t n, float* __restrict__ found_inf,
                  int32_t* __restrict__ skipped_dev, uint32_t* __restrict__ scratch) {
  const uint64_t n4 = n / 4;
  bool bad = false;
  const uint4* g4 = reinterpret_cast<const uint4*>(g);
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint4 q = g4[i];
    // exponent all ones <=> inf or NaN
    bad |= ((q.x & 0x7f800000u) == 0x7f800000u) | ((q.y & 0x7f800000u) == 0x7f800000u) |
           ((q.z & 0x7f800000u) == 0x7f800000u) | ((q.w & 0x7f800000u) == 0x7f800000u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u))
    bad |= (__float_as_uint(g[n4 * 4 + threadIdx.x]) & 0x7f800000u) == 0x7f800000u;
  const int any = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) {
    if (any) atomicOr(&scratch[1], 1u);
    __threadfence();
    const uint32_t ticket = atomicAdd(&scratch[0], 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const uint32_t flag = atomicExch(&scratch[1], 0u);
      *found_inf = flag ? 1.0f : 0.0f;
      if (flag && skipped_dev != nullptr) *skipped_dev += 1;
      scratch[0] = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient exchange fused into the optimizer, over NVLink / NVSwitch peer memory (row (e) of SURVEY.md section 8).
// Every rank holds the full parameters (fp32 masters + fp16 working copy) and a full gradient buffer in SYMMETRIC
// memory.  Rank r owns the slice [begin, end) of the flat parameter space:
//     g      = sum over ranks of grad[i]            one multimem.ld_reduce (the switch adds) or P2P loads
//     p, m, v -> Adam                               moments exist only for the owned slice
//     p'     -> every rank's masters and fp16 copy  one multimem.st each (the switch replicates) or P2P stores
// i.e. reduce-scatter + optimizer + all-gather in one pass, 1/world of the Adam work per rank, and every copy of a
// parameter comes from the same arithmetic (bit-identical replicas).  The caller brackets the launch with two
// cross-rank barriers (gradients complete before, parameters visible after).
// Measured on B200 (bench.py, 13.1 M parameters): kernel 0.108 ms with peer loads / stores vs 0.187 ms with multimem
// at 2 GPUs, 0.205 ms vs 0.181 ms at 8 GPUs (step 3.26 -> 3.18 ms; NCCL all-reduce + replicated Adam: 3.27 ms):
// the multimem operations cost more per 16 bytes but their traffic does not grow with the world size.
constexpr int kMaxPeers = UCSA_MAX_PEERS;
struct PeerPtrs {
  float* grad[kMaxPeers];
  float* param[kMaxPeers];
  __half* param_h[kMaxPeers];
  const float* found_inf[kMaxPeers];  // per-rank overflow flags (ucsa_grad_check), or all null
};

__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f32x4(float* mc, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st_f16x4(__half* mc, uint32_t lo, uint32_t hi) {
  asm volatile("multimem.st.weak.global.v2.f16x2 [%0], {%1,%2};" ::"l"(mc), "r"(lo), "r"(hi) : "memory");
}

template <bool MC_LOAD, bool MC_STORE>
__global__ void __launch_bounds__(256)
adam_exchange_kernel(const PeerPtrs peers, const float* __restrict__ mc_grad, float* __restrict__ mc_param,
                     __half* __restrict__ mc_param_h, uint32_t world, uint32_t rank, uint64_t begin, uint64_t end,
                     uint64_t wd_begin, float* __restrict__ m, float* __restrict__ v, double lr, double b1, double b2,
                     float eps, float wd, int step, const int32_t* __restrict__ step_dev,
                     int32_t* __restrict__ skipped_dev, int broadcast_masters) {
  // GradScaler semantics across the job: an overflow on ANY rank skips the step on every rank (each rank reads all
  // the flags, so no extra collective is needed); skipped steps do not advance Adam's step count.
  if (peers.found_inf[0] != nullptr) {
    bool any = false;
    for (uint32_t r = 0; r < world; ++r) any |= *peers.found_inf[r] != 0.f;
    if (any) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && skipped_dev != nullptr) *skipped_dev += 1;
      return;
    }
  }
  __shared__ float bc[2];
  if (threadIdx.x == 0) {
    if (step_dev != nullptr) step = *step_dev - (skipped_dev != nullptr ? *skipped_dev : 0);
    adam_scalars(lr, b1, b2, step, &bc[0], &bc[1]);
  }
  __syncthreads();
  // kUnroll float4 groups per thread and iteration, all remote gradient loads issued before the first use: the
  // loads cross NVLink (microseconds of latency), so the bytes in flight per SM decide the throughput
  constexpr int kUnroll = 4;
  const uint64_t tile = static_cast<uint64_t>(blockDim.x) * 4 * kUnroll;  // parameters per CTA and iteration
  for (uint64_t base = begin + static_cast<uint64_t>(blockIdx.x) * tile; base < end;
       base += static_cast<uint64_t>(gridDim.x) * tile) {
    float4 g[kUnroll];
    uint64_t idx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      idx[u] = base + (static_cast<uint64_t>(u) * blockDim.x + threadIdx.x) * 4;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (MC_LOAD) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (idx[u] < end) g[u] = mc_ld_reduce_add(mc_grad + idx[u]);
    } else {
      for (uint32_t r = 0; r < world; ++r) {  // fixed order: the sum does not depend on the owner
        float4 t[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          t[u] = idx[u] < end ? *reinterpret_cast<const float4*>(peers.grad[r] + idx[u])
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) g[u].x += t[u].x, g[u].y += t[u].y, g[u].z += t[u].z, g[u].w += t[u].w;
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t i = idx[u];
      if (i >= end) continue;
      const AdamCoef c{bc[0], static_cast<float>(b1), static_cast<float>(1.0 - b1), static_cast<float>(b2),
                       static_cast<float>(1.0 - b2), eps, i >= wd_begin ? wd : 0.f, 1.0f, bc[1]};
      float4 pp = *reinterpret_cast<const float4*>(peers.param[rank] + i);
      const uint64_t k = i - begin;
      float4 mm = *reinterpret_cast<const float4*>(m + k);
      float4 vv = *reinterpret_cast<const float4*>(v + k);
      adam_one(pp.x, g[u].x, mm.x, vv.x, c);
      adam_one(pp.y, g[u].y, mm.y, vv.y, c);
      adam_one(pp.z, g[u].z, mm.z, vv.z, c);
      adam_one(pp.w, g[u].w, mm.w, vv.w, c);
      *reinterpret_cast<float4*>(m + k) = mm;
      *reinterpret_cast<float4*>(v + k) = vv;
      const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
      const uint32_t lo_b = *reinterpret_cast<const uint32_t*>(&lo), hi_b = *reinterpret_cast<const uint32_t*>(&hi);
      // broadcast_masters == 0: the fp32 masters of a slice live on its owner only (one third of the store traffic);
      // every rank still receives the fp16 working copy the kernels read
      if (MC_STORE) {
        if (broadcast_masters) mc_st_f32x4(mc_param + i, pp);
        else *reinterpret_cast<float4*>(peers.param[rank] + i) = pp;
        mc_st_f16x4(mc_param_h + i, lo_b, hi_b);
      } else {
        if (!broadcast_masters) *reinterpret_cast<float4*>(peers.param[rank] + i) = pp;
        for (uint32_t r = 0; r < world; ++r) {
          if (broadcast_masters) *reinterpret_cast<float4*>(peers.param[r] + i) = pp;
          *reinterpret_cast<uint2*>(peers.param_h[r] + i) = make_uint2(lo_b, hi_b);
        }
      }
    }
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_cast_f32_to_f16(const float* src, uint64_t n, void* dst_h, void* stream) {
  UCSA_REQUIRE(src && dst_h, "cast_f32_to_f16: null pointer");
  UCSA_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst_h) % 8 == 0,
               "cast_f32_to_f16: buffers must be 16-byte (src) / 8-byte (dst) aligned");
  if (n == 0) return UCSA_OK;
  cast_kernel<<<ceil_div((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(src, n, static_cast<__half*>(dst_h));
  return check_launch("cast_f32_to_f16");
}

extern "C" int ucsa_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_h,
                              uint64_t n, double lr, double beta1, double beta2, float eps, float weight_decay,
                              uint64_t wd_begin, float grad_scale_inv, const float* grad_scale_dev,
                              const float* found_inf, uint32_t step, const int32_t* step_dev,
                              const int32_t* skipped_dev, void* stream) {
  UCSA_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
  UCSA_REQUIRE(step >= 1 || step_dev != nullptr, "adam_step: step counts from 1");
  UCSA_REQUIRE(wd_begin % 4 == 0, "adam_step: wd_begin must be a multiple of 4 parameters");
  UCSA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                 reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(param_h) & 7u) == 0,
               "adam_step: buffers must be 16-byte aligned (fp16 copy: 8-byte)");
  if (n == 0) return UCSA_OK;
  uint32_t blocks = ceil_div((n + 3) / 4, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq,
                                                               static_cast<__half*>(param_h), n, lr, beta1, beta2,
                                                               eps, weight_decay, wd_begin, grad_scale_inv, grad_scale_dev,
                                                               found_inf, static_cast<int>(step), step_dev, skipped_dev);
  return check_launch("adam_step");
}

extern "C" int ucsa_grad_check(const float* grad, uint64_t n, float* found_inf, int32_t* skipped_dev,
                               uint32_t* scratch2, void* stream) {
  UCSA_REQUIRE(grad && found_inf && scratch2, "grad_check: null pointer");
  UCSA_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15u) == 0, "grad_check: grad must be 16-byte aligned");
  const uint64_t n4 = n / 4;
  uint32_t blocks = ceil_div(n4 > 0 ? n4 : 1, 256 * 8);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  grad_check_kernel<<<blocks, 256, 0, as_stream(stream)>>>(grad, n, found_inf, skipped_dev, scratch2);
  return check_launch("grad_check");
}

extern "C" int ucsa_adam_exchange(const uint64_t* grad_ptrs_host, const uint64_t* param_ptrs_host,
                                  const uint64_t* param_h_ptrs_host, const float* mc_grad, float* mc_param,
                                  void* mc_param_h, uint32_t world, uint32_t rank, uint64_t begin, uint64_t end,
                                  uint64_t wd_begin, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                                  double beta2, float eps, float weight_decay, uint32_t step, const int32_t* step_dev,
                                  const uint64_t* found_inf_ptrs_host, int32_t* skipped_dev, int broadcast_masters,
                                  void* stream) {
  UCSA_REQUIRE(grad_ptrs_host && param_ptrs_host && param_h_ptrs_host && exp_avg && exp_avg_sq,
               "adam_exchange: null pointer");
  UCSA_REQUIRE(world >= 1 && world <= UCSA_MAX_PEERS && rank < world, "adam_exchange: 1 <= world <= %d, rank < world",
               UCSA_MAX_PEERS);
  UCSA_REQUIRE(begin % 4 == 0 && end % 4 == 0 && begin <= end && wd_begin % 4 == 0,
               "adam_exchange: slice bounds must be multiples of 4 parameters");
  UCSA_REQUIRE(step >= 1 || step_dev != nullptr, "adam_exchange: step counts from 1");
  const bool multicast = mc_grad != nullptr;
  UCSA_REQUIRE(!multicast || (mc_param_h && (mc_param || !broadcast_masters)),
               "adam_exchange: give the multicast addresses of every buffer that is broadcast, or none");
  if (begin == end) return UCSA_OK;
  PeerPtrs peers{};
  for (uint32_t r = 0; r < world; ++r) {
    peers.grad[r] = reinterpret_cast<float*>(grad_ptrs_host[r]);
    peers.param[r] = reinterpret_cast<float*>(param_ptrs_host[r]);
    peers.param_h[r] = reinterpret_cast<__half*>(param_h_ptrs_host[r]);
    peers.found_inf[r] = found_inf_ptrs_host ? reinterpret_cast<const float*>(found_inf_ptrs_host[r]) : nullptr;
    UCSA_REQUIRE(peers.grad[r] && peers.param_h[r] && (peers.param[r] || (!broadcast_masters && r != rank)),
                 "adam_exchange: null peer pointer");
    UCSA_REQUIRE(!found_inf_ptrs_host || peers.found_inf[r], "adam_exchange: null found_inf pointer");
  }
  const uint64_t tiles = ((end - begin) + 256 * 4 * 4 - 1) / (256 * 4 * 4);  // 256 threads x float4 x kUnroll
  const uint32_t cap = kNumSMs * 8;
  const uint32_t blocks = static_cast<uint32_t>(tiles < cap ? tiles : cap);
  auto kernel = multicast ? adam_exchange_kernel<true, true> : adam_exchange_kernel<false, false>;
  kernel<<<blocks, 256, 0, as_stream(stream)>>>(peers, mc_grad, mc_param, static_cast<__half*>(mc_param_h), world, rank,
                                               begin, end, wd_begin, exp_avg, exp_avg_sq, lr, beta1, beta2, eps,
                                               weight_decay, static_cast<int>(step), step_dev, skipped_dev,
                                               broadcast_masters);
  return check_launch("adam_exchange");
}
