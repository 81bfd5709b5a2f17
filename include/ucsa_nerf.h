/*
 * ucsa_nerf.h -- C ABI of libucsa_nerf.so, the B200-native (sm_100a) Semantic-NeRF hot path.
 *
 * Drop-in boundary for the native layer of ethz-asl/ucsa_neural_rendering (nr4seg/nerf):
 *   - replaces the pybind11 module `_raymarching`
 *       (nr4seg/nerf/raymarching/src/bindings.cpp:5-18, raymarching.h:7-18), and
 *   - replaces the tiny-cuda-nn modules the network calls
 *       (nr4seg/nerf/network_tcnn_semantics.py:36-100: HashGrid, SphericalHarmonics, FullyFusedMLP), and
 *   - provides the fused kernels behind SemanticNeRFRenderer.run()
 *       (nr4seg/nerf/renderer_semantics.py:123-299) that the reference executes as eager torch ops.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all memory; nothing is allocated, freed or synchronised inside;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and is CUDA-graph capturable;
 *   - return value 0 = success, negative = error (UCSA_ERR_*); ucsa_last_error_string() describes it;
 *   - fp16 buffers are IEEE binary16 (`__half`), passed as void*;
 *   - "cat order" of a ray's samples: slot k in [0,Tc) is coarse sample k, slot Tc+j is fine sample j
 *     (the order of torch.cat at renderer_semantics.py:221); "sorted order" is after the sort at :222.
 */
#ifndef UCSA_NERF_H
#define UCSA_NERF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(UCSA_BUILDING_LIBRARY)
#define UCSA_API __attribute__((visibility("default")))
#else
#define UCSA_API
#endif

#define UCSA_OK 0
#define UCSA_ERR_INVALID_ARGUMENT (-1)
#define UCSA_ERR_CUDA (-2)
#define UCSA_ERR_UNSUPPORTED (-3)

#define UCSA_ABI_VERSION 1
#define UCSA_GRID_LEVELS 16
#define UCSA_SIGMA_PARAMS 3072  /* 32->64->16            network_tcnn_semantics.py:48-58  */
#define UCSA_COLOR_PARAMS 7168  /* 32->64->64->16        network_tcnn_semantics.py:74-84  */
#define UCSA_MAX_CLASSES 48     /* semantics 16->64->pad16(C)  network_tcnn_semantics.py:90-100 */
#define UCSA_LOSS_SCRATCH_BYTES 2048 /* ucsa_nerf_loss: per-CTA partial sums + done counter */
#define UCSA_MAX_PEERS 16       /* ranks of one NVLink domain in ucsa_adam_exchange */
#define UCSA_TILE_ROWS(rows) (((rows) + 127u) / 128u * 128u) /* rows of a tile-layout activation buffer */

/* Multiresolution hash grid geometry (tcnn HashGrid config at network_tcnn_semantics.py:36-46).
 * Filled by ucsa_grid_desc_init() on the host and passed BY VALUE to kernels. */
typedef struct ucsa_grid_desc {
  float scale[UCSA_GRID_LEVELS];     /* exp2f(l*log2f(pls))*base - 1                     */
  uint32_t res[UCSA_GRID_LEVELS];    /* ceil(scale)+1                                    */
  uint32_t entries[UCSA_GRID_LEVELS];/* min(round_up(res^3,8), 2^log2_hashmap)           */
  uint32_t offset[UCSA_GRID_LEVELS]; /* first entry of the level (in entries of 2 feats) */
  uint32_t hashed[UCSA_GRID_LEVELS]; /* 1 = prime-xor hash, 0 = dense x+y*res+z*res^2    */
  uint32_t total_entries;            /* sum(entries); the table holds 2*total fp16/fp32   */
} ucsa_grid_desc;

UCSA_API int ucsa_abi_version(void);
UCSA_API const char* ucsa_last_error_string(void);

/* Host helper: geometry for `bound` (per_level_scale = 2^(log2(2048*bound/16)/15), network_tcnn_semantics.py:34). */
UCSA_API int ucsa_grid_desc_init(float per_level_scale, uint32_t base_resolution, uint32_t log2_hashmap_size,
                        ucsa_grid_desc* out_host);

/* ---- a2. near/far.  Replaces near_far_from_aabb (raymarching.h:7, raymarching.cu:62-126). */
UCSA_API int ucsa_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb6, uint32_t n_rays,
                            float min_near, float* nears, float* fars, void* stream);

/* ---- a3. coarse sampling (renderer_semantics.py:154-168).  Writes slots [0,Tc) of z_cat [N,T].
 * lin[Tc] is torch.linspace(0,1,Tc).  perturb: stratified jitter; the uniform numbers come from t_rand [N,Tc]
 * when non-null, else from the counter-based generator keyed by (seed, ray_base+n, k).  step_dev (nullable, device
 * int32) is mixed into the seed on the device, so a captured CUDA graph draws fresh numbers on every replay. */
UCSA_API int ucsa_sample_coarse(const float* nears, const float* fars, const float* lin, const float* t_rand,
                       uint64_t seed, const int32_t* step_dev, uint32_t ray_base, int perturb, uint32_t n_rays,
                       uint32_t tc, uint32_t t, float* z_cat, void* stream);

/* ---- a4/a5/a6/a8. density = hash-grid encode + sigma MLP + trunc_exp (network_tcnn_semantics.py:130-144).
 * Sample positions are either xyz [S,3] (then rays_*, z_cat are ignored and a "ray" has T=1 slot), or
 * o + d*z clipped to aabb6 (renderer_semantics.py:171-173) for slots [k0,k1) of every ray.
 * table_h: fp16 [2*total_entries]; w_sigma_h: fp16 [3072], layer-major, each layer [out][in] row-major.
 * Outputs, indexed n*T+k: sigma f32; h fp16 [.,16] (h[0] = log-density, h[1:16] = geo_feat);
 * enc fp16 [.,32] and hid fp16 [.,64] are saved for the backward pass when non-null.
 * tiled != 0: enc / hid are OPAQUE tile-layout buffers of UCSA_TILE_ROWS(rows) rows (128-sample tiles stored as
 * contiguous blocks in the tensor-core operand layout and moved with one bulk copy each); in ray mode this needs
 * T, k0 and k1-k0 to be multiples of 128.  tiled == 0: plain row-major rows.  Forward and backward must agree. */
UCSA_API int ucsa_density_fwd(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                     const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
                     const void* table_h, const ucsa_grid_desc* grid_host, const void* w_sigma_h,
                     float* sigma, void* h, void* enc, void* hid, int tiled, void* stream);

/* Backward of ucsa_density_fwd for slots [k0,k1).  d_sigma f32 [N,T] (cat order); dh fp16 [N,T,16] whose
 * elements 1..15 hold loss_scale * dL/dgeo_feat for samples with use_geo[n*T+k] != 0 (element 0 is ignored);
 * trunc_exp backward (activation.py:16-19) is applied here.  Accumulates (atomicAdd, fp32) into
 * grad_table [2*total] and grad_w_sigma [3072]; both already divided by loss_scale.
 * grad_replicas (nullable) = n_replicas zero-filled private copies [R][2*dense_entries] of the dense levels' part of
 * the table: CTA b adds the dense levels into copy b % R (all rays of a batch leave one camera, so a few coarse cells
 * take most of the adds and serialise the L2 atomic units); fold them with ucsa_reduce_grad_replicas afterwards. */
UCSA_API int ucsa_density_bwd(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                     const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
                     const ucsa_grid_desc* grid_host, const void* w_sigma_h, const void* h, const void* enc,
                     const void* hid, int tiled, const float* d_sigma, const void* dh, const void* dh2,
                     const uint8_t* use_geo, float loss_scale, float* grad_table, float* grad_replicas,
                     uint32_t n_replicas, float* grad_w_sigma, void* stream);
/* grad_table[dense part] += sum of the replicas; the replicas are zero again on return. */
UCSA_API int ucsa_reduce_grad_replicas(float* grad_replicas, uint32_t n_replicas, const ucsa_grid_desc* grid_host,
                              float* grad_table, void* stream);

/* ---- a9/a10. importance resampling + merge (renderer_semantics.py:182-222, sample_pdf :10-46).
 * Reads coarse z / sigma (slots [0,Tc)), writes the fine z into slots [Tc,Tc+Tf) IN ASCENDING ORDER (the same
 * set of samples as sample_pdf draws; only their order inside the cat buffer differs from torch.cat, which nothing
 * downstream observes) and order [N,T] (sorted position -> cat slot, the z_index of :222).
 * u [N,Tf] when non-null, else generated (seed). */
UCSA_API int ucsa_resample_merge(const float* sigma, float* z_cat, const float* u, uint64_t seed,
                        const int32_t* step_dev, uint32_t ray_base, uint32_t n_rays, uint32_t tc, uint32_t tf,
                        float density_scale, int32_t* order, void* stream);

/* ---- a11. weights, masks, depth (renderer_semantics.py:238-250,270-277).  order may be null (identity).
 * Writes w_sorted [N,T] (un-masked weights), depth [N] (sum of masked w*z / direction_norm),
 * ray_count [N] (#samples with w > 1e-4) and use_geo [N,T] (cat order, 1 where masked-in). */
UCSA_API int ucsa_weights_fwd(const float* z_cat, const float* sigma, const int32_t* order, const float* direction_norms,
                     uint32_t n_rays, uint32_t t, float density_scale, float* w_sorted, float* depth,
                     int32_t* ray_count, uint8_t* use_geo, void* stream);

/* exclusive scan of ray_count -> ray_off [N+1] (ray_off[N] = K, the number of masked-in samples). */
UCSA_API int ucsa_scan_counts(const int32_t* ray_count, uint32_t n_rays, int32_t* ray_off, void* stream);

/* compaction: sel [K] = n*T + cat slot, w_sel [K], z_sel [K] in (ray, sorted position) order. */
UCSA_API int ucsa_compact_masked(const float* w_sorted, const float* z_cat, const int32_t* order, const int32_t* ray_off,
                        uint32_t n_rays, uint32_t t, int32_t* sel, float* w_sel, float* z_sel, void* stream);

/* The three calls above as ONE launch (what the render pipeline uses): weights, masks and depth per ray, the global
 * offsets by a decoupled look-back scan over the CTAs, and the compaction written from the weights still held in shared
 * memory.  Same outputs bit for bit (w_sorted, depth, use_geo, ray_off [N+1], sel / w_sel / z_sel [K]).
 * scratch: UCSA_WEIGHTS_SCRATCH_WORDS(n_rays) 32-bit words, 8-byte aligned, ZERO before the first call; the kernel
 * leaves it zeroed, so one block serves every later call on the same stream (one block per concurrent stream). */
#define UCSA_WEIGHTS_SCRATCH_WORDS(n_rays) (2u + 2u * (((n_rays) + 3u) / 4u))
UCSA_API int ucsa_weights_compact(const float* z_cat, const float* sigma, const int32_t* order,
                         const float* direction_norms, uint32_t n_rays, uint32_t t, float density_scale,
                         float* w_sorted, float* depth, int32_t* ray_off, uint8_t* use_geo, int32_t* sel, float* w_sel,
                         float* z_sel, uint32_t* scratch, void* stream);

/* ---- a7/a12/a13 (+a14 fused). colour + semantic heads on the K masked-in rows (network_tcnn_semantics.py:147-207)
 * (two kernels: colour, then semantics) and, when image/semantics are non-null, the compositing of
 * renderer_semantics.py:279-285 fused into them:
 * image [N,3] += sum w*rgb, semantics [N,C] += sum w*softmax(logits) (atomicAdd; the caller zero-fills them).
 * K is read on the device from ray_off[n_rays]; k_max bounds the launch.  rgb [K,3] f32 (fp16 sigmoid values);
 * logits fp16 [K,48] row-major, optional (may be null: the backward pass recomputes them).
 * hc1, hc2, hs (all three or none): hidden activations saved for ucsa_heads_bwd, OPAQUE tile-layout buffers of
 * UCSA_TILE_ROWS(k_max) * 64 halves each - 128-row tiles stored as contiguous 16 KB blocks in the tensor-core
 * operand layout (csrc/mlp_umma.cuh), moved with one bulk copy per tile.
 * Two kernels (colour, semantics) that run concurrently: the second is launched on a library-owned helper stream
 * forked from and joined back to `stream` with events, so for the caller everything is ordered on `stream` and the
 * call can be captured into a CUDA graph (make the first call on a device outside a capture: it creates the helper
 * stream). */
UCSA_API int ucsa_heads_fwd(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t, uint32_t k_max,
                   const float* rays_d, const void* h, const void* w_color_h, const void* w_sem_h,
                   uint32_t n_classes, const float* w_sel, float* rgb, void* logits, void* hc1, void* hc2, void* hs,
                   float* image, float* semantics, void* stream);

/* Backward of ucsa_heads_fwd including the compositing: from g_image [N,3], g_depth [N], g_semantics [N,C] it forms
 * dL/drgb, dL/dlogits (soft-max backward; semantic weights are detached, renderer_semantics.py:270) per row on the
 * fly, writes d_w_sel [K] = dL/dw of every masked-in sample (for ucsa_weights_bwd), dh[sel][1..15] =
 * loss_scale * dL/dgeo_feat (fp16), and accumulates grad_w_color [7168], grad_w_sem [4096] (fp32, unscaled).
 * hc1, hc2, hs are the tile-layout buffers ucsa_heads_fwd filled.  Two kernels, colour and semantics.  dh_sem
 * (optional, same shape as dh): the semantic kernel writes its share of dL/dgeo_feat there instead of adding it
 * to dh, which makes the two kernels independent -- they then run concurrently (helper stream, see ucsa_heads_fwd) --
 * and ucsa_density_bwd takes the second share as `dh2`.  With dh_sem = null the kernels run one after the other
 * and dh holds the sum.  Both forms give bit-identical gradients. */
UCSA_API int ucsa_heads_bwd(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t, uint32_t k_max,
                   const float* rays_d, const void* h, const void* w_color_h, const void* w_sem_h,
                   uint32_t n_classes, const float* rgb, const void* hc1, const void* hc2,
                   const void* hs, const float* w_sel, const float* z_sel, const float* g_image,
                   const float* g_depth, const float* g_semantics, const float* direction_norms, float loss_scale,
                   void* dh, void* dh_sem, float* d_w_sel, float* grad_w_color, float* grad_w_sem, void* stream);

/* ---- a14, stand-alone form over the compact rows (renderer_semantics.py:279-285): image = sum w*rgb,
 * semantics = sum w*softmax(logits).  One pass, fp32.  The rendering pipeline uses the fused form inside
 * ucsa_heads_fwd/bwd; these entry points serve callers that hold rgb / logits already (and the cross-check tests). */
UCSA_API int ucsa_composite_fwd(const int32_t* ray_off, const float* w_sel, const float* rgb, const void* logits,
                       uint32_t n_rays, uint32_t n_classes, float* image, float* semantics, void* stream);

/* Backward of composite+weights: from g_image [N,3], g_depth [N], g_semantics [N,C] produce
 * d_rgb [K,3], d_logits f32 [K,48] (softmax backward; semantic weights are detached, :270),
 * and d_sigma [N,T] in cat order (through weights, masked like :271). */
UCSA_API int ucsa_composite_bwd(const int32_t* ray_off, const int32_t* sel, const float* w_sel, const float* z_sel,
                       const float* rgb, const void* logits, const float* g_image, const float* g_depth,
                       const float* g_semantics, const float* direction_norms, uint32_t n_rays,
                       uint32_t n_classes, float* d_rgb, float* d_logits, float* d_w_sel, void* stream);
UCSA_API int ucsa_weights_bwd(const float* z_cat, const float* sigma, const int32_t* order, const float* w_sorted,
                     const int32_t* ray_off, const float* d_w_sel, uint32_t n_rays, uint32_t t,
                     float density_scale, float* d_sigma, void* stream);

/* ---- a14 (dense form, BASELINE.json configs[0]).  Everything per sample is given:
 * sigma [N,T], z [N,T] (sorted), rgb [N,T,3], prob [N,T,C] f32.  Computes weights, the w>1e-4 masks and
 * the three composites exactly like renderer_semantics.py:238-285. */
UCSA_API int ucsa_composite_dense_fwd(const float* sigma, const float* z, const float* rgb, const float* prob,
                             const float* direction_norms, uint32_t n_rays, uint32_t t, uint32_t n_classes,
                             float density_scale, float* weights, float* depth, float* image, float* semantics,
                             void* stream);
UCSA_API int ucsa_composite_dense_bwd(const float* sigma, const float* z, const float* rgb, const float* weights,
                             const float* direction_norms, const float* g_depth, const float* g_image,
                             const float* g_semantics, uint32_t n_rays, uint32_t t, uint32_t n_classes,
                             float density_scale, float* d_sigma, float* d_rgb, float* d_prob, void* stream);

/* ---- a16-a20. occupancy-grid path: replaces the bound kernels of raymarching.h:9-18 / bindings.cpp:5-18.
 * The occupancy test reads `bitfield` when non-null (ucsa_grid_packbits), else the float `grid` [C,H,H,H] exactly like
 * raymarching.cu:157,204-210 (density > min(0.01, mean_density)).  Sample offsets are handed out in ray order:
 * rays[n] = (n, offset, count); counter[0] += total samples, counter[1] += n_rays.  `scratch` is int32
 * [UCSA_MARCH_SCRATCH_INTS(n_rays)] (counts, offsets and the block sums of the two-level scan used above 4096 rays).
 * t_stage (optional, float [n_rays * 1024]): staging for the parameter t of every occupied step; with it every ray is
 * marched once and the samples are written by a sample-parallel pass, without it every ray is marched twice (count,
 * then write).  Same samples either way. */
#define UCSA_MARCH_SCRATCH_INTS(n_rays) (2ull * (n_rays) + 2ull * (((n_rays) + 1023ull) / 1024ull) + 2ull)
UCSA_API int ucsa_march_rays_train(const float* rays_o, const float* rays_d, const float* grid, const uint32_t* bitfield,
                          float mean_density, float bound, float dt_gamma, uint32_t n_rays, uint32_t C, uint32_t H,
                          uint32_t max_points, const float* nears, const float* fars, float* xyzs, float* dirs,
                          float* deltas, int32_t* rays, int32_t* counter, uint32_t perturb, int32_t* scratch,
                          float* t_stage, void* stream);
UCSA_API int ucsa_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t C, uint32_t H,
                    const float* grid, const uint32_t* bitfield, float mean_density, const float* nears,
                    const float* fars, float* xyzs, float* dirs, float* deltas, uint32_t perturb, void* stream);
/* rgb + depth as raymarching.cu:318-487; local_semantics [M,C] / semantics [N,C] optional (both null or both set):
 * the semantic kernels the reference declares (raymarching.h:12-13) but never implemented.  Instead of the fp32
 * probabilities the forward pass also takes the fp16 logits of the semantic head (logits_h [M, logits_ld], soft-max
 * inside the kernel, as ucsa_composite_rays does); the backward pass hands out dL/d(probabilities) either way.
 * deltas [M,2] must be 8-byte aligned. */
UCSA_API int ucsa_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* local_semantics,
                                      const void* logits_h, uint32_t logits_ld, const float* deltas,
                                      const int32_t* rays, uint32_t M, uint32_t N, uint32_t n_classes,
                                      float* weights_sum, float* depth, float* image, float* semantics,
                                      void* stream);
UCSA_API int ucsa_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                       const float* grad_semantics, const float* sigmas, const float* rgbs,
                                       const float* deltas, const int32_t* rays, const float* weights_sum,
                                       const float* image, uint32_t M, uint32_t N, uint32_t n_classes,
                                       float* grad_sigmas, float* grad_rgbs, float* grad_local_semantics,
                                       void* stream);
/* inference wavefront: raymarching.cu:647-729 (+ semantics), order-preserving compaction for :837-855 */
/* semantic input: class probabilities local_semantics [M,C] f32, OR the semantic head's fp16 logits [M,logits_ld]
 * (the soft-max over the first C columns is then taken inside the kernel), or neither (rgb + depth only). */
UCSA_API int ucsa_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                        const float* sigmas, const float* rgbs, const float* local_semantics, const void* logits_h,
                        uint32_t logits_ld, const float* deltas, uint32_t n_classes, float* weights_sum, float* depth,
                        float* image, float* semantics, void* stream);
/* scratch (optional): ceil(n_alive / 1024) ints; with it a long list is compacted by many CTAs (count, scan, write)
 * instead of one CTA walking it -- the result is the same order-preserving list. */
UCSA_API int ucsa_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                      const float* rays_t_old, int32_t* alive_counter, int32_t* scratch, void* stream);
/* occupancy grid maintenance (not in the reference; torch-ngp's rule): grid = max(grid*decay, fresh) where fresh >= 0;
 * bitfield bit i = grid[i] > min(0.01, mean_density). */
UCSA_API int ucsa_grid_update(float* density_grid, const float* fresh, uint64_t n_cells, float decay, void* stream);
UCSA_API int ucsa_grid_packbits(const float* density_grid, uint64_t n_cells, float mean_density, uint32_t* bitfield,
                       void* stream);
/* The whole grid refresh as a kernel chain (the reference allocates density_grid, renderer_semantics.py:89-103, but
 * ships no update; rule after torch-ngp, which it is adapted from):
 * ucsa_grid_density: sigma of ONE jittered point per cell of the [cascades, H, H, H] grid (cell order of
 *   raymarching.cu:204) -- hash-grid encode + sigma MLP + trunc_exp in the density kernel, the points are generated
 *   on the fly (seed = counter-based stream), nothing but the [cascades*H^3] sigmas touches memory;
 * ucsa_grid_update_pack: grid = max(grid*decay, fresh*fresh_scale), mean_density_dev[0] = mean(max(grid,0)),
 *   bitfield bit i = grid[i] > min(0.01, mean).  sum_scratch: one double owned by the caller. */
UCSA_API int ucsa_grid_density(const void* table_h, const ucsa_grid_desc* grid_host, const void* w_sigma_h, float bound,
                       uint32_t cascades, uint32_t grid_h, uint64_t seed, float* sigma_cells, void* stream);
UCSA_API int ucsa_grid_update_pack(float* density_grid, const float* fresh, uint64_t n_cells, float decay,
                       float fresh_scale, double* sum_scratch, float* mean_density_dev, uint32_t* bitfield,
                       void* stream);

/* ---- stand-alone encoders / MLP, the module-level API the reference network exposes
 * (self.encoder, self.encoder_dir, tcnn.Network; network_tcnn_semantics.py:108,117,121,125). */
UCSA_API int ucsa_hashgrid_fwd(const float* x01, uint32_t n, const void* table_h, const ucsa_grid_desc* grid_host,
                      void* enc, void* stream);
UCSA_API int ucsa_hashgrid_bwd(const float* x01, uint32_t n, const ucsa_grid_desc* grid_host, const void* d_enc,
                      float inv_loss_scale, float* grad_table, void* stream);
UCSA_API int ucsa_hashgrid_indices(const float* x01, uint32_t n, const ucsa_grid_desc* grid_host, uint32_t* idx,
                          void* stream); /* [n,16,8] global entry indices, the bit-exact contract */
UCSA_API int ucsa_sh4_fwd(const float* d01, uint32_t n, void* out_h, void* stream);
/* dims: n_layers+1 widths, each a multiple of 16 and <= 64; x fp16 [n,dims[0]] -> y fp16 [n,dims[L]];
 * acts fp16 [n, sum(dims[1..L-1])] saved when non-null. */
UCSA_API int ucsa_mlp_fwd(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims_host, uint32_t n_layers,
                 void* y_h, void* acts_h, void* stream);
UCSA_API int ucsa_mlp_bwd(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims_host, uint32_t n_layers,
                 const void* acts_h, const void* dy_h, float inv_loss_scale, void* dx_h, float* grad_w,
                 void* stream);

UCSA_API int ucsa_density_fwd_simt(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                     const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
                     const void* table_h, const ucsa_grid_desc* grid_host, const void* w_sigma_h,
                     float* sigma, void* h, void* enc, void* hid, void* stream);
UCSA_API int ucsa_density_bwd_simt(const float* xyz, const float* rays_o, const float* rays_d, const float* aabb6,
                     const float* z_cat, uint32_t n_rays, uint32_t t, uint32_t k0, uint32_t k1, float bound,
                     const ucsa_grid_desc* grid_host, const void* w_sigma_h, const void* h, const void* enc,
                     const void* hid, const float* d_sigma, const void* dh, const uint8_t* use_geo,
                     float loss_scale, float* grad_table, float* grad_w_sigma, void* stream);
UCSA_API int ucsa_heads_fwd_simt(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t, uint32_t k_max,
                   const float* rays_d, const void* h, const void* w_color_h, const void* w_sem_h,
                   uint32_t n_classes, float* rgb, void* logits, void* hc1, void* hc2, void* hs, void* stream);
UCSA_API int ucsa_heads_bwd_simt(const int32_t* sel, const int32_t* ray_off, uint32_t n_rays, uint32_t t, uint32_t k_max,
                   const float* rays_d, const void* h, const void* w_color_h, const void* w_sem_h,
                   uint32_t n_classes, const float* rgb, const void* hc1, const void* hc2, const void* hs,
                   const float* d_rgb, const float* d_logits, float loss_scale, void* dh, float* grad_w_color,
                   float* grad_w_sem, void* stream);
/* CUDA-core (no tensor core) realisation of the same MLP contract; kept as the on-device cross-check of the
 * tcgen05 path (tests/) -- not used by the rendering pipeline. */
UCSA_API int ucsa_mlp_fwd_simt(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims_host,
                               uint32_t n_layers, void* y_h, void* acts_h, void* stream);
UCSA_API int ucsa_mlp_bwd_simt(const void* x_h, uint32_t n, const void* w_h, const uint32_t* dims_host,
                               uint32_t n_layers, const void* acts_h, const void* dy_h, float inv_loss_scale,
                               void* dx_h, float* grad_w, void* stream);

/* ---- parameter plumbing: fp32 master -> fp16 working copy; fused Adam (row f1:
 * joint_train_lightning_net.py:897-919: lr, betas (0.9,0.99), eps 1e-15, weight_decay on the MLPs).  lr and the
 * betas are doubles: torch.optim.Adam forms 1 - beta, the bias corrections and the step size in Python doubles and
 * rounds once, and the kernel does the same so that it tracks torch to the last bits.  weight_decay applies to the
 * parameters with index >= wd_begin (a multiple of 4): one launch covers the "encoding" group (no decay) followed by
 * the "net" group of a flat parameter buffer.  Gradients are multiplied by grad_scale_inv and, when grad_scale_dev
 * is given (torch.amp.GradScaler keeps its scale on the device), by 1 / grad_scale_dev[0]. */
UCSA_API int ucsa_cast_f32_to_f16(const float* src, uint64_t n, void* dst_h, void* stream);
UCSA_API int ucsa_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_h,
                   uint64_t n, double lr, double beta1, double beta2, float eps, float weight_decay,
                   uint64_t wd_begin, float grad_scale_inv, const float* grad_scale_dev, const float* found_inf,
                   uint32_t step, const int32_t* step_dev, const int32_t* skipped_dev, void* stream);
/* GradScaler's overflow check (joint_train_lightning_net.py:46,509-513; torch.amp.GradScaler.unscale_ / step): one
 * pass over a gradient buffer.  found_inf[0] = 1 if any entry is inf / NaN else 0; when it is 1 and skipped_dev is
 * given, skipped_dev[0] += 1 (ucsa_adam_step / ucsa_adam_exchange then skip the update, and Adam's step count is
 * step_dev[0] - skipped_dev[0]).  scratch2: two zero-initialised uint32 words owned by the caller (re-armed by the
 * kernel, so the launch can be replayed from a CUDA graph). */
UCSA_API int ucsa_grad_check(const float* grad, uint64_t n, float* found_inf, int32_t* skipped_dev, uint32_t* scratch2,
                   void* stream);

/* ---- (e)+f1. gradient exchange fused into Adam over peer memory (the reference: DDP all-reduce inside Lightning,
 * then torch.optim.Adam).  All ranks keep parameters (fp32 masters + fp16 copy) and the gradient buffer in symmetric
 * memory; *_ptrs_host are host arrays of `world` device addresses (rank order) of those flat buffers, mc_* their
 * multicast addresses (NVLS; all three or all null -> plain peer loads / stores).  The calling rank owns parameters
 * [begin, end): it sums the gradients of all ranks, applies Adam (moments exp_avg / exp_avg_sq hold end-begin
 * entries; weight decay applies to parameters >= wd_begin) and writes the new values into every rank's buffers.
 * found_inf_ptrs_host (optional): `world` addresses of the ranks' overflow flags (ucsa_grad_check, in symmetric
 * memory): when any is set the update is skipped on every rank and skipped_dev[0] += 1 on the caller.
 * broadcast_masters = 0 keeps the fp32 masters of a slice on its owner only (param_ptrs_host[rank] is then the only
 * master address used; mc_param may be null): peers receive just the fp16 working copy the kernels read, and a
 * checkpoint gathers the owned slices explicitly.
 * Bracket the launch with cross-rank barriers: gradients complete before, parameters visible after. */
UCSA_API int ucsa_adam_exchange(const uint64_t* grad_ptrs_host, const uint64_t* param_ptrs_host,
                   const uint64_t* param_h_ptrs_host, const float* mc_grad, float* mc_param, void* mc_param_h,
                   uint32_t world, uint32_t rank, uint64_t begin, uint64_t end, uint64_t wd_begin, float* exp_avg,
                   float* exp_avg_sq, double lr, double beta1, double beta2, float eps, float weight_decay,
                   uint32_t step, const int32_t* step_dev, const uint64_t* found_inf_ptrs_host, int32_t* skipped_dev,
                   int broadcast_masters, void* stream);

/* ---- f2. losses of forward_nerf_train (joint_train_lightning_net.py:199-221,503-507) and their gradients w.r.t.
 * image / depth / semantics in one kernel.  gt_rgb as fp16 [N,3] (batch["img_fp16"]) or fp32; labels int64 with -1 =
 * ignore; loss4 = (total, colour, semantics, depth).  total is multiplied by global_scale (1/world for sharded rays).
 * scratch: UCSA_LOSS_SCRATCH_BYTES of device memory owned by the caller, zero-filled once (the kernel re-arms it);
 * one block per engine / stream, so concurrent launches never share partial sums. */
UCSA_API int ucsa_nerf_loss(const float* image, const float* depth, const float* semantics, const void* gt_rgb_h,
                   const float* gt_rgb_f, const int64_t* labels, const float* gt_depth, uint32_t n_rays,
                   uint32_t n_classes, float one_m_to_scene_uom, float weight_semantics, float weight_depth,
                   float global_scale, float* loss4, float* g_image, float* g_depth, float* g_semantics,
                   void* scratch, void* stream);

/* ---- f2 (front). pinhole rays of the pixels `inds` [N] (row-major, may repeat; null = pixels 0..N-1) of one view:
 * dataset/ngp_utils.py:28-70 (get_rays), lightning/joint_train_lightning_net.py:109-151 (get_rays_train).
 * pose16 = cam2world 4x4 row-major on the device (nerf_matrix_to_ngp convention).  Outputs rays_o, rays_d [N,3],
 * direction_norms [N] (the norm of the un-normalised camera-frame direction). */
UCSA_API int ucsa_generate_rays(const float* pose16, float fx, float fy, float cx, float cy, uint32_t width,
                   uint32_t height, const int64_t* inds, uint32_t n, float* rays_o, float* rays_d,
                   float* direction_norms, void* stream);
/* ground truth of those pixels (joint_train_lightning_net.py:180-187: three torch.gather): image fp16 channel planes
 * [C,H*W] (batch["img_fp16"]), labels int64 [H*W] and depth f32 [H*W] (either may be null with its output)
 * -> gt_rgb fp16 [N,C], gt_labels [N], gt_depth [N]. */
UCSA_API int ucsa_gather_gt(const void* image_h, const int64_t* labels, const float* depth, uint64_t hw, uint32_t channels,
                   const int64_t* inds, uint32_t n, void* gt_rgb_h, int64_t* gt_labels, float* gt_depth, void* stream);

/* ---- f3. pseudo-label epilogue of a rendered view (joint_train_lightning_net.py:246-250,755-768): rows of semantics
 * [N,C] without mass become uniform, rows are normalised, label_u8 = argmax + 1 (lowest index on ties);
 * rgb_u8 [N,3] = (image * 255) truncated to u8, channel order RGB or BGR (cv2).  Either output may be null. */
UCSA_API int ucsa_label_epilogue(const float* image, const float* semantics, uint32_t n_pixels, uint32_t n_classes, int bgr,
                   uint8_t* label_u8, uint8_t* rgb_u8, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UCSA_NERF_H */
