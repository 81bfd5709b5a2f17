// The two ends of the rendering path that the reference runs as eager torch / numpy / OpenCV code
// (rows f2 and f3 of SURVEY.md section 8):
//   * ray generation for a pinhole camera + gather of the ground truth of the sampled pixels
//     (dataset/ngp_utils.py:28-70 get_rays, lightning/joint_train_lightning_net.py:109-151 get_rays_train, :180-187);
//   * the pseudo-label epilogue of a rendered view: zero-sum -> uniform, normalise, arg-max -> label + 1 as u8,
//     colour * 255 -> u8 (joint_train_lightning_net.py:246-250, :755-768).
#include "common.cuh"

namespace ucsa {
namespace {

// pixel p (row-major, p = row * W + col) of a pinhole camera: i = col + 0.5, j = row + 0.5;
// d_cam = ((i - cx) / fx, (j - cy) / fy, 1); norm = |d_cam|; rays_d = R (d_cam / norm); rays_o = t.
// Arithmetic as the eager fp32 expression (no FMA contraction); the 3x3 product accumulates k = 0, 1, 2 in order.
__global__ void generate_rays_kernel(const float* __restrict__ pose, float fx, float fy, float cx, float cy,
                                     uint32_t width, const int64_t* __restrict__ inds, uint32_t n,
                                     float* __restrict__ rays_o, float* __restrict__ rays_d,
                                     float* __restrict__ direction_norms) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint64_t p = inds != nullptr ? static_cast<uint64_t>(inds[r]) : r;
  const float i = __fadd_rn(static_cast<float>(p % width), 0.5f);
  const float j = __fadd_rn(static_cast<float>(p / width), 0.5f);
  const float x = __fdiv_rn(__fsub_rn(i, cx), fx), y = __fdiv_rn(__fsub_rn(j, cy), fy), z = 1.0f;
  const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  const float d[3] = {__fdiv_rn(x, norm), __fdiv_rn(y, norm), __fdiv_rn(z, norm)};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float acc = __fmul_rn(d[0], pose[4 * k + 0]);
    acc = __fadd_rn(acc, __fmul_rn(d[1], pose[4 * k + 1]));
    acc = __fadd_rn(acc, __fmul_rn(d[2], pose[4 * k + 2]));
    rays_d[3ull * r + k] = acc;
    rays_o[3ull * r + k] = pose[4 * k + 3];
  }
  direction_norms[r] = norm;
}

// ground truth of the sampled pixels from device-resident planes: image fp16 [C,H*W] (batch["img_fp16"], channel
// planes), labels int64 [H*W], depth f32 [H*W]  ->  gt_rgb fp16 [N,C], labels [N], depth [N]  (torch.gather x 3)
__global__ void gather_gt_kernel(const __half* __restrict__ image, const int64_t* __restrict__ labels,
                                 const float* __restrict__ depth, uint64_t hw, uint32_t channels,
                                 const int64_t* __restrict__ inds, uint32_t n, __half* __restrict__ gt_rgb,
                                 int64_t* __restrict__ gt_labels, float* __restrict__ gt_depth) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint64_t p = static_cast<uint64_t>(inds[r]);
  for (uint32_t c = 0; c < channels; ++c) gt_rgb[static_cast<uint64_t>(r) * channels + c] = image[c * hw + p];
  if (labels != nullptr) gt_labels[r] = labels[p];
  if (depth != nullptr) gt_depth[r] = depth[p];
}

// one warp per pixel: lanes stride over the classes
__global__ void __launch_bounds__(256)
label_epilogue_kernel(const float* __restrict__ image, const float* __restrict__ semantics, uint32_t n, uint32_t c,
                      int bgr, uint8_t* __restrict__ label_u8, uint8_t* __restrict__ rgb_u8) {
  const int lane = threadIdx.x & 31;
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const float* s = semantics + static_cast<uint64_t>(r) * c;
  float part = 0.f;
  for (uint32_t k = lane; k < c; k += 32) part += s[k];
  const float total = warp_sum(part);
  const bool invalid = total == 0.f;                           // no semantic mass: every class becomes 1 / C
  const float denom = invalid ? static_cast<float>(c) : total;
  float best = -INFINITY;
  uint32_t best_k = 0xffffffffu;
  for (uint32_t k = lane; k < c; k += 32) {
    const float p = __fdiv_rn(invalid ? 1.0f : s[k], denom);   // arg-max of the NORMALISED values, as the reference
    if (p > best) {
      best = p;
      best_k = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // first index among equal maxima (torch.argmax on a tie-free row; ties: lowest)
    const float ob = __shfl_xor_sync(kFullMask, best, o);
    const uint32_t ok = __shfl_xor_sync(kFullMask, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (lane == 0 && label_u8 != nullptr) label_u8[r] = static_cast<uint8_t>(best_k + 1u);  // stored labels are class + 1
  if (lane < 3 && rgb_u8 != nullptr && image != nullptr) {
    const float v = image[3ull * r + lane] * 255.0f;  // numpy astype(uint8) of a value in [0, 255]: truncation
    const int out_c = bgr ? 2 - lane : lane;
    rgb_u8[3ull * r + out_c] = static_cast<uint8_t>(static_cast<int>(fminf(fmaxf(v, 0.f), 255.f)));
  }
}

}  // namespace
}  // namespace ucsa

using namespace ucsa;

extern "C" int ucsa_generate_rays(const float* pose16, float fx, float fy, float cx, float cy, uint32_t width,
                                  uint32_t height, const int64_t* inds, uint32_t n, float* rays_o, float* rays_d,
                                  float* direction_norms, void* stream) {
  UCSA_REQUIRE(pose16 && rays_o && rays_d && direction_norms, "generate_rays: null pointer");
  UCSA_REQUIRE(width >= 1 && height >= 1 && fx != 0.f && fy != 0.f, "generate_rays: bad camera");
  UCSA_REQUIRE(inds != nullptr || n <= static_cast<uint64_t>(width) * height,
               "generate_rays: without pixel indices n must not exceed width * height");
  if (n == 0) return UCSA_OK;
  generate_rays_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(pose16, fx, fy, cx, cy, width, inds, n, rays_o,
                                                                        rays_d, direction_norms);
  return check_launch("generate_rays");
}

extern "C" int ucsa_gather_gt(const void* image_h, const int64_t* labels, const float* depth, uint64_t hw,
                              uint32_t channels, const int64_t* inds, uint32_t n, void* gt_rgb_h, int64_t* gt_labels,
                              float* gt_depth, void* stream) {
  UCSA_REQUIRE(image_h && inds && gt_rgb_h, "gather_gt: null pointer");
  UCSA_REQUIRE((labels == nullptr) == (gt_labels == nullptr) && (depth == nullptr) == (gt_depth == nullptr),
               "gather_gt: give labels / depth together with their outputs");
  UCSA_REQUIRE(channels >= 1 && channels <= 4 && hw >= 1, "gather_gt: 1..4 image channels");
  if (n == 0) return UCSA_OK;
  gather_gt_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(static_cast<const __half*>(image_h), labels, depth,
                                                                    hw, channels, inds, n,
                                                                    static_cast<__half*>(gt_rgb_h), gt_labels,
                                                                    gt_depth);
  return check_launch("gather_gt");
}

extern "C" int ucsa_label_epilogue(const float* image, const float* semantics, uint32_t n_pixels, uint32_t n_classes,
                                   int bgr, uint8_t* label_u8, uint8_t* rgb_u8, void* stream) {
  UCSA_REQUIRE(semantics != nullptr, "label_epilogue: null semantics");
  UCSA_REQUIRE(n_classes >= 1 && n_classes <= 255, "label_epilogue: 1 <= classes <= 255 (labels are class + 1 as u8)");
  UCSA_REQUIRE(rgb_u8 == nullptr || image != nullptr, "label_epilogue: rgb output needs the image");
  if (n_pixels == 0) return UCSA_OK;
  label_epilogue_kernel<<<ceil_div(static_cast<uint64_t>(n_pixels) * 32, 256), 256, 0, as_stream(stream)>>>(
      image, semantics, n_pixels, n_classes, bgr, label_u8, rgb_u8);
  return check_launch("label_epilogue");
}
