// Real spherical harmonics, degree 4 (16 coefficients), of a unit direction.
// Row a7 of SURVEY.md section 8 (tcnn SphericalHarmonics, network_tcnn_semantics.py:64-70,116-117,164-165).
// The reference maps d to [0,1] ((d+1)/2) because tcnn maps it back with 2*x-1; both steps are kept so the
// fp32 value entering the polynomials is the same one.
#pragma once

#include "common.cuh"

namespace ucsa {

// input already mapped to [0,1] (what tcnn's encoder receives)
__device__ __forceinline__ void sh4_from01(float ux, float uy, float uz, float (&out)[16]) {
  const float x = __fsub_rn(__fmul_rn(ux, 2.0f), 1.0f);
  const float y = __fsub_rn(__fmul_rn(uy, 2.0f), 1.0f);
  const float z = __fsub_rn(__fmul_rn(uz, 2.0f), 1.0f);
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  out[0] = 0.28209479177387814f;
  out[1] = -0.48860251190291987f * y;
  out[2] = 0.48860251190291987f * z;
  out[3] = -0.48860251190291987f * x;
  out[4] = 1.0925484305920792f * xy;
  out[5] = -1.0925484305920792f * yz;
  out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  out[7] = -1.0925484305920792f * xz;
  out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  out[10] = 2.8906114426405538f * xy * z;
  out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  out[14] = 1.4453057213202769f * z * (x2 - y2);
  out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// input is the ray direction in [-1,1]; (d+1)/2 as at network_tcnn_semantics.py:116,164
__device__ __forceinline__ void sh4_eval(float dx, float dy, float dz, float (&out)[16]) {
  sh4_from01(__fdiv_rn(__fadd_rn(dx, 1.0f), 2.0f), __fdiv_rn(__fadd_rn(dy, 1.0f), 2.0f),
             __fdiv_rn(__fadd_rn(dz, 1.0f), 2.0f), out);
}

}  // namespace ucsa
