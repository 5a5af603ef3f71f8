// The scalar arithmetic of the fused multi-scale head (head_kernel<K, HEAD_MS>): torch's upsample_bilinear2d
// (align_corners=False) tap weights and blend, and the division of one scale's value by the number of scales, with the
// FMA contraction pinned to the one torch's CPU and CUDA kernels evaluate -- anomaly/models/models.py:659-661 +
// anomaly/eval_ood_traditional.py:198-208 (`scores += interpolate(z) / 5`).
//
// __host__ __device__ and free of CUDA-only constructs: tests/host/ms_emulation.cpp runs the same functions on the CPU
// (fmaf is exact there too) and tests/test_ms_emulation.py compares the result bit for bit with a torch-CPU replay of the
// reference loop.  Host builds must use -ffp-contract=off so that bl_mul / bl_add stay separate roundings.
#pragma once
#include <math.h>

#ifndef DML_HD
#if defined(__CUDACC__)
#define DML_HD __host__ __device__ __forceinline__
#else
#define DML_HD inline
#endif
#endif

namespace dml {

DML_HD float bl_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
DML_HD float bl_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);   // never contracted into a following add
#else
  return a * b;
#endif
}
DML_HD float bl_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

// source coordinate of output index `dst` (torch area_pixel_compute_source_index, align_corners=False, clamped at 0):
// scale * (dst + 0.5) - 0.5 evaluated as one FMA; scale = float(in) / out
DML_HD float bilinear_src(float scale, int dst) {
  const float r = bl_fma(scale, dst + 0.5f, -0.5f);
  return r < 0.f ? 0.f : r;
}

// h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11) in torch's contraction
DML_HD float bilinear_blend(float w0, float w1, float h0, float h1, float v00, float v01, float v10, float v11) {
  const float t0 = bl_fma(w0, v00, bl_mul(w1, v01));
  const float t1 = bl_fma(w0, v10, bl_mul(w1, v11));
  return bl_fma(h0, t0, bl_mul(h1, t1));
}

// one scale's contribution val / n: correctly rounded (q = val * fl(1/n), one FMA residual step: Markstein) like the
// CPU `scores_tmp / 5`, or -- `recip` -- the plain product by fl(1/n) that torch's CUDA div-by-scalar kernel computes
DML_HD float scale_share(float val, float inv_n, float n, bool recip) {
  float t = bl_mul(val, inv_n);
  if (!recip) t = bl_fma(bl_fma(-n, t, val), inv_n, t);
  return t;
}

}  // namespace dml
