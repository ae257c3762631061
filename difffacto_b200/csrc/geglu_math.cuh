// GEGLU / dropout device math shared by the training kernels (train_ops.cu) and the fused FF-in GEMM epilogue (gemm_tc.cu).
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace dfb200 {
// Phi(g) and phi(g) of the standard normal with ONE exponential: erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z),
// z = |g| / sqrt 2 (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7 -- the accuracy of erff's fp32 result), and phi(g) is that same
// exp(-g^2 / 2) / sqrt(2 pi).  erff + expf cost ~40 instructions per element and bound the fused GEGLU kernels; this costs ~18.
__device__ __forceinline__ void normal_cdf_pdf(float g, float& cdf, float& pdf) {
  const float z = fabsf(g) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-z * z);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.f);
  cdf = 0.5f * (1.f + copysignf(erf_abs, g));
  pdf = 0.39894228040143267794f * e;
}
// Dropout keep-mask of 4 consecutive elements (Philox quad = element index / 4): the mask dfb200_dropout draws, so fused and unfused
// paths are interchangeable.
__device__ __forceinline__ void dropout_keep4(float p, float scale, uint64_t seed, uint64_t offset, long long quad, float (&m)[4]) {
  uint32_t r[4];
  philox4x32_10((uint64_t)quad, offset, seed, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = ((float)(r[i] >> 8) * 5.9604644775390625e-8f >= p) ? scale : 0.f;
}
}  // namespace dfb200
