// The anchored-DDPM per-element arithmetic, shared by the standalone update kernel (ddpm.cu) and
// the fused denoiser epilogue.  Every product and sum is rounded separately (explicit
// __fmul_rn/__fadd_rn, no FMA contraction) in the order the reference evaluates its torch ops, so
// fp32 results are bit-identical to the reference given the same eps and noise.
#pragma once
#include "common.cuh"

namespace dfb200 {

struct StepCoef {  // float32(table[t]) of anchored_diffusion.py:62-112
  float sqrt_recip, sqrt_recipm1, post_var, c1, c2, c3;
  float nonzero;  // (t != 0).float()
};

__device__ __forceinline__ StepCoef load_step_coef(const float* __restrict__ sched, int T, int t) {
  StepCoef c;
  c.sqrt_recip = __ldg(sched + DFB200_SCHED_SQRT_RECIP_ALPHAS_CUMPROD * T + t);
  c.sqrt_recipm1 = __ldg(sched + DFB200_SCHED_SQRT_RECIPM1_ALPHAS_CUMPROD * T + t);
  c.post_var = __ldg(sched + DFB200_SCHED_POSTERIOR_VARIANCE * T + t);
  c.c1 = __ldg(sched + DFB200_SCHED_POSTERIOR_MEAN_COEF1 * T + t);
  c.c2 = __ldg(sched + DFB200_SCHED_POSTERIOR_MEAN_COEF2 * T + t);
  c.c3 = __ldg(sched + DFB200_SCHED_POSTERIOR_MEAN_COEF3 * T + t);
  c.nonzero = t != 0 ? 1.f : 0.f;
  return c;
}

// pred_xstart: anchored_diffusion.py:401-409 (_predict_xstart_from_eps, learn_anchor path)
//   sqrt_recip*(x_t - a) + a - sqrt_recipm1 * sqrt(var) * eps
__device__ __forceinline__ float ddpm_xstart(const StepCoef& c, float x, float a, float var, float eps) {
  const float L = __fsqrt_rn(var);
  const float lhs = __fadd_rn(__fmul_rn(c.sqrt_recip, __fsub_rn(x, a)), a);
  const float rhs = __fmul_rn(__fmul_rn(c.sqrt_recipm1, L), eps);
  return __fsub_rn(lhs, rhs);
}

// sample: q_posterior_mean (:184-188) + fixed-small variance scaled by the point variance (:313)
// + `mean + nonzero_mask * sqrt(variance) * noise` (:476-483)
__device__ __forceinline__ float ddpm_prev(const StepCoef& c, float x, float a, float var, float x0, float z) {
  const float mean = __fadd_rn(__fadd_rn(__fmul_rn(c.c1, x0), __fmul_rn(c.c2, x)), __fmul_rn(c.c3, a));
  const float sd = __fsqrt_rn(__fmul_rn(c.post_var, var));
  return __fadd_rn(mean, __fmul_rn(__fmul_rn(c.nonzero, sd), z));
}

// DDIM variant (anchored_diffusion.py:368-374 xt_dir, :480-481 sample) in the reference's op order:
//   sample = (((x0 - a) * sqrt(acp_t) + a) + (L * dir_t) * eps) + ((eta * nonzero) * sqrt(pv_t * var)) * z
// sqrt_acp = sqrt(float32(alphas_cumprod_prev[t])), dir = float32(sqrt(1 - ac - eta^2 * posterior_variance))[t]
__device__ __forceinline__ float ddim_prev(const StepCoef& c, float a, float var, float x0, float eps, float z, float sqrt_acp,
                                           float dir, float eta) {
  const float L = __fsqrt_rn(var);
  const float lhs = __fadd_rn(__fmul_rn(__fsub_rn(x0, a), sqrt_acp), a);
  const float xt_dir = __fmul_rn(__fmul_rn(L, dir), eps);
  const float sd = __fsqrt_rn(__fmul_rn(c.post_var, var));
  const float nz = __fmul_rn(__fmul_rn(__fmul_rn(eta, c.nonzero), sd), z);
  return __fadd_rn(__fadd_rn(lhs, xt_dir), nz);
}

}  // namespace dfb200
