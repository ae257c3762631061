// Anchored DDPM elementwise kernels: eps -> x_{t-1} update, forward noising, x_T init, Philox.
// Reference: python/difffacto/models/diffusions/anchored_diffusion.py:148-173 (q_sample),
// :227-395 + :401-409 + :175-193 (p_mean_variance on the config path), :450-484 (p_sample), :564.
// The reference spends ~25 tiny torch kernels and 10 host->device table uploads per step here; this
// is one HBM-bound pass: 16 B read per element (x, eps|noise, anchor, variance) + 4..8 B written.
#include "common.cuh"
#include "ddpm.cuh"
#include "philox.cuh"

namespace dfb200 {

// One thread per 4 consecutive elements of the flattened (B,3,N) tensors (N % 4 == 0 path) or per
// element (generic path).
template <bool VEC4, bool PHILOX>
__global__ void __launch_bounds__(256)
ddpm_step_kernel(long long total, int per_sample, int T, const float* __restrict__ sched,
                 const int* __restrict__ t, const float* __restrict__ x_t, const float* __restrict__ eps,
                 const float* __restrict__ anchors, const float* __restrict__ variance,
                 const float* __restrict__ noise, uint64_t seed, uint64_t offset,
                 float* __restrict__ x_prev, float* __restrict__ pred_xstart) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC4) {
    const long long e = q * 4;
    if (e >= total) return;
    const int b = (int)(e / per_sample);
    const StepCoef c = load_step_coef(sched, T, __ldg(t + b));
    const float4 x = __ldg(reinterpret_cast<const float4*>(x_t + e));
    const float4 ep = __ldg(reinterpret_cast<const float4*>(eps + e));
    const float4 a = __ldg(reinterpret_cast<const float4*>(anchors + e));
    const float4 v = __ldg(reinterpret_cast<const float4*>(variance + e));
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (PHILOX) z = philox_normal4((uint64_t)q, offset, seed);
    else if (noise != nullptr) z = __ldg(reinterpret_cast<const float4*>(noise + e));
    float4 x0, xp;
    x0.x = ddpm_xstart(c, x.x, a.x, v.x, ep.x); xp.x = ddpm_prev(c, x.x, a.x, v.x, x0.x, z.x);
    x0.y = ddpm_xstart(c, x.y, a.y, v.y, ep.y); xp.y = ddpm_prev(c, x.y, a.y, v.y, x0.y, z.y);
    x0.z = ddpm_xstart(c, x.z, a.z, v.z, ep.z); xp.z = ddpm_prev(c, x.z, a.z, v.z, x0.z, z.z);
    x0.w = ddpm_xstart(c, x.w, a.w, v.w, ep.w); xp.w = ddpm_prev(c, x.w, a.w, v.w, x0.w, z.w);
    *reinterpret_cast<float4*>(x_prev + e) = xp;
    if (pred_xstart != nullptr) *reinterpret_cast<float4*>(pred_xstart + e) = x0;
  } else {
    if (q >= total) return;
    const int b = (int)(q / per_sample);
    const StepCoef c = load_step_coef(sched, T, __ldg(t + b));
    float z = 0.f;
    if (PHILOX) {
      const float4 z4 = philox_normal4((uint64_t)(q >> 2), offset, seed);
      const int l = (int)(q & 3);
      z = l == 0 ? z4.x : l == 1 ? z4.y : l == 2 ? z4.z : z4.w;
    } else if (noise != nullptr) {
      z = __ldg(noise + q);
    }
    const float x = __ldg(x_t + q), a = __ldg(anchors + q), v = __ldg(variance + q);
    const float x0 = ddpm_xstart(c, x, a, v, __ldg(eps + q));
    x_prev[q] = ddpm_prev(c, x, a, v, x0, z);
    if (pred_xstart != nullptr) pred_xstart[q] = x0;
  }
}

// x_t = sqrt_ac*(x0 - a) + a + sqrt_1mac * sqrt(var) * noise      (anchored_diffusion.py:169-173)
__global__ void __launch_bounds__(256)
q_sample_kernel(long long total, int per_sample, int T, const float* __restrict__ sched,
                const int* __restrict__ t, const float* __restrict__ x_start,
                const float* __restrict__ anchors, const float* __restrict__ variance,
                const float* __restrict__ noise, float* __restrict__ x_t) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int b = (int)(q / per_sample);
  const int tt = __ldg(t + b);
  const float sa = __ldg(sched + DFB200_SCHED_SQRT_ALPHAS_CUMPROD * T + tt);
  const float sb = __ldg(sched + DFB200_SCHED_SQRT_ONE_MINUS_ALPHAS_CUMPROD * T + tt);
  const float a = __ldg(anchors + q);
  const float L = __fsqrt_rn(__ldg(variance + q));
  const float lhs = __fadd_rn(__fmul_rn(sa, __fsub_rn(__ldg(x_start + q), a)), a);
  x_t[q] = __fadd_rn(lhs, __fmul_rn(__fmul_rn(sb, L), __ldg(noise + q)));
}

// x_T = sqrt(var) * z + anchors   (anchored_diffusion.py:564); z read from x itself or Philox
template <bool PHILOX>
__global__ void __launch_bounds__(256)
xT_init_kernel(long long total, float* __restrict__ x, const float* __restrict__ anchors,
               const float* __restrict__ variance, uint64_t seed, uint64_t offset) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  float z;
  if (PHILOX) {
    const float4 z4 = philox_normal4((uint64_t)(q >> 2), offset, seed);
    const int l = (int)(q & 3);
    z = l == 0 ? z4.x : l == 1 ? z4.y : l == 2 ? z4.z : z4.w;
  } else {
    z = x[q];
  }
  x[q] = __fadd_rn(__fmul_rn(__fsqrt_rn(__ldg(variance + q)), z), __ldg(anchors + q));
}

__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ out, size_t count, uint64_t seed, uint64_t offset) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t e = q * 4;
  if (e >= count) return;
  const float4 z = philox_normal4((uint64_t)q, offset, seed);
  if (e + 3 < count && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    *reinterpret_cast<float4*>(out + e) = z;
  } else {
    out[e] = z.x;
    if (e + 1 < count) out[e + 1] = z.y;
    if (e + 2 < count) out[e + 2] = z.z;
    if (e + 3 < count) out[e + 3] = z.w;
  }
}

// DDIM variant of the reverse step (anchored_diffusion.py:368-374 xt_dir, :480-481 sample), evaluated in the reference's
// op order:  sample = (((x0 - a) * sqrt(acp_t) + a) + (L * dir_t) * eps) + ((eta * nonzero) * sqrt(pv_t * var)) * z
// acp = float32(alphas_cumprod_prev), dir = float32(sqrt(1 - ac - eta^2 * posterior_variance)), both indexed by the
// FULL-schedule timestep t (the reference does not re-derive them for the strided DDIM step list).
__global__ void __launch_bounds__(256)
ddim_step_kernel(long long total, int per_sample, int T, const float* __restrict__ sched, const int* __restrict__ t,
                 const float* __restrict__ x_t, const float* __restrict__ eps, const float* __restrict__ anchors,
                 const float* __restrict__ variance, const float* __restrict__ noise, const float* __restrict__ acp,
                 const float* __restrict__ dir_coeff, float eta, float* __restrict__ x_prev, float* __restrict__ pred_xstart) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int b = (int)(q / per_sample);
  const int tt = __ldg(t + b);
  const StepCoef c = load_step_coef(sched, T, tt);
  const float x = __ldg(x_t + q), a = __ldg(anchors + q), v = __ldg(variance + q), e = __ldg(eps + q);
  const float x0 = ddpm_xstart(c, x, a, v, e);
  x_prev[q] = ddim_prev(c, a, v, x0, e, noise != nullptr ? __ldg(noise + q) : 0.f, __fsqrt_rn(__ldg(acp + tt)), __ldg(dir_coeff + tt), eta);
  if (pred_xstart != nullptr) pred_xstart[q] = x0;
}

// classifier-free guidance mix (anchored_diffusion.py:263-266): out = (1 - w) * uncond + w * cond
__global__ void __launch_bounds__(256)
guidance_mix_kernel(long long total, float one_minus_w, float w, const float* __restrict__ uncond, const float* __restrict__ cond,
                    float* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  out[q] = __fadd_rn(__fmul_rn(one_minus_w, __ldg(uncond + q)), __fmul_rn(w, __ldg(cond + q)));
}

int launch_ddpm_step(int B, int N, int T, const float* sched, const int* t, const float* x_t,
                     const float* eps, const float* anchors, const float* variance, const float* noise,
                     bool philox, uint64_t seed, uint64_t offset, float* x_prev, float* pred_xstart,
                     cudaStream_t st) {
  const long long total = (long long)B * 3 * N;
  if (total == 0) return DFB200_OK;
  const int per_sample = 3 * N;
  auto aligned = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (N % 4 == 0) && aligned(x_t) && aligned(eps) && aligned(anchors) && aligned(variance) &&
                   aligned(noise) && aligned(x_prev) && aligned(pred_xstart);
  const int grid = cdiv(vec ? total / 4 : total, 256);
#define DDPM_LAUNCH(V, P)                                                                              \
  ddpm_step_kernel<V, P><<<grid, 256, 0, st>>>(total, per_sample, T, sched, t, x_t, eps, anchors, variance, \
                                               noise, seed, offset, x_prev, pred_xstart)
  if (vec && philox) DDPM_LAUNCH(true, true);
  else if (vec) DDPM_LAUNCH(true, false);
  else if (philox) DDPM_LAUNCH(false, true);
  else DDPM_LAUNCH(false, false);
#undef DDPM_LAUNCH
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

int launch_xT_init(long long total, float* x, const float* anchors, const float* variance, bool philox,
                   uint64_t seed, uint64_t offset, cudaStream_t st) {
  if (total == 0) return DFB200_OK;
  if (philox) xT_init_kernel<true><<<cdiv(total, 256), 256, 0, st>>>(total, x, anchors, variance, seed, offset);
  else xT_init_kernel<false><<<cdiv(total, 256), 256, 0, st>>>(total, x, anchors, variance, seed, offset);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_ddpm_step(int B, int N, int T, const float* sched, const int* t, const float* x_t,
                                const float* eps, const float* anchors, const float* variance,
                                const float* noise, float* x_prev, float* pred_xstart,
                                dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && T > 0, DFB200_ERR_INVALID_ARG, "ddpm_step: bad sizes B=%d N=%d T=%d", B, N, T);
  return launch_ddpm_step(B, N, T, sched, t, x_t, eps, anchors, variance, noise, false, 0, 0, x_prev,
                          pred_xstart, as_stream(stream));
}

extern "C" int dfb200_q_sample(int B, int N, int T, const float* sched, const int* t, const float* x_start,
                               const float* anchors, const float* variance, const float* noise, float* x_t,
                               dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && T > 0, DFB200_ERR_INVALID_ARG, "q_sample: bad sizes B=%d N=%d T=%d", B, N, T);
  const long long total = (long long)B * 3 * N;
  if (total == 0) return DFB200_OK;
  q_sample_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(total, 3 * N, T, sched, t, x_start, anchors,
                                                                   variance, noise, x_t);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_philox_normal(float* out, size_t count, uint64_t seed, uint64_t offset,
                                    dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  philox_normal_kernel<<<cdiv((long long)((count + 3) / 4), 256), 256, 0, as_stream(stream)>>>(out, count, seed, offset);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_ddim_step(int B, int N, int T, const float* sched, const int* t, const float* x_t, const float* eps,
                                const float* anchors, const float* variance, const float* noise,
                                const float* alphas_cumprod_prev, const float* xt_dir_coeff, float eta, float* x_prev,
                                float* pred_xstart, dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && T > 0, DFB200_ERR_INVALID_ARG, "ddim_step: bad sizes B=%d N=%d T=%d", B, N, T);
  const long long total = (long long)B * 3 * N;
  if (total == 0) return DFB200_OK;
  ddim_step_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(total, 3 * N, T, sched, t, x_t, eps, anchors, variance, noise,
                                                                    alphas_cumprod_prev, xt_dir_coeff, eta, x_prev, pred_xstart);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_guidance_mix(size_t count, float classifier_weight, const float* eps_uncond, const float* eps_cond,
                                   float* eps_out, dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  guidance_mix_kernel<<<cdiv((long long)count, 256), 256, 0, as_stream(stream)>>>((long long)count, 1.f - classifier_weight,
                                                                                  classifier_weight, eps_uncond, eps_cond, eps_out);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
