// Counter-based Philox4x32-10 + Box-Muller normals (Salmon et al., SC'11 constants).
// Element e of draw `offset` under `seed` is lane e%4 of philox(ctr = {lo(e/4), hi(e/4), lo(offset),
// hi(offset)}, key = seed); every kernel that consumes noise (ddpm_step, the fused denoiser
// epilogue, philox_normal) uses this mapping, so explicit-noise and in-kernel-noise runs agree.
#pragma once
#include <stdint.h>

namespace dfb200 {

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint64_t quad, uint64_t offset, uint64_t seed,
                                                      uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

#ifdef __CUDACC__
// four N(0,1) samples for elements 4*quad .. 4*quad+3
__device__ __forceinline__ float4 philox_normal4(uint64_t quad, uint64_t offset, uint64_t seed) {
  uint32_t r[4];
  philox4x32_10(quad, offset, seed, r);
  // u in (0,1): 24 random bits centred in their bucket
  const float u0 = (float)(r[0] >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
  const float u1 = (float)(r[1] >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
  const float u2 = (float)(r[2] >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
  const float u3 = (float)(r[3] >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
  // Box-Muller on the SFU: r = sqrt(-2 ln u) = sqrt(-2 ln2 * log2 u) (MUFU.LG2), angle uniform in (-pi, pi) (MUFU.SIN / MUFU.COS,
  // absolute error ~5e-7 on that range).  Explicitly rounded products: every kernel that inlines this draws bit-identical values.
  // (The libm forms logf / sincospif cost ~150 instructions per call on the critical path of the fused kernel's head phase.)
  const float ra = __fsqrt_rn(__fmul_rn(-1.3862943611198906f, __log2f(u0)));
  const float rb = __fsqrt_rn(__fmul_rn(-1.3862943611198906f, __log2f(u2)));
  const float ta = __fmaf_rn(u1, 6.283185307179586f, -3.141592653589793f);
  const float tb = __fmaf_rn(u3, 6.283185307179586f, -3.141592653589793f);
  const float sa = __sinf(ta), ca = __cosf(ta), sb = __sinf(tb), cb = __cosf(tb);
  return make_float4(__fmul_rn(ra, ca), __fmul_rn(ra, sa), __fmul_rn(rb, cb), __fmul_rn(rb, sb));
}
#endif

}  // namespace dfb200
