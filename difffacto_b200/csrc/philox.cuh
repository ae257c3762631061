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
  const float ra = sqrtf(-2.f * logf(u0));
  const float rb = sqrtf(-2.f * logf(u2));
  float sa, ca, sb, cb;
  sincospif(2.f * u1, &sa, &ca);
  sincospif(2.f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
#endif

}  // namespace dfb200
