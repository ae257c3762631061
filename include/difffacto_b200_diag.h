/* difffacto_b200 -- DIAGNOSTIC entry points (tcgen05 building-block self-tests, MMA issue-rate microbenchmarks, the
 * phase-timeline hook of the fused denoiser).  They are NOT part of the product ABI: libdifffacto_b200.so does not
 * export them.  They exist only in libdifffacto_b200_diag.so, the same sources compiled with -DDFB200_DIAGNOSTICS
 * (`python -m difffacto_b200.build --diag`), which tests/ and tools/ load through difffacto_b200._lib.load_diag(). */
#ifndef DIFFFACTO_B200_DIAG_H_
#define DIFFFACTO_B200_DIAG_H_
#include "difffacto_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Library self-test of the tcgen05 building blocks the bf16 denoiser is made of (canonical K-major
 * UMMA shared-memory tiles, cp.async.bulk staging, TMEM alloc/store/load, accumulate onto pre-stored
 * TMEM, bias as an extra MMA against a ones tile): D[128,N] = Cin + A[128,K].W[N,K]^T + bias with bf16
 * operands and fp32 accumulation.  N in {32,64,128}, K a multiple of 16 <= 128; bias/Cin may be NULL;
 * variant: bit0 swap LBO/SBO roles, bit1 two-slab bias tile, bit2 stage W through `scratch`
 * (>= N*K*2 bytes) with cp.async.bulk.  All pointers are device pointers. */
int dfb200_selftest_umma(int variant, int N, int K, const float* A, const float* W, const float* bias,
                         const float* Cin, float* D, void* scratch, dfb200_stream_t stream);

/* The same with a CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256): D[256,N] = Cin + A[256,K].W[N,K]^T + bias.
 * CTA rank r stages A rows [128r,128r+128) and W rows [r*N/2,(r+1)*N/2); rank 1 signals rank 0 by remote mbarrier
 * arrives, rank 0 issues the MMAs and a multicast commit.  variant: bit2 stage W through `scratch` with cp.async.bulk,
 * bit3 A operand from tensor memory (TS form). */
int dfb200_selftest_umma2(int variant, int N, int K, const float* A, const float* W, const float* bias,
                          const float* Cin, float* D, void* scratch, dfb200_stream_t stream);

/* Microbenchmark: SM cycles for `iters` back-to-back tcgen05.mma (M=128, N, K=16, bf16) from shared-memory operands
 * cycling over `ksteps` K-slabs; layout 0 = canonical no-swizzle tiles, 1 = SWIZZLE_128B.  out_cycles: device int64. */
int dfb200_bench_umma(int layout, int N, int iters, int ksteps, long long* out_cycles, dfb200_stream_t stream);
/* Same for a CTA pair (cta_group::2, M = 256); mode bit0: A operand from tensor memory, bit1: single accumulator. */
int dfb200_bench_umma2(int mode, int N, int iters, int ksteps, long long* out_cycles, dfb200_stream_t stream);

/* Profiling hook: a device buffer of 1024 int64 that CTA 0 of the fused bf16 denoiser kernel fills with
 * clock64() stamps at the phase boundaries of its `item`-th work item ([0..511] tile-0 epilogue, [512..1023] MMA
 * issuer); NULL disables. */
int dfb200_debug_tc_timeline(long long* device_buffer, int item);

#ifdef __cplusplus
}
#endif
#endif /* DIFFFACTO_B200_DIAG_H_ */
