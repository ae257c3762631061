// tcgen05 building-block self-test and issue-rate microbenchmark (dfb200_selftest_umma, dfb200_bench_umma).
// Diagnostics only: compiled into libdifffacto_b200_diag.so (-DDFB200_DIAGNOSTICS), never into the product library.
#ifdef DFB200_DIAGNOSTICS
#include "../../include/difffacto_b200_diag.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace dfb200 {
using namespace tc;

// ---------------------------------------------------------------------------------------------
// UMMA self-test: D[128 x N] = (Cin) + A[128 x K] . W[N x K]^T (+ bias), bf16 operands, one CTA.
// Exercises exactly the building blocks of the fused kernel: canonical no-swizzle K-major tiles written
// by threads, bulk-copied B tile, TMEM alloc / st / ld, accumulate onto pre-stored TMEM, bias-by-ones-MMA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int variant, int N, int K, const float* __restrict__ A, const float* __restrict__ W,
                     const float* __restrict__ bias, const float* __restrict__ Cin, float* __restrict__ D,
                     uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem;                 // 128 x 128 bf16 max = 32768
  uint8_t* b_tile = smem + 32768;         // 128 x 128 bf16 max = 32768
  uint8_t* ones = smem + 65536;           // 4096
  uint8_t* bslab = smem + 69632;          // 2 slabs x 2048
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 73728);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 73728 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool swap_lbo_sbo = variant & 1, bias_two_slabs = variant & 2, use_bulk = variant & 4, a_in_tmem = variant & 8;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  // A: row tid
  for (int k = 0; k < K; k += 2)
    *reinterpret_cast<uint32_t*>(a_tile + tile_off(128, tid, k)) = pack_bf16(A[tid * K + k], A[tid * K + k + 1]);
  uint8_t* bdst = use_bulk ? scratch : b_tile;
  for (int i = tid; i < N * K / 2; i += 128) {
    const int n = i / (K / 2), k = (i - n * (K / 2)) * 2;
    *reinterpret_cast<uint32_t*>(bdst + tile_off(N, n, k)) = pack_bf16(W[n * K + k], W[n * K + k + 1]);
  }
  {
    *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ones + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (tid < N) {
      const float v = bias != nullptr ? bias[tid] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) << 16);
      *reinterpret_cast<uint4*>(bslab + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(bslab + N * 16 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async();
  __threadfence();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  if (use_bulk && tid == 0) {
    mbar_arrive_expect_tx(&bars[1], (uint32_t)(N * K * 2));
    bulk_g2s(b_tile, scratch, (uint32_t)(N * K * 2), &bars[1]);
  }
  if (a_in_tmem) {  // A (bf16, two k per 32-bit column) into TMEM columns [128, 128 + K/2)
    for (int cb = 0; cb < (K / 2 + 31) / 32; ++cb) {
      float h[32];
      for (int j = 0; j < 32; ++j) {
        const int k = (cb * 32 + j) * 2;
        h[j] = k + 1 < K ? __uint_as_float(pack_bf16(A[tid * K + k], A[tid * K + k + 1])) : 0.f;
      }
      tmem_st32(row_addr + 128 + cb * 32, h);
    }
    tmem_wait_st();
  }
  if (Cin != nullptr) {
    for (int cb = 0; cb < N / 32; ++cb) {
      float h[32];
      for (int k = 0; k < 32; ++k) h[k] = Cin[tid * N + cb * 32 + k];
      tmem_st32(row_addr + cb * 32, h);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (use_bulk) mbar_wait(&bars[1], 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      const uint32_t a_lbo = swap_lbo_sbo ? TILE_SBO : 2048u, a_sbo = swap_lbo_sbo ? 2048u : TILE_SBO;
      const uint32_t b_lbo = swap_lbo_sbo ? TILE_SBO : (uint32_t)(N * 16), b_sbo = swap_lbo_sbo ? (uint32_t)(N * 16) : TILE_SBO;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = make_smem_desc(smem_u32(a_tile) + ks * 4096, a_lbo, a_sbo);
        const uint64_t bd = make_smem_desc(smem_u32(b_tile) + ks * (N * 32), b_lbo, b_sbo);
        if (a_in_tmem) umma_bf16_ts(tmem, tmem + 128 + ks * 8, bd, idesc, (ks > 0 || Cin != nullptr) ? 1u : 0u);
        else umma_bf16(tmem, ad, bd, idesc, (ks > 0 || Cin != nullptr) ? 1u : 0u);
      }
      if (bias != nullptr) {
        const uint64_t od = make_smem_desc(smem_u32(ones), a_lbo, a_sbo);
        const uint64_t bd = bias_two_slabs ? make_smem_desc(smem_u32(bslab), b_lbo, b_sbo)
                                           : (swap_lbo_sbo ? make_smem_desc(smem_u32(bslab), TILE_SBO, 0u)
                                                           : make_smem_desc(smem_u32(bslab), 0u, TILE_SBO));
        umma_bf16(tmem, od, bd, idesc, 1u);
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  for (int cb = 0; cb < N / 32; ++cb) {
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int k = 0; k < 32; ++k) D[tid * N + cb * 32 + k] = h[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair UMMA self-test: D[256 x N] = Cin + A[256 x K] . W[N x K]^T + bias with cta_group::2 (M = 256).
// CTA rank r: A rows [128r, 128r+128) (shared memory, or its TMEM when variant&8), B rows [r*N/2, (r+1)*N/2) as an
// (N/2)-row tile, bias slab rows likewise; rank 1 tells rank 0 "my operands are in place" with one remote mbarrier
// arrive per warp (variant&4: B arrives by bulk copy and the full-barrier completion is relayed); rank 0 issues the
// MMAs and the multicast commit; each CTA reads its 128 rows of D out of its own TMEM.
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(int variant, int N, int K, const float* __restrict__ A, const float* __restrict__ W,
                      const float* __restrict__ bias, const float* __restrict__ Cin, float* __restrict__ D,
                      uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem;                 // 128 x 128 bf16 max = 32768
  uint8_t* b_tile = smem + 32768;         // 64 x 128 bf16 max = 16384
  uint8_t* ones = smem + 49152;           // 4096
  uint8_t* bslab = smem + 53248;          // 64 rows x 16 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 57344);  // 0: D ready (both), 1: local bulk full, 2: operands ready (rank 0; 8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 57344 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool use_bulk = variant & 4, a_in_tmem = variant & 8;
  const int NH = N / 2;
  const float* Ar = A + (size_t)rank * 128 * K;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 8); fence_barrier_init(); }
  for (int k = 0; k < K; k += 2)
    *reinterpret_cast<uint32_t*>(a_tile + tile_off(128, tid, k)) = pack_bf16(Ar[tid * K + k], Ar[tid * K + k + 1]);
  uint8_t* bdst = use_bulk ? scratch + (size_t)rank * NH * K * 2 : b_tile;
  for (int i = tid; i < NH * K / 2; i += 128) {
    const int n = i / (K / 2), k = (i - n * (K / 2)) * 2;
    const float* w = W + (size_t)(rank * NH + n) * K + k;
    *reinterpret_cast<uint32_t*>(bdst + tile_off(NH, n, k)) = pack_bf16(w[0], w[1]);
  }
  *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(ones + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < NH) {
    const float v = bias != nullptr ? bias[rank * NH + tid] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) << 16);
    *reinterpret_cast<uint4*>(bslab + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
  }
  fence_proxy_async();
  __threadfence();
  cluster_sync_all();  // barriers initialised in both CTAs before any remote arrive / multicast commit
  if (warp == 0) tmem_alloc2(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  if (use_bulk && tid == 0) {
    mbar_arrive_expect_tx(&bars[1], (uint32_t)(NH * K * 2));
    bulk_g2s(b_tile, bdst, (uint32_t)(NH * K * 2), &bars[1]);
  }
  if (a_in_tmem) {
    for (int cb = 0; cb < (K / 2 + 31) / 32; ++cb) {
      float h[32];
      for (int j = 0; j < 32; ++j) {
        const int k = (cb * 32 + j) * 2;
        h[j] = k + 1 < K ? __uint_as_float(pack_bf16(Ar[tid * K + k], Ar[tid * K + k + 1])) : 0.f;
      }
      tmem_st32(row_addr + 128 + cb * 32, h);
    }
    tmem_wait_st();
  }
  if (Cin != nullptr) {
    for (int cb = 0; cb < N / 32; ++cb) {
      float h[32];
      for (int k = 0; k < 32; ++k) h[k] = Cin[(size_t)(rank * 128 + tid) * N + cb * 32 + k];
      tmem_st32(row_addr + cb * 32, h);
    }
    tmem_wait_st();
  }
  if (use_bulk) mbar_wait(&bars[1], 0);  // every thread observes its CTA's B half before signalling
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars[2]), 0));
  if (rank == 0 && warp == 0) {
    mbar_wait_cluster(&bars[2], 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = make_smem_desc(smem_u32(a_tile) + ks * 4096, 2048, TILE_SBO);
        const uint64_t bd = make_smem_desc(smem_u32(b_tile) + ks * (NH * 32), NH * 16, TILE_SBO);
        if (a_in_tmem) umma2_bf16_ts(tmem, tmem + 128 + ks * 8, bd, idesc, (ks > 0 || Cin != nullptr) ? 1u : 0u);
        else umma2_bf16(tmem, ad, bd, idesc, (ks > 0 || Cin != nullptr) ? 1u : 0u);
      }
      if (bias != nullptr)
        umma2_bf16(tmem, make_smem_desc(smem_u32(ones), 2048, TILE_SBO), make_smem_desc(smem_u32(bslab), 0u, TILE_SBO), idesc, 1u);
      umma2_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  for (int cb = 0; cb < N / 32; ++cb) {
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int k = 0; k < 32; ++k) D[(size_t)(rank * 128 + tid) * N + cb * 32 + k] = h[k];
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// UMMA issue-rate microbenchmark: `iters` back-to-back K=16 MMAs (M=128, N) from smem operands, cycles from the
// first issue to the commit's arrival.  layout 0 = no-swizzle canonical tiles (as used by the fused kernel),
// 1 = SWIZZLE_128B descriptors.  Operand contents are irrelevant (timing only).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int layout, int N, int iters, int ksteps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 131072 + 64);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = make_idesc_bf16(128, N);
    const bool same_acc = (layout & 2) != 0;  // every MMA accumulates into the same TMEM tile (dependent chain)
    layout &= 1;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      // descriptors precomputed; the issue loop is 8 unrolled MMAs per iteration (like the fused kernel's sequences)
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int kk = ks % ksteps;
        if (layout == 0) {
          ad[ks] = make_smem_desc(sbase + kk * 4096, 2048, TILE_SBO);
          bd[ks] = make_smem_desc(sbase + 65536 + kk * (N * 32), N * 16, TILE_SBO);
        } else {  // SWIZZLE_128B K-major: rows of 128 B, 8-row atoms of 1024 B; a K=16 step advances the start by 32 B
          ad[ks] = make_smem_desc(sbase + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
          bd[ks] = make_smem_desc(sbase + 65536 + (kk >> 2) * (N * 128) + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
        }
      }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_bf16(tmem + (same_acc ? 0 : (ks & 1) * 256), ad[ks], bd[ks], idesc, 1u);
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    t1 = clock64();
    if (elect_one()) { out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// CTA-pair variant of the rate microbenchmark: cta_group::2 MMAs (M = 256, N) issued by rank 0; mode bit0: A from tensor
// memory (TS form), bit1: every MMA accumulates into the same TMEM tile, bit2: SWIZZLE_128B operand descriptors instead of
// the no-swizzle layout.  The B half tile has N/2 rows per CTA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma2_rate_kernel(int mode, int N, int iters, int ksteps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 131072 + 64);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  fence_proxy_async();
  cluster_sync_all();
  if (warp == 0) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  cluster_sync_all();
  if (warp == 0 && rank == 0) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = make_idesc_bf16(256, N);
    const bool ts = mode & 1, same_acc = mode & 2;
    const int NH = N / 2;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int kk = ks % ksteps;
        if (mode & 4) {  // SWIZZLE_128B K-major operands (rows of 128 B, 8-row atoms of 1024 B, K=16 step = +32 B)
          ad[ks] = make_smem_desc(sbase + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
          bd[ks] = make_smem_desc(sbase + 65536 + (kk >> 2) * (NH * 128) + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
        } else {
          ad[ks] = make_smem_desc(sbase + kk * 4096, 2048, TILE_SBO);
          bd[ks] = make_smem_desc(sbase + 65536 + kk * (NH * 32), NH * 16, TILE_SBO);
        }
      }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t d = tmem + (same_acc ? 0 : (ks & 1) * 256);
          if (ts) umma2_bf16_ts(d, tmem + 128 + (ks % ksteps) * 8, bd[ks], idesc, 1u);
          else umma2_bf16(d, ad[ks], bd[ks], idesc, 1u);
        }
      }
      umma2_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    t1 = clock64();
    if (elect_one()) { out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_bench_umma(int layout, int N, int iters, int ksteps, long long* out_cycles, dfb200_stream_t stream) {
  DFB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && ksteps >= 1 && ksteps <= 8, DFB200_ERR_INVALID_ARG, "bench_umma: bad shape");
  const int smem = 131072 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<1, 128, smem, as_stream(stream)>>>(layout, N, iters, ksteps, out_cycles);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_selftest_umma(int variant, int N, int K, const float* A, const float* W, const float* bias,
                                    const float* Cin, float* D, void* scratch, dfb200_stream_t stream) {
  DFB_REQUIRE((N == 32 || N == 64 || N == 128) && K >= 16 && K <= 128 && K % 16 == 0, DFB200_ERR_INVALID_ARG,
              "selftest_umma: N in {32,64,128}, K multiple of 16 in [16,128]");
  DFB_REQUIRE(!(variant & 4) || scratch != nullptr, DFB200_ERR_INVALID_ARG, "selftest_umma: bulk variant needs scratch");
  const int smem = 73728 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(variant, N, K, A, W, bias, Cin, D, reinterpret_cast<uint8_t*>(scratch));
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_selftest_umma2(int variant, int N, int K, const float* A, const float* W, const float* bias,
                                     const float* Cin, float* D, void* scratch, dfb200_stream_t stream) {
  DFB_REQUIRE((N == 32 || N == 64 || N == 128) && K >= 16 && K <= 128 && K % 16 == 0, DFB200_ERR_INVALID_ARG,
              "selftest_umma2: N in {32,64,128}, K multiple of 16 in [16,128]");
  DFB_REQUIRE(!(variant & 4) || scratch != nullptr, DFB200_ERR_INVALID_ARG, "selftest_umma2: bulk variant needs scratch");
  const int smem = 57344 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma2_selftest_kernel<<<2, 128, smem, as_stream(stream)>>>(variant, N, K, A, W, bias, Cin, D, reinterpret_cast<uint8_t*>(scratch));
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_bench_umma2(int mode, int N, int iters, int ksteps, long long* out_cycles, dfb200_stream_t stream) {
  DFB_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0 && ksteps >= 1 && ksteps <= 8, DFB200_ERR_INVALID_ARG, "bench_umma2: bad shape");
  DFB_REQUIRE(!(mode & 1) || N <= 128 || (mode & 2), DFB200_ERR_INVALID_ARG, "bench_umma2: TS mode keeps A in TMEM columns [128,192)");
  const int smem = 131072 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma2_rate_kernel<<<2, 128, smem, as_stream(stream)>>>(mode, N, iters, ksteps, out_cycles);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
#endif  // DFB200_DIAGNOSTICS
