// bf16 tensor-core GEMM for the TRAINING path (tcgen05, fp32 accumulation in TMEM): the mixed-precision counterpart of
// dfb200_sgemm with the same operand-layout flags, used by difffacto_b200/train_ops.py for the large Linear layers (forward,
// dgrad, wgrad) when the model runs with precision = "bf16" (BASELINE config 4: bf16 training).
//
//   C[M,N] = (beta ? C : 0) + bias[j] + sum_k bf16(A(i,k)) * bf16(B(k,j))        fp32 in, fp32 out
//
// One CTA (256 threads) owns a 128 x 128 output tile.  Operands are read as fp32 from global memory in either layout
// (k-contiguous or row-contiguous), converted to bf16 in registers and written as canonical K-major no-swizzle UMMA tiles
// (16-byte chunks, conflict-free) into a 2-stage shared-memory ring of 64-deep k-slices; one thread issues the 4 K=16 MMAs
// of a slice and commits to the slice's mbarrier.  Round 2: the operands of slice i+1 are requested into registers (64 per thread)
// before slice i's barrier, so their global-memory latency overlaps the MMAs and the other resident CTAs.  gridDim.z > 1
// splits K (wgrad reduces over the B*N token rows) with an atomicAdd epilogue.
#include <stdlib.h>
#include "common.cuh"
#include <cuda.h>
#include "tc_common.cuh"
#include "geglu_math.cuh"

namespace dfb200 {
using namespace tc;

constexpr int GT_M = 128, GT_N = 128, GT_K = 64, GT_THREADS = 256;
constexpr uint32_t GT_TILE_BYTES = GT_M * GT_K * 2;  // 16 KB per operand per stage

// Staging of rows [r0, r0+128) x k [k0, k0+64) of an operand as a bf16 UMMA tile (R = 128), in TWO phases so that all global
// loads of a k-slice (both operands) are in flight together: gt_load fetches the thread's 4 chunks of 8 k-values into
// registers, gt_store converts and writes them.  (Round 1 fetched, converted and stored chunk by chunk: 16 dependent
// global-memory round trips per slice -- ncu showed 62 % of the stall samples on the first F2FP after each load pair.)
// KC: element (r,k) = P[r*ld + k]; else P[k*ld + r].
//   KC mapping: thread -> rows rblk*16 + rsub + 8j (rsub = tid % 8), chunks kq + 4i (kq = (tid / 8) % 4): a warp-wide LDG.128
//     touches 8 rows x 128 contiguous bytes (8 cache lines; the row-per-lane mapping of round 1 touched 32), and a quarter warp
//     stores 8 consecutive rows of one chunk (conflict-free STS.128).
//   row-contiguous mapping: thread -> row tid % 128, chunks (tid / 128) * 4 + ch; lanes read consecutive rows of one k (coalesced).
// operand layouts of the register-staged kernel (template parameters: each mapping keeps its own register budget)
constexpr int GT_ROWS = 0, GT_KC = 1, GT_ROWS4 = 2;  // row-contiguous 32-bit loads, k-contiguous, row-contiguous 128-bit loads
template <int L>
__device__ __forceinline__ void gt_load(float (&v)[4][8], const float* __restrict__ P, int ld, int r0, int rows, int k0, int kend, bool vec_ok) {
  const int t = threadIdx.x;
  if (L == GT_KC) {
    const int rsub = t & 7, kq = (t >> 3) & 3, rblk = t >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = c >> 1, i = c & 1;
      const int gr = r0 + rblk * 16 + rsub + 8 * j, gk = k0 + (kq + 4 * i) * 8;
      const float* p = P + (size_t)gr * ld + gk;
      if (vec_ok && gr < rows && gk + 7 < kend) {
        // one 256-bit load per chunk (LDG.E.256, sm_100): every 32-byte sector is requested once (two 128-bit loads touched
        // each sector twice, and the L1 data pipe -- 73 % busy in the ncu capture of round 2 -- is what bounds this kernel)
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[c][0]), "=f"(v[c][1]), "=f"(v[c][2]), "=f"(v[c][3]), "=f"(v[c][4]), "=f"(v[c][5]), "=f"(v[c][6]), "=f"(v[c][7])
                     : "l"(p));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = (gr < rows && gk + e < kend) ? __ldg(p + e) : 0.f;
      }
    }
  } else {
    if (L == GT_ROWS4) {
      // 128-bit form (ld % 4 == 0, 16-byte aligned base): thread -> rows 4 (t % 32) .. + 3, k-chunk t / 32; a warp reads 512 contiguous
      // bytes of ONE k per instruction, 8 LDG.128 per operand and slice instead of 32 LDG.32.  v[i][e] = (row 4 (t % 32) + i, k 8 (t / 32) + e)
      const int rq = t & 31, kc = t >> 5, gr = r0 + 4 * rq, gk = k0 + kc * 8;
      if (r0 + GT_M <= rows && k0 + GT_K <= kend) {
        const float* p = P + (size_t)gk * ld + gr;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(p + (size_t)e * ld));
          v[0][e] = f.x; v[1][e] = f.y; v[2][e] = f.z; v[3][e] = f.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) v[i][e] = (gr + i < rows && gk + e < kend) ? __ldg(P + (size_t)(gk + e) * ld + gr + i) : 0.f;
      }
      return;
    }
    const int r = t & 127, kh = t >> 7, gr = r0 + r;
    if (r0 + GT_M <= rows && k0 + GT_K <= kend) {
      // interior slice (uniform per CTA): one pointer, 32 loads at constant multiples of ld.  The guarded form below spends ~5
      // instructions per element on bounds and addresses; ncu of the wgrad shapes showed the kernel issue-bound (52 % of the issue
      // slots busy at 3.1 TB/s) on exactly that
      const float* p = P + (size_t)(k0 + kh * 32) * ld + gr;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = __ldg(p + (size_t)(c * 8 + e) * ld);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gk = k0 + (kh * 4 + c) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = (gr < rows && gk + e < kend) ? __ldg(P + (size_t)(gk + e) * ld + gr) : 0.f;
      }
    }
  }
}
template <int L>
__device__ __forceinline__ void gt_store(uint8_t* tile, const float (&v)[4][8]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int r, kc;
    if (L == GT_KC) { r = (t >> 5) * 16 + (t & 7) + 8 * (c >> 1); kc = ((t >> 3) & 3) + 4 * (c & 1); }
    else if (L == GT_ROWS4) { r = 4 * (t & 31) + c; kc = t >> 5; }  // 128-bit load mapping (4-way bank conflict on this store, measured cheaper than the loads it saves)
    else { r = t & 127; kc = (t >> 7) * 4 + c; }
    *reinterpret_cast<uint4*>(tile + kc * (GT_M * 16) + r * 16) =
        make_uint4(pack_bf16(v[c][0], v[c][1]), pack_bf16(v[c][2], v[c][3]), pack_bf16(v[c][4], v[c][5]), pack_bf16(v[c][6], v[c][7]));
  }
}

template <int A_L, int B_L>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
               float* __restrict__ C, int ldc, const float* __restrict__ bias, int beta, int k_per_split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tiles = smem;                          // 2 stages
  uint8_t* b_tiles = smem + 2 * GT_TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * GT_TILE_BYTES);  // [2] stage free, [1] all done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4 * GT_TILE_BYTES + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int i0 = blockIdx.y * GT_M, j0 = blockIdx.x * GT_N;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = make_idesc_bf16(GT_M, GT_N);
  uint32_t ph[2] = {0, 0};
  int it = 0;
  // 32-byte loads need k-contiguous rows that start on 32-byte boundaries (uniform per launch)
  const bool a_vec = (A_L == GT_KC) && (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 31) == 0) && (kbeg % 8 == 0);
  const bool b_vec = (B_L == GT_KC) && (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(B) & 31) == 0) && (kbeg % 8 == 0);
  float va[4][8], vb[4][8];
  if (kbeg < kend) {
    gt_load<A_L>(va, A, lda, i0, M, kbeg, kend, a_vec);
    gt_load<B_L>(vb, B, ldb, j0, N, kbeg, kend, b_vec);
  }
  for (int k0 = kbeg; k0 < kend; k0 += GT_K, ++it) {
    const int s = it & 1;
    if (it >= 2) {  // the MMAs that read this stage two slices ago have completed
      mbar_wait(&bars[s], ph[s]);
      ph[s] ^= 1;
    }
    gt_store<A_L>(a_tiles + s * GT_TILE_BYTES, va);
    gt_store<B_L>(b_tiles + s * GT_TILE_BYTES, vb);
    if (k0 + GT_K < kend) {  // the next slice's operands fly while this slice's MMAs run
      gt_load<A_L>(va, A, lda, i0, M, k0 + GT_K, kend, a_vec);
      gt_load<B_L>(vb, B, ldb, j0, N, k0 + GT_K, kend, b_vec);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(a_tiles + s * GT_TILE_BYTES), bb = smem_u32(b_tiles + s * GT_TILE_BYTES);
#pragma unroll
      for (int ks = 0; ks < GT_K / 16; ++ks)
        umma_bf16(tmem, make_smem_desc(aa + ks * 2 * (GT_M * 16), GT_M * 16, TILE_SBO),
                  make_smem_desc(bb + ks * 2 * (GT_N * 16), GT_N * 16, TILE_SBO), idesc, (it > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&bars[s]);
    }
  }
  if (tid == 0) umma_commit(&bars[2]);
  __syncwarp();
  if (it > 0) mbar_wait(&bars[2], 0);
  tc_fence_after();
  // epilogue: warps w and w+4 share TMEM lanes 32*(w%4)..; warps 0-3 take columns [0,64), warps 4-7 [64,128).  A thread
  // holds one ROW of the accumulator; the 32x32 block of a warp is transposed through shared memory (the operand stages are
  // free now) so that every global store instruction writes 128 contiguous bytes of one row of C.
  {
    const int lane = tid & 31;
    float* tw = reinterpret_cast<float*>(smem) + warp * (32 * 33);  // [32 rows][33]
    const int rbase = i0 + (warp & 3) * 32;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float h[32];
      if (it > 0) {
        tmem_ld32(taddr + cb * 32, h);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) h[e] = 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 32; ++e) tw[lane * 33 + e] = h[e];
      __syncwarp();
      const int j = j0 + (warp >> 2) * 64 + cb * 32 + lane;
      const float bj = (bias != nullptr && blockIdx.z == 0 && j < N) ? __ldg(bias + j) : 0.f;
      if (j < N) {
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int i = rbase + rr;
          if (i >= M) break;
          const float v = tw[rr * 33 + lane] + bj;
          float* o = C + (size_t)i * ldc + j;
          if (split) atomicAdd(o, v);
          else *o = beta ? *o + v : v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Round 2: ASYNCHRONOUS-COPY form for k-contiguous operands (forward GEMMs; dgrad with a pre-transposed weight).  The kernel above
// stages fp32 operands through REGISTERS (convert to bf16, store): one 64-deep k-slice = 64 KB of loads per CTA is all that can be
// in flight, and the training GEMMs ran 2.5-4x off their memory floor for it.  Here nothing passes through registers:
//   * the tensor cores take the fp32 containers as they are (tcgen05.mma kind::tf32: the top 19 bits of each value, fp32
//     accumulation -- 10 mantissa bits instead of the 8 of the bf16 staging; the MMA runs at a third of the bf16 rate, which does not
//     matter for GEMMs that are memory bound by a wide margin),
//   * 4 producer warps copy 128 x 32 fp32 slices global -> shared with 16-byte cp.async (LDGSTS) straight into the SWIZZLE_128B
//     K-major UMMA layout (a row of a slice = 128 contiguous bytes, the 16-byte chunk index XORed with row % 8: coalesced global
//     reads, conflict-free shared-memory writes), 3 slices in flight per thread in a 5-stage ring,
//   * one warp issues the MMAs (4 x K=8 per slice), 4 epilogue warps drain two double-buffered TMEM accumulators through a
//     shared-memory transpose into 128-bit global stores while the next tile's MMAs run; CTAs are persistent over (tile, k-split)
//     work items, n fastest, so A row blocks are shared through L2.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int GA_STAGES = 6, GA_K = 32, GA_PROD = 128, GA_THREADS = GA_PROD + 5 * 32, GA_INFLIGHT = 5;
constexpr uint32_t GA_TILE = 128 * GA_K * 4;                 // 16 KB per operand and slice
constexpr uint32_t GA_STAGE = 2 * GA_TILE;
constexpr uint32_t GA_TW_STRIDE = 36;
constexpr uint32_t GA_SM_TW = GA_STAGES * GA_STAGE;
constexpr uint32_t GA_SM_BAR = GA_SM_TW + 4 * 32 * GA_TW_STRIDE * 4;
constexpr uint32_t GA_SMEM = GA_SM_BAR + 256;
enum GaBar { GA_FULL = 0, GA_EMPTY = GA_STAGES, GA_ACC_FULL = 2 * GA_STAGES, GA_ACC_EMPTY = 2 * GA_STAGES + 2 };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// copy rows [r0, r0+128) x k [k0, k0+32) of a k-contiguous fp32 operand into a SWIZZLE_128B tile (zero fill outside rows / kend)
__device__ __forceinline__ void ga_copy(uint32_t tile, const float* __restrict__ P, int ld, int r0, int rows, int k0, int kend) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 16 + (t >> 3), c = t & 7;
    const int gr = r0 + row, gk = k0 + c * 4;
    int bytes = 0;
    if (gr < rows && gk < kend) bytes = min(4, kend - gk) * 4;
    const float* src = P + (size_t)(gr < rows ? gr : 0) * ld + (gk < kend ? gk : 0);
    cp_async16(tile + row * 128 + ((c ^ (row & 7)) << 4), src, (uint32_t)bytes);
  }
}

// interior form of ga_copy (all 128 rows and all 32 k inside the operand): the per-thread source pointer and the swizzled destination
// are loop invariants of the caller, a copy is one LDGSTS plus one 64-bit add.  The general form recomputes row / k bounds, the
// zero-fill size and the address for each of its 8 copies (~230 dependent integer instructions per slice and warp), which made the
// 4 producer warps -- not the memory system -- the second bottleneck of the memory-bound shapes.
__device__ __forceinline__ void ga_copy_interior(uint32_t dst, const float* __restrict__ src, size_t row_step) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 2048), "l"(src + i * row_step) : "memory");
}

__global__ void __launch_bounds__(GA_THREADS, 1)
gemm_tf32_async_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
                       int ldc, const float* __restrict__ bias, int beta, int k_per_split, int tiles_n, int tiles_mn, int items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GA_SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GA_SM_BAR + 192);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  constexpr int W_MMA = GA_PROD / 32;
  if (tid == 0) {
    for (int i = 0; i < GA_STAGES; ++i) { mbar_init(&bars[GA_FULL + i], GA_PROD); mbar_init(&bars[GA_EMPTY + i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[GA_ACC_FULL + i], 1); mbar_init(&bars[GA_ACC_EMPTY + i], 128); }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool split = items > tiles_mn;
  const uint32_t sbase = smem_u32(smem);

  if (warp < W_MMA) {
    // =========================== producers ===========================
    int it = 0;       // slices issued
    int done = 0;     // slices published
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int z = item / tiles_mn, mn = item - z * tiles_mn;
      const int i0 = (mn / tiles_n) * GT_M, j0 = (mn % tiles_n) * GT_N;
      const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
      const bool interior = i0 + GT_M <= M && j0 + GT_N <= N && (ke - kb) % GA_K == 0;
      // thread t copies 16-byte chunk t % 8 of rows t / 8 + 16 i; (row & 7) does not depend on i, so the swizzle is a thread constant
      const uint32_t dst0 = (uint32_t)(tid >> 3) * 128u + (uint32_t)(((tid & 7) ^ ((tid >> 3) & 7)) << 4);
      const float* pa = A + (size_t)(i0 + (tid >> 3)) * lda + (tid & 7) * 4 + kb;
      const float* pb = B + (size_t)(j0 + (tid >> 3)) * ldb + (tid & 7) * 4 + kb;
#pragma unroll 1
      for (int k0 = kb; k0 < ke; k0 += GA_K, ++it, pa += GA_K, pb += GA_K) {
        const int s = it % GA_STAGES;
        mbar_wait(&bars[GA_EMPTY + s], (((uint32_t)(it / GA_STAGES)) & 1u) ^ 1u);
        const uint32_t st = sbase + s * GA_STAGE;
        if (interior) {
          ga_copy_interior(st + dst0, pa, (size_t)16 * lda);
          ga_copy_interior(st + GA_TILE + dst0, pb, (size_t)16 * ldb);
        } else {
          ga_copy(st, A, lda, i0, M, k0, ke);
          ga_copy(st + GA_TILE, B, ldb, j0, N, k0, ke);
        }
        cp_async_commit();
        if (it - done >= GA_INFLIGHT - 1) {  // the slice issued GA_INFLIGHT - 1 iterations ago has landed: publish it
          cp_async_wait<GA_INFLIGHT - 1>();
          fence_proxy_async();
          mbar_arrive(&bars[GA_FULL + done % GA_STAGES]);
          ++done;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; done < it; ++done) mbar_arrive(&bars[GA_FULL + done % GA_STAGES]);
  } else if (warp == W_MMA) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_tf32(GT_M, GT_N);
    int it = 0, t = 0;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
      const int z = item / tiles_mn;
      const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
      const int b = t & 1;
      mbar_wait(&bars[GA_ACC_EMPTY + b], (((uint32_t)(t >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      bool first = true;
#pragma unroll 1
      for (int k0 = kb; k0 < ke; k0 += GA_K, ++it) {
        const int s = it % GA_STAGES;
        mbar_wait(&bars[GA_FULL + s], ((uint32_t)(it / GA_STAGES)) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aa = sbase + s * GA_STAGE, bb = aa + GA_TILE;
#pragma unroll
          for (int ks = 0; ks < GA_K / 8; ++ks)  // SWIZZLE_128B K-major: 8-row atoms of 1024 B; a K = 8 step (32 B) advances the start address
            umma_tf32(tmem + b * 128, make_smem_desc(aa + ks * 32, 16, 1024) | ((uint64_t)2 << 61),
                      make_smem_desc(bb + ks * 32, 16, 1024) | ((uint64_t)2 << 61), idesc, (!first || ks > 0) ? 1u : 0u);
          umma_commit(&bars[GA_EMPTY + s]);
        }
        __syncwarp();
        first = false;
      }
      if (elect_one()) umma_commit(&bars[GA_ACC_FULL + b]);
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // =========================== epilogue (4 warps; TMEM lanes 32 * (warp % 4) ..) ===========================
    const int q = warp & 3;
    float* tw = reinterpret_cast<float*>(smem + GA_SM_TW) + (warp - W_MMA - 1) * (32 * GA_TW_STRIDE);
    const bool vec_c = (N % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    int t = 0;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
      const int z = item / tiles_mn, mn = item - z * tiles_mn;
      const int i0 = (mn / tiles_n) * GT_M, j0 = (mn % tiles_n) * GT_N;
      const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
      const int b = t & 1;
      const bool interior = vec_c && !split && i0 + GT_M <= M && j0 + GT_N <= N;
      // bias of the 4 column blocks, requested BEFORE the wait for the accumulator (a global load per column block on the critical
      // path of each block cost ~4 x 700 cycles per tile)
      float4 bj4[4];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        bj4[cb] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (interior && bias != nullptr) bj4[cb] = __ldg(reinterpret_cast<const float4*>(bias + j0 + cb * 32 + 4 * (lane & 7)));
      }
      mbar_wait(&bars[GA_ACC_FULL + b], ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + b * 128;
      const int rbase = i0 + q * 32;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float h[32];
        if (kb < ke) {
          tmem_ld32(taddr + cb * 32, h);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) h[e] = 0.f;
        }
        if (cb == 3) {  // the accumulator is in registers: hand the buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&bars[GA_ACC_EMPTY + b]);
        }
        __syncwarp();
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4)
          *reinterpret_cast<float4*>(tw + lane * GA_TW_STRIDE + 4 * e4) = make_float4(h[4 * e4], h[4 * e4 + 1], h[4 * e4 + 2], h[4 * e4 + 3]);
        __syncwarp();
        const int jb = j0 + cb * 32;
        if (interior) {
          // interior tile: branch-free, all 8 shared-memory reads (and the 8 loads of C for beta) in flight before the first
          // store.  The general path below runs one LDS -> FADD -> STG chain at a time behind its bounds checks; with 4 epilogue
          // warps that chain, not the copies or the MMAs, set the pace of the memory-bound shapes (ncu: MMA warp waiting on
          // ACC_EMPTY, producers on EMPTY, epilogue warps never idle)
          const int j = jb + 4 * (lane & 7);
          const float4 bj = bj4[cb];
          float* o = C + (size_t)(rbase + (lane >> 3)) * ldc + j;
          const float* ts = tw + (lane >> 3) * GA_TW_STRIDE + 4 * (lane & 7);
          float4 v[8];
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4) {
            v[r4] = *reinterpret_cast<const float4*>(ts + r4 * 4 * GA_TW_STRIDE);
            v[r4].x += bj.x; v[r4].y += bj.y; v[r4].z += bj.z; v[r4].w += bj.w;  // same order as the general path: (acc + bias) + C
          }
          if (beta) {
            float4 c0[8];
#pragma unroll
            for (int r4 = 0; r4 < 8; ++r4) c0[r4] = *reinterpret_cast<const float4*>(o + (size_t)r4 * 4 * ldc);
#pragma unroll
            for (int r4 = 0; r4 < 8; ++r4) { v[r4].x += c0[r4].x; v[r4].y += c0[r4].y; v[r4].z += c0[r4].z; v[r4].w += c0[r4].w; }
          }
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4) *reinterpret_cast<float4*>(o + (size_t)r4 * 4 * ldc) = v[r4];
        } else if (vec_c) {
          const int j = jb + 4 * (lane & 7);
          float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias != nullptr && z == 0 && j < N) bj = __ldg(reinterpret_cast<const float4*>(bias + j));
#pragma unroll 2
          for (int r4 = 0; r4 < 8; ++r4) {
            const int rr = r4 * 4 + (lane >> 3), i = rbase + rr;
            if (i < M && j < N) {
              float4 v = *reinterpret_cast<const float4*>(tw + rr * GA_TW_STRIDE + 4 * (lane & 7));
              v.x += bj.x; v.y += bj.y; v.z += bj.z; v.w += bj.w;
              float* o = C + (size_t)i * ldc + j;
              if (split) { atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w); }
              else {
                if (beta) { const float4 c0 = *reinterpret_cast<const float4*>(o); v.x += c0.x; v.y += c0.y; v.z += c0.z; v.w += c0.w; }
                *reinterpret_cast<float4*>(o) = v;
              }
            }
          }
        } else {
          const int j = jb + lane;
          const float bj = (bias != nullptr && z == 0 && j < N) ? __ldg(bias + j) : 0.f;
          if (j < N) {
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              const int i = rbase + rr;
              if (i >= M) break;
              const float v = tw[rr * GA_TW_STRIDE + lane] + bj;
              float* o = C + (size_t)i * ldc + j;
              if (split) atomicAdd(o, v);
              else *o = beta ? *o + v : v;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// TMA form of the kernel above (the default for k-contiguous operands).  ncu of the cp.async form on the memory-bound training
// shapes (32 768 x 1 024 x 128: 42 us against a 21 us HBM floor) showed producers and epilogue warps both > 80 % busy while DRAM,
// L2 and the tensor pipe idled: every operand byte went through the LSU as a 16-byte LDGSTS, and every result byte through it
// three times (STS, LDS, STG).  Here
//   * ONE thread issues two cp.async.bulk.tensor.2d loads per k-slice (box 32 k x 128 rows, SWIZZLE_128B tensor maps built on the
//     host per call; rows / k outside the operand are zero-filled by the TMA unit, so there is no edge form), completion by
//     mbarrier complete_tx: all GA2_STAGES slices can be in flight and the LSU sees none of it;
//   * 8 epilogue warps (two per TMEM lane quadrant, two 32-column blocks each) add the bias in registers, stage their 32 x 32
//     block row-wise in shared memory and every lane sends ITS row to global memory with one 128-byte cp.async.bulk -- or
//     cp.reduce.async.bulk .add.f32 for beta = 1 (C is never read into the SM) and for split-K partial sums (no scalar atomics).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int GA2_STAGES = 5, GA2_EPI = 8, GA2_THREADS = (2 + GA2_EPI) * 32;
constexpr int GA2_EPI_GEGLU = 16, GA2_THREADS_GEGLU = (2 + GA2_EPI_GEGLU) * 32;  // GEGLU epilogue: 16 warps x 16 hidden units (latency-bound math)
constexpr uint32_t GA2_SM_TW = GA2_STAGES * GA_STAGE;
constexpr uint32_t GA2_TW_BYTES = 32 * 128;  // one 32 x 32 fp32 block, SWIZZLE_128B rows
constexpr uint32_t GA2_SM_BAR = GA2_SM_TW + GA2_EPI * 2 * GA2_TW_BYTES;
constexpr uint32_t GA2_SMEM = GA2_SM_BAR + 256;
enum Ga2Bar { GA2_FULL = 0, GA2_EMPTY = GA2_STAGES, GA2_ACC_FULL = 2 * GA2_STAGES, GA2_ACC_EMPTY = 2 * GA2_STAGES + 2 };

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
               "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// GEGLU = true: FeedForward's first half in one kernel (attention.py:77-94: Linear(128, 2H) -> GEGLU -> Dropout).  A tile covers 64
// hidden units: the B operand is TWO 64-row boxes of W1 (value rows n0.., gate rows H + n0..), so the accumulator holds the value
// columns in [0, 64) and their gates in [64, 128); an epilogue warp reads a 32-column value block and its gate block, stores both
// to h (kept for the backward) and u = dropout(value * gelu(gate)) to a second output -- the (M, 2H) pre-activation is not read again.
struct GegluArgs {
  int H;
  float p, scale;
  unsigned long long seed, offset;
  const unsigned long long* step;
};
template <bool GEGLU>
__global__ void __launch_bounds__(GEGLU ? GA2_THREADS_GEGLU : GA2_THREADS, 1)
gemm_tf32_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_u, const GegluArgs G, int M, int N,
                     int K, const float* __restrict__ bias, int beta, int k_per_split, int tiles_n, int tiles_mn,
                     int items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GA2_SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GA2_SM_BAR + 192);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  if (tid == 0) {
    for (int i = 0; i < GA2_STAGES; ++i) { mbar_init(&bars[GA2_FULL + i], 1); mbar_init(&bars[GA2_EMPTY + i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[GA2_ACC_FULL + i], 1); mbar_init(&bars[GA2_ACC_EMPTY + i], (GEGLU ? GA2_EPI_GEGLU : GA2_EPI) * 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool split = items > tiles_mn;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
    // =========================== TMA producer (one lane) ===========================
    if (lane == 0) {
      int it = 0;
#pragma unroll 1
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int z = item / tiles_mn, mn = item - z * tiles_mn;
        const int i0 = (mn / tiles_n) * GT_M, j0 = (mn % tiles_n) * GT_N;
        const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
#pragma unroll 1
        for (int k0 = kb; k0 < ke; k0 += GA_K, ++it) {
          const int s = it % GA2_STAGES;
          mbar_wait(&bars[GA2_EMPTY + s], (((uint32_t)(it / GA2_STAGES)) & 1u) ^ 1u);
          const uint32_t st = sbase + s * GA_STAGE;
          mbar_arrive_expect_tx(&bars[GA2_FULL + s], GA_STAGE);
          tma_load_2d(st, &map_a, k0, i0, &bars[GA2_FULL + s]);
          if constexpr (GEGLU) {
            const int n0 = (mn % tiles_n) * 64;
            tma_load_2d(st + GA_TILE, &map_b, k0, n0, &bars[GA2_FULL + s]);
            tma_load_2d(st + GA_TILE + 64 * 128, &map_b, k0, G.H + n0, &bars[GA2_FULL + s]);
          } else {
            tma_load_2d(st + GA_TILE, &map_b, k0, j0, &bars[GA2_FULL + s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_tf32(GT_M, GT_N);
    int it = 0, t = 0;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
      const int z = item / tiles_mn;
      const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
      const int b = t & 1;
      mbar_wait(&bars[GA2_ACC_EMPTY + b], (((uint32_t)(t >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      bool first = true;
#pragma unroll 1
      for (int k0 = kb; k0 < ke; k0 += GA_K, ++it) {
        const int s = it % GA2_STAGES;
        mbar_wait(&bars[GA2_FULL + s], ((uint32_t)(it / GA2_STAGES)) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aa = sbase + s * GA_STAGE, bb = aa + GA_TILE;
#pragma unroll
          for (int ks = 0; ks < GA_K / 8; ++ks)
            umma_tf32(tmem + b * 128, make_smem_desc(aa + ks * 32, 16, 1024) | ((uint64_t)2 << 61),
                      make_smem_desc(bb + ks * 32, 16, 1024) | ((uint64_t)2 << 61), idesc, (!first || ks > 0) ? 1u : 0u);
          umma_commit(&bars[GA2_EMPTY + s]);
        }
        __syncwarp();
        first = false;
      }
      if (elect_one()) umma_commit(&bars[GA2_ACC_FULL + b]);
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // =========================== epilogue: 8 warps, TMEM lanes 32 (warp % 4).., column blocks 2 half, 2 half + 1 ===========================
    const int q = warp & 3, half = (warp - 2) >> 2;
    // two staging blocks per warp: 32 x 32 fp32 (4 KB, SWIZZLE_128B) or, GEGLU form, 32 x 16 (2 KB, SWIZZLE_64B); 64 KB in all either way
    uint8_t* tw = smem + GA2_SM_TW + (warp - 2) * (GEGLU ? GA2_TW_BYTES : 2 * GA2_TW_BYTES);
    int t = 0, nst = 0;
    auto stage_store16 = [&](const float (&h)[16], const CUtensorMap* map, int col, int row, bool issue) {
      uint8_t* buf = tw + (nst & 1) * (GA2_TW_BYTES / 2);
      ++nst;
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4)  // SWIZZLE_64B: chunk e4 of row `lane` at position e4 ^ ((row / 2) % 4)
        *reinterpret_cast<float4*>(buf + lane * 64 + ((e4 ^ ((lane >> 1) & 3)) << 4)) = make_float4(h[4 * e4], h[4 * e4 + 1], h[4 * e4 + 2], h[4 * e4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (issue) tma_store_2d(map, col, row, smem_u32(buf));
        bulk_commit();
      }
    };
    // one 32 x 32 block (row = lane) -> staging -> global: rows >= M / columns outside the tensor are clipped by the TMA unit
    auto stage_store = [&](const float (&h)[32], const CUtensorMap* map, int col, int row, bool issue, bool reduce) {
      uint8_t* buf = tw + (nst & 1) * GA2_TW_BYTES;
      ++nst;
      if (lane == 0) bulk_wait_read<1>();  // the store issued two blocks ago has read this buffer
      __syncwarp();
#pragma unroll
      for (int e4 = 0; e4 < 8; ++e4)  // 16-byte chunk e4 of row `lane` at position e4 ^ (row % 8): conflict-free, and what the tensor map expects
        *reinterpret_cast<float4*>(buf + lane * 128 + ((e4 ^ (lane & 7)) << 4)) = make_float4(h[4 * e4], h[4 * e4 + 1], h[4 * e4 + 2], h[4 * e4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (issue) {
          if (reduce) tma_reduce_add_2d(map, col, row, smem_u32(buf));
          else tma_store_2d(map, col, row, smem_u32(buf));
        }
        bulk_commit();
      }
    };
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
      const int z = item / tiles_mn, mn = item - z * tiles_mn;
      const int i0 = (mn / tiles_n) * GT_M;
      const int kb = z * k_per_split, ke = min(K, kb + k_per_split);
      const int b = t & 1;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + b * 128;
      const int rbase = i0 + q * 32;
      if constexpr (GEGLU) {
        // warp = (quadrant q, slice): 16 hidden units; the Philox + normal-CDF math of the GEGLU is a long dependent chain per element,
        // so the block is kept small and 16 warps (4 per scheduler) hide each other's latencies (8 warps x 32 units: 67 us per
        // 32 768 x 1 024 x 128 launch against 38 us for the plain GEMM)
        const int slice = (warp - 2) >> 2;
        const int hcol = (mn % tiles_n) * 64 + 16 * slice;  // first hidden unit of this warp's block
        float bb = 0.f;  // lanes 0-15: value bias of unit hcol + lane, lanes 16-31: gate bias of unit hcol + lane - 16
        if (bias != nullptr) bb = __ldg(bias + (lane < 16 ? hcol + lane : G.H + hcol + lane - 16));
        mbar_wait(&bars[GA2_ACC_FULL + b], ((uint32_t)(t >> 1)) & 1u);
        tc_fence_after();
        float v[16], g[16];
        tmem_ld16(taddr + 16 * slice, v);
        tmem_ld16(taddr + 64 + 16 * slice, g);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&bars[GA2_ACC_EMPTY + b]);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          v[e] += __shfl_sync(0xFFFFFFFFu, bb, e);
          g[e] += __shfl_sync(0xFFFFFFFFu, bb, 16 + e);
        }
        const bool in = rbase < M;
        stage_store16(v, &map_c, hcol, rbase, in);
        stage_store16(g, &map_c, G.H + hcol, rbase, in);
        unsigned long long seed = G.seed;
        if (G.p > 0.f && G.step != nullptr) seed += __ldg(G.step) * 0x9E3779B97F4A7C15ull;
        const long long quad0 = (long long)(rbase + lane) * (G.H >> 2) + (hcol >> 2);
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          float m[4] = {1.f, 1.f, 1.f, 1.f};
          if (G.p > 0.f) dropout_keep4(G.p, G.scale, seed, G.offset, quad0 + e4, m);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float cdf, pdf;
            normal_cdf_pdf(g[4 * e4 + i], cdf, pdf);
            v[4 * e4 + i] = m[i] * v[4 * e4 + i] * (g[4 * e4 + i] * cdf);  // same expression as geglu_dropout_fwd_kernel
          }
        }
        stage_store16(v, &map_u, hcol, rbase, in);
      } else {
        const int j0 = (mn % tiles_n) * GT_N;
        // bias of this warp's two column blocks: lane l holds column l; requested before the wait for the accumulator
        float bl[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = j0 + (2 * half + c) * 32 + lane;
          bl[c] = (bias != nullptr && z == 0 && j < N) ? __ldg(bias + j) : 0.f;
        }
        mbar_wait(&bars[GA2_ACC_FULL + b], ((uint32_t)(t >> 1)) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cb = 2 * half + c;
          float h[32];
          if (kb < ke) {
            tmem_ld32(taddr + cb * 32, h);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) h[e] = 0.f;
          }
          if (c == 1) {  // this warp's part of the accumulator is in registers
            tc_fence_before();
            mbar_arrive(&bars[GA2_ACC_EMPTY + b]);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) h[e] += __shfl_sync(0xFFFFFFFFu, bl[c], e);
          const int jb = j0 + cb * 32;
          stage_store(h, &map_c, jb, rbase, rbase < M && jb < N, split || beta);
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on a driver symbol)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// k-contiguous fp32 operand (rows x K, leading dimension ld) as a 2-D tensor map with a 32 k x 128 row box, SWIZZLE_128B
static bool make_operand_map(CUtensorMap* map, const float* P, int rows, int K, int ld, int box_rows = 128, int box_cols = GA_K) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};  // inner extent 128 bytes -> SWIZZLE_128B, 64 bytes -> SWIZZLE_64B
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(P), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_gemm_bf16(int a_k_contiguous, int b_k_contiguous, int M, int N, int K, const float* A, int lda, const float* B,
                                int ldb, float* C, int ldc, const float* bias, int beta, int split_k, dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && split_k >= 1, DFB200_ERR_INVALID_ARG, "gemm_bf16: bad sizes M=%d N=%d K=%d split=%d", M, N, K, split_k);
  if (M == 0 || N == 0) return DFB200_OK;
  int kps = cdiv(cdiv(K, split_k), GT_K) * GT_K;
  if (kps == 0) kps = GT_K;
  const int splits = K == 0 ? 1 : cdiv(K, kps);
  // asynchronous-copy tf32 kernel for k-contiguous, 16-byte aligned operands (DFB200_GEMM_ASYNC=0 keeps the register-staged
  // bf16 kernel for A/B measurements)
  static const bool use_async = [] { const char* e = getenv("DFB200_GEMM_ASYNC"); return e == nullptr || e[0] != '0'; }();
  if (use_async && a_k_contiguous && b_k_contiguous && K > 0 && lda % 4 == 0 && ldb % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0) {
    const int kps_a = cdiv(cdiv(K, split_k), GA_K) * GA_K;
    const int splits_a = cdiv(K, kps_a);
    const int tiles_n = cdiv(N, GT_N), tiles_m = cdiv(M, GT_M);
    const long long items = (long long)tiles_n * tiles_m * splits_a;
    DFB_REQUIRE(items <= 0x7fffffffLL, DFB200_ERR_INVALID_ARG, "gemm_bf16: too many tiles");
    const int n_sm = current_device_sm_count();
    DFB_REQUIRE(n_sm > 0, DFB200_ERR_CUDA, "gemm_bf16: cannot query the SM count of the current device");
    // DFB200_GEMM_ASYNC=cp keeps the cp.async form (A/B measurements)
    static const bool use_tma = [] { const char* e = getenv("DFB200_GEMM_ASYNC"); return e == nullptr || e[0] != 'c'; }();
    if (use_tma && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
      CUtensorMap map_a, map_b, map_c;  // C: 32 x 32 boxes (the epilogue's staging blocks)
      DFB_REQUIRE(make_operand_map(&map_a, A, M, K, lda) && make_operand_map(&map_b, B, N, K, ldb) && make_operand_map(&map_c, C, M, N, ldc, 32),
                  DFB200_ERR_CUDA,
                  "gemm_bf16: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d ldb=%d)", M, N, K, lda, ldb);
      static DeviceOnce once2;
      if (once2.first_time())
        DFB_CUDA(cudaFuncSetAttribute(gemm_tf32_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GA2_SMEM));
      gemm_tf32_tma_kernel<false><<<(int)(items < n_sm ? items : n_sm), GA2_THREADS, GA2_SMEM, as_stream(stream)>>>(
          map_a, map_b, map_c, map_c, GegluArgs{}, M, N, K, bias, beta, kps_a, tiles_n, tiles_n * tiles_m, (int)items);
      DFB_LAUNCH_CHECK();
      return DFB200_OK;
    }
    static DeviceOnce once;
    if (once.first_time()) DFB_CUDA(cudaFuncSetAttribute(gemm_tf32_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GA_SMEM));
    gemm_tf32_async_kernel<<<(int)(items < n_sm ? items : n_sm), GA_THREADS, GA_SMEM, as_stream(stream)>>>(
        M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps_a, tiles_n, tiles_n * tiles_m, (int)items);
    DFB_LAUNCH_CHECK();
    return DFB200_OK;
  }
  dim3 grid(cdiv(N, GT_N), cdiv(M, GT_M), splits);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "gemm_bf16: grid too large");
  const int smem = 4 * GT_TILE_BYTES + 128;
  cudaStream_t st = as_stream(stream);
#define GT_LAUNCH(AK, BK)                                                                                                      \
  do {                                                                                                                         \
    static DeviceOnce once; /* cudaFuncSetAttribute is per device */                                                          \
    if (once.first_time()) DFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    gemm_tc_kernel<AK, BK><<<grid, GT_THREADS, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps);                  \
  } while (0)
  // row-contiguous operands take 128-bit loads when their rows start on 16-byte boundaries
  const int la = a_k_contiguous ? GT_KC : (lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) ? GT_ROWS4 : GT_ROWS;
  const int lb = b_k_contiguous ? GT_KC : (ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0) ? GT_ROWS4 : GT_ROWS;
#define GT_ROW(LA)                                  \
  if (lb == GT_KC) GT_LAUNCH(LA, GT_KC);            \
  else if (lb == GT_ROWS4) GT_LAUNCH(LA, GT_ROWS4); \
  else GT_LAUNCH(LA, GT_ROWS)
  if (la == GT_KC) { GT_ROW(GT_KC); }
  else if (la == GT_ROWS4) { GT_ROW(GT_ROWS4); }
  else { GT_ROW(GT_ROWS); }
#undef GT_ROW
#undef GT_LAUNCH
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_ff_in_forward(long long M, int H, int K, const float* x, int ldx, const float* w1, int ldw, const float* b1, float p,
                                    uint64_t seed, uint64_t offset, const unsigned long long* step, float* h, float* u,
                                    dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 0 && M <= 0x7fffff00LL && H > 0 && K > 0, DFB200_ERR_INVALID_ARG, "ff_in_forward: bad sizes M=%lld H=%d K=%d", M, H, K);
  DFB_REQUIRE(p >= 0.f && p < 1.f, DFB200_ERR_INVALID_ARG, "ff_in_forward: dropout p must be in [0, 1) (got %g)", (double)p);
  if (M == 0) return DFB200_OK;
  DFB_REQUIRE(H % 64 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ldx >= K && ldw >= K &&
                  ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(u)) & 15) == 0,
              DFB200_ERR_UNSUPPORTED, "ff_in_forward: needs H %% 64 == 0, leading dimensions that are multiples of 4 and 16-byte aligned pointers (H=%d)", H);
  CUtensorMap map_a, map_b, map_c, map_u;
  DFB_REQUIRE(make_operand_map(&map_a, x, (int)M, K, ldx) && make_operand_map(&map_b, w1, 2 * H, K, ldw, 64) &&
                  make_operand_map(&map_c, h, (int)M, 2 * H, 2 * H, 32, 16) && make_operand_map(&map_u, u, (int)M, H, H, 32, 16),
              DFB200_ERR_CUDA, "ff_in_forward: cuTensorMapEncodeTiled failed (M=%lld H=%d K=%d)", M, H, K);
  const int tiles_n = H / 64, tiles_m = cdiv((int)M, GT_M);
  const long long items = (long long)tiles_n * tiles_m;
  DFB_REQUIRE(items <= 0x7fffffffLL, DFB200_ERR_INVALID_ARG, "ff_in_forward: too many tiles");
  const int n_sm = current_device_sm_count();
  DFB_REQUIRE(n_sm > 0, DFB200_ERR_CUDA, "ff_in_forward: cannot query the SM count of the current device");
  static DeviceOnce once;
  if (once.first_time()) DFB_CUDA(cudaFuncSetAttribute(gemm_tf32_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GA2_SMEM));
  const GegluArgs G{H, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, offset, step};
  gemm_tf32_tma_kernel<true><<<(int)(items < n_sm ? items : n_sm), GA2_THREADS_GEGLU, GA2_SMEM, as_stream(stream)>>>(
      map_a, map_b, map_c, map_u, G, (int)M, 2 * H, K, b1, 0, cdiv(K, GA_K) * GA_K, tiles_n, tiles_n * tiles_m, (int)items);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
