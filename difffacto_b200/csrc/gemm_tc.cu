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
#include "common.cuh"
#include "tc_common.cuh"

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
template <bool KC>
__device__ __forceinline__ void gt_load(float (&v)[4][8], const float* __restrict__ P, int ld, int r0, int rows, int k0, int kend, bool vec_ok) {
  const int t = threadIdx.x;
  if (KC) {
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
    const int r = t & 127, kh = t >> 7, gr = r0 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gk = k0 + (kh * 4 + c) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[c][e] = (gr < rows && gk + e < kend) ? __ldg(P + (size_t)(gk + e) * ld + gr) : 0.f;
    }
  }
}
template <bool KC>
__device__ __forceinline__ void gt_store(uint8_t* tile, const float (&v)[4][8]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int r, kc;
    if (KC) { r = (t >> 5) * 16 + (t & 7) + 8 * (c >> 1); kc = ((t >> 3) & 3) + 4 * (c & 1); }
    else { r = t & 127; kc = (t >> 7) * 4 + c; }
    *reinterpret_cast<uint4*>(tile + kc * (GT_M * 16) + r * 16) =
        make_uint4(pack_bf16(v[c][0], v[c][1]), pack_bf16(v[c][2], v[c][3]), pack_bf16(v[c][4], v[c][5]), pack_bf16(v[c][6], v[c][7]));
  }
}

template <bool A_KC, bool B_KC>
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
  const bool a_vec = A_KC && (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 31) == 0) && (kbeg % 8 == 0);
  const bool b_vec = B_KC && (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(B) & 31) == 0) && (kbeg % 8 == 0);
  float va[4][8], vb[4][8];
  if (kbeg < kend) {
    gt_load<A_KC>(va, A, lda, i0, M, kbeg, kend, a_vec);
    gt_load<B_KC>(vb, B, ldb, j0, N, kbeg, kend, b_vec);
  }
  for (int k0 = kbeg; k0 < kend; k0 += GT_K, ++it) {
    const int s = it & 1;
    if (it >= 2) {  // the MMAs that read this stage two slices ago have completed
      mbar_wait(&bars[s], ph[s]);
      ph[s] ^= 1;
    }
    gt_store<A_KC>(a_tiles + s * GT_TILE_BYTES, va);
    gt_store<B_KC>(b_tiles + s * GT_TILE_BYTES, vb);
    if (k0 + GT_K < kend) {  // the next slice's operands fly while this slice's MMAs run
      gt_load<A_KC>(va, A, lda, i0, M, k0 + GT_K, kend, a_vec);
      gt_load<B_KC>(vb, B, ldb, j0, N, k0 + GT_K, kend, b_vec);
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

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_gemm_bf16(int a_k_contiguous, int b_k_contiguous, int M, int N, int K, const float* A, int lda, const float* B,
                                int ldb, float* C, int ldc, const float* bias, int beta, int split_k, dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && split_k >= 1, DFB200_ERR_INVALID_ARG, "gemm_bf16: bad sizes M=%d N=%d K=%d split=%d", M, N, K, split_k);
  if (M == 0 || N == 0) return DFB200_OK;
  int kps = cdiv(cdiv(K, split_k), GT_K) * GT_K;
  if (kps == 0) kps = GT_K;
  const int splits = K == 0 ? 1 : cdiv(K, kps);
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
  if (a_k_contiguous && b_k_contiguous) GT_LAUNCH(true, true);
  else if (a_k_contiguous) GT_LAUNCH(true, false);
  else if (b_k_contiguous) GT_LAUNCH(false, true);
  else GT_LAUNCH(false, false);
#undef GT_LAUNCH
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
