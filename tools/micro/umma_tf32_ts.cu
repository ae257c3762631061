// Probe (GPU box): tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (TS form): correctness and issue rate vs the SS form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I difffacto_b200/csrc -o tools/micro/_bin/umma_tf32_ts tools/micro/umma_tf32_ts.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dfb200::tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mode 0: SS, mode 1: TS (A tf32 in TMEM columns [256, 256+K)); rate != 0: `rate` x 16 MMAs back to back, alternating accumulators
__global__ void __launch_bounds__(128, 1) k(int mode, int rate, const float* A, const float* W, float* D, long long* cyc) {
  constexpr int N = 128, K = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem; uint8_t* b_tile = smem + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 131072 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  for (int kk = 0; kk < K; ++kk) *reinterpret_cast<float*>(a_tile + tile_off32(128, tid, kk)) = to_tf32(A[tid * K + kk]);
  for (int i = tid; i < N * K; i += 128) { const int n = i / K, kk = i - n * K; *reinterpret_cast<float*>(b_tile + tile_off32(N, n, kk)) = to_tf32(W[n * K + kk]); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    for (int j = 0; j < 32; ++j) h[j] = to_tf32(A[tid * K + cb * 32 + j]);
    tmem_st32(row_addr + 256 + cb * 32, h);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  long long t0 = 0;
  if (warp == 0) {
    tc_fence_after();
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, N);
      const uint32_t aa = smem_u32(a_tile), bb = smem_u32(b_tile);
      t0 = clock64();
      const int reps = rate ? rate : 1;
#pragma unroll 1
      for (int it = 0; it < reps; ++it) {
        const uint32_t d = tmem + (it & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t bd = make_smem_desc(bb + ks * 4096, 2048, TILE_SBO);
          if (mode == 1) umma_tf32_ts(d, tmem + 256 + ks * 8, bd, idesc, ks > 0 ? 1u : 0u);
          else umma_tf32(d, make_smem_desc(aa + ks * 4096, 2048, TILE_SBO), bd, idesc, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  if (tid == 0) *cyc = clock64() - t0;
  tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) D[tid * 128 + cb * 32 + j] = h[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
int main() {
  const int N = 128, K = 128;
  std::vector<float> A(128 * K), W(N * K), D(128 * 128);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 400.f;
  for (auto& v : W) v = (rand() % 2001 - 1000) / 8000.f;
  float *dA, *dW, *dD; long long* dc;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, 128 * 128 * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 131072 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<double> ref(128 * N);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)A[r * K + kk] * W[n * K + kk]; ref[r * N + n] = s; }
  for (int mode : {0, 1}) {
    k<<<1, 128, smem>>>(mode, 0, dA, dW, dD, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double me = 0; for (int i = 0; i < 128 * N; ++i) me = fmax(me, fabs(D[i] - ref[i]));
    printf("kind::tf32 %s form: max |err| vs fp64 %.3e\n", mode ? "TS (A in TMEM)" : "SS", me);
    k<<<1, 128, smem>>>(mode, 32, dA, dW, dD, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("kind::tf32 %s form: %.1f cycles per M128 N128 K8 MMA (512 back to back)\n", mode ? "TS" : "SS", (double)c / 512);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
