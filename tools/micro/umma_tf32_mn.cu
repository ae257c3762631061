// Probe (GPU box): shared-memory layout / descriptor of an MN-MAJOR tf32 B operand for tcgen05.mma kind::tf32 (SWIZZLE_128B).
// D[128 x 128] = A[128 x 32] (K-major, known-good SW128 layout) . B[32 x 128] with B given as B[k][n] (n contiguous).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dfb200::tc;
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int amn, int bmn, int bit_a, int bit_b) {
  return make_idesc_tf32(M, N) | ((uint32_t)amn << bit_a) | ((uint32_t)bmn << bit_b);
}
// variant: bits 0-1 smem arrangement (0: atom a at a*4096, k-group g at g*1024; 1: a*1024, g*4096), bit 2: swap LBO/SBO in the descriptor,
// bit 3: no XOR swizzle in the data (descriptor still SW128), bit 4: major bits at 16/17 instead of 15/16, bit 5: k-step advance by 128 B (one k row x8?) experimental
__global__ void __launch_bounds__(128, 1) k(int variant, const float* A, const float* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem; uint8_t* b_tile = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  const int arr = variant & 3;
  const uint32_t a_str = arr == 0 ? 4096 : 1024, g_str = arr == 0 ? 1024 : 4096;
  // A: row tid, K-major SW128: row r at r*128, chunk c at (c ^ (r&7))*16
  for (int c = 0; c < 8; ++c) {
    float4 v = make_float4(A[tid * 32 + 4 * c], A[tid * 32 + 4 * c + 1], A[tid * 32 + 4 * c + 2], A[tid * 32 + 4 * c + 3]);
    *reinterpret_cast<float4*>(a_tile + tid * 128 + ((c ^ (tid & 7)) << 4)) = v;
  }
  // B[k][n]: k = 0..31, n = 0..127
  for (int i = tid; i < 32 * 32; i += 128) {
    const int kr = i / 32, c16 = i % 32;
    float4 v = make_float4(B[kr * 128 + 4 * c16], B[kr * 128 + 4 * c16 + 1], B[kr * 128 + 4 * c16 + 2], B[kr * 128 + 4 * c16 + 3]);
    const int sw = (variant & 8) ? (c16 & 7) : ((c16 & 7) ^ (kr & 7));
    *reinterpret_cast<float4*>(b_tile + (c16 >> 3) * a_str + (kr >> 3) * g_str + (kr & 7) * 128 + (sw << 4)) = v;
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      const int ba = (variant & 16) ? 16 : 15, bb = (variant & 16) ? 17 : 16;
      const uint32_t idesc = idesc_tf32(128, 128, 0, 1, ba, bb);
      const uint32_t lbo = (variant & 4) ? g_str : a_str, sbo = (variant & 4) ? a_str : g_str;
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_smem_desc(smem_u32(a_tile) + ks * 32, 16, 1024) | ((uint64_t)2 << 61);
        const uint64_t bd = make_smem_desc(smem_u32(b_tile) + ks * g_str, lbo, sbo) | ((uint64_t)2 << 61);
        umma_tf32(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) D[tid * 128 + cb * 32 + j] = h[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}
int main() {
  std::vector<float> A(128 * 32), B(32 * 128), D(128 * 128);
  srand(3);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 500.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 500.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  std::vector<double> ref(128 * 128);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 128; ++n) { double s = 0; for (int kk = 0; kk < 32; ++kk) s += (double)A[r * 32 + kk] * B[kk * 128 + n]; ref[r * 128 + n] = s; }
  for (int v = 0; v < 32; ++v) {
    if ((v & 3) > 1) continue;
    cudaMemset(dD, 0, D.size() * 4);
    k<<<1, 128, 40000>>>(v, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double me = 0; int bad = 0;
    for (int i = 0; i < 128 * 128; ++i) { double d = fabs(D[i] - ref[i]); me = fmax(me, d); bad += d > 0.05; }
    printf("variant %2d (arr %d, swapLS %d, noswz %d, bits1617 %d): max |err| %.3e, %d / 16384 entries off\n", v, v & 3, (v >> 2) & 1, (v >> 3) & 1, (v >> 4) & 1, me, bad);
  }
  return 0;
}
