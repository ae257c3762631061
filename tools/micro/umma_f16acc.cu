// Microbenchmark / probe (GPU box): tcgen05.mma kind::f16 with F16 operands and an F16 accumulator in TMEM.
//   (1) how are the N accumulator columns packed into 32-bit TMEM columns, (2) accuracy vs fp32 accumulation,
//   (3) issue rate, (4) TS form with an f16 A operand in TMEM accumulating into an fp32 D.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I difffacto_b200/csrc -o tools/micro/_bin/umma_f16acc tools/micro/umma_f16acc.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace dfb200::tc;

__host__ __device__ constexpr uint32_t make_idesc(int cfmt, int afmt, int bfmt, int M, int N) {
  return ((uint32_t)cfmt << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) { __half2 t = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&t); }

// mode 0: D(f16, TMEM cols [0,N/2)) = A.W^T ; mode 1: D(f32) = A.W^T (f16 operands) ; mode 2: TS form, A f16 in TMEM cols [256, 256+K/2), D f32
// mode 3: rate of mode 0 (iters MMAs alternating two accumulators), mode 4: rate of mode 1
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int K, const float* A, const float* W, uint32_t* Draw, long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem; uint8_t* b_tile = smem + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  for (int kk = 0; kk < K; kk += 2) *reinterpret_cast<uint32_t*>(a_tile + tile_off(128, tid, kk)) = pack_f16(A[tid * K + kk], A[tid * K + kk + 1]);
  for (int i = tid; i < N * K / 2; i += 128) {
    const int n = i / (K / 2), kk = (i - n * (K / 2)) * 2;
    *reinterpret_cast<uint32_t*>(b_tile + tile_off(N, n, kk)) = pack_f16(W[n * K + kk], W[n * K + kk + 1]);
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  if (mode == 2) {
    for (int cb = 0; cb < K / 2 / 32; ++cb) {
      float h[32];
      for (int j = 0; j < 32; ++j) { const int kk = (cb * 32 + j) * 2; h[j] = __uint_as_float(pack_f16(A[tid * K + kk], A[tid * K + kk + 1])); }
      tmem_st32(row_addr + 256 + cb * 32, h);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    tc_fence_after();
    if (lane == 0) {
      const int cf = (mode == 0 || mode == 3) ? 0 : 1;
      const uint32_t idesc = make_idesc(cf, 0, 0, 128, N);
      t0 = clock64();
      const int reps = mode >= 3 ? iters : 1;
      for (int it = 0; it < reps; ++it) {
        const uint32_t d = tmem + (it & 1) * 128;
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t ad = make_smem_desc(smem_u32(a_tile) + ks * 4096, 2048, TILE_SBO);
          const uint64_t bd = make_smem_desc(smem_u32(b_tile) + ks * (N * 32), (uint32_t)(N * 16), TILE_SBO);
          if (mode == 2) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d), "r"(tmem + 256 + ks * 8), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
          }
        }
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  if (tid == 0) { t1 = clock64(); *cyc = t1 - t0; }
  tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {  // dump the first 128 TMEM columns raw
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) Draw[tid * 128 + cb * 32 + j] = __float_as_uint(h[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static float h2f(uint16_t h) { __half x; *reinterpret_cast<uint16_t*>(&x) = h; return __half2float(x); }
int main() {
  const int N = 128, K = 128;
  std::vector<float> A(128 * K), W(N * K);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 400.f;   // ~ LayerNorm output scale
  for (auto& v : W) v = (rand() % 2001 - 1000) / 8000.f;  // ~ weights
  float *dA, *dW; uint32_t* dD; long long* dc;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, 128 * 128 * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  std::vector<double> ref(128 * N);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)A[r * K + kk] * W[n * K + kk]; ref[r * N + n] = s; }
  std::vector<uint32_t> D(128 * 128);
  for (int mode : {1, 2, 0}) {
    cudaMemset(dD, 0, 128 * 128 * 4);
    k<<<1, 128, 70000>>>(mode, N, K, dA, dW, dD, dc, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    if (mode != 0) {
      double me = 0, mr = 0;
      for (int i = 0; i < 128 * N; ++i) { float v; memcpy(&v, &D[i], 4); me = fmax(me, fabs(v - ref[i])); mr = fmax(mr, fabs(ref[i])); }
      printf("mode %d (%s, f32 accumulator): max |err| %.3e (max |ref| %.3f)\n", mode, mode == 1 ? "SS f16 operands" : "TS f16 A in TMEM", me, mr);
    } else {
      // hypothesis P: column j holds (n=2j low half, n=2j+1 high half); hypothesis Q: column j holds (n=j low, n=j+64 high)
      double eP = 0, eQ = 0;
      for (int r = 0; r < 128; ++r) for (int j = 0; j < N / 2; ++j) {
        const uint32_t w = D[r * 128 + j];
        const float lo = h2f(w & 0xffff), hi = h2f(w >> 16);
        eP = fmax(eP, fmax(fabs(lo - ref[r * N + 2 * j]), fabs(hi - ref[r * N + 2 * j + 1])));
        eQ = fmax(eQ, fmax(fabs(lo - ref[r * N + j]), fabs(hi - ref[r * N + j + 64])));
      }
      printf("mode 0 (f16 accumulator): max |err| if column j = (n=2j, 2j+1): %.3e ; if column j = (n=j, j+N/2): %.3e\n", eP, eQ);
      double rms = 0, rr = 0; for (int r = 0; r < 128; ++r) for (int j = 0; j < N / 2; ++j) { const uint32_t w = D[r * 128 + j]; double d0 = h2f(w & 0xffff) - ref[r * N + 2 * j]; rms += d0 * d0; rr += ref[r * N + 2 * j] * ref[r * N + 2 * j]; }
      printf("   relative rms error under hypothesis P: %.3e ; raw row 0 cols 0..3: %08x %08x %08x %08x ; col 64..65: %08x %08x\n", sqrt(rms / rr), D[0], D[1], D[2], D[3], D[64], D[65]);
    }
  }
  for (int mode : {3, 4}) {
    k<<<1, 128, 70000>>>(mode, N, K, dA, dW, dD, dc, 64);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d rate (%s accumulator): %.1f cycles per M128 N128 K16 MMA (512 MMAs)\n", mode, mode == 3 ? "f16" : "f32", (double)c / (64 * 8));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
