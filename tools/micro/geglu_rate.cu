// Microbenchmark (GPU box): issue rate of the bf16 kernel's GEGLU epilogue math as a function of warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/geglu_rate tools/micro/geglu_rate.cu && /tmp/geglu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int BIAS>
__device__ __forceinline__ float2 geglu2(float2 a_half, float2 g, float2 ba_half, float2 bg) {
  if (BIAS) { a_half = __fadd2_rn(a_half, ba_half); g = __fadd2_rn(g, bg); }
  const float2 g2 = __fmul2_rn(g, g);
  const float2 in = __fmul2_rn(g, __ffma2_rn(g2, f2s(0.034700932528f), f2s(0.800156991001f)));
  const float2 t = f2(tanh_approx(in.x), tanh_approx(in.y));
  const float2 ag = __fmul2_rn(a_half, g);
  return __ffma2_rn(ag, t, ag);
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) { __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<unsigned*>(&t); }

// each thread: ITER rounds of PAIRS independent column pairs (as one epilogue block), data-dependent across rounds so nothing is hoisted
template <int PAIRS, int BIAS>
__global__ void k(const float* in, unsigned* out, long long* cyc, int iters) {
  float a[2 * PAIRS], g[2 * PAIRS], ba[2 * PAIRS], bg[2 * PAIRS];
  for (int i = 0; i < 2 * PAIRS; ++i) { a[i] = in[threadIdx.x + i * 32]; g[i] = in[threadIdx.x + i * 64 + 7]; ba[i] = in[i]; bg[i] = in[i + 64]; }
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    unsigned u[PAIRS];
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
      const float2 y = geglu2<BIAS>(f2(a[2 * p], a[2 * p + 1]), f2(g[2 * p], g[2 * p + 1]), f2(ba[2 * p], ba[2 * p + 1]), f2(bg[2 * p], bg[2 * p + 1]));
      u[p] = pack_bf16(y.x, y.y);
    }
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) { acc ^= u[p]; a[2 * p] = __uint_as_float((u[p] & 0x007fffffu) | 0x3f000000u); g[2 * p + 1] += 0.01f; }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int PAIRS, int BIAS>
void run(int warps, const float* in, unsigned* out, long long* cyc) {
  const int iters = 200;
  k<PAIRS, BIAS><<<148, warps * 32>>>(in, out, cyc, iters);
  k<PAIRS, BIAS><<<148, warps * 32>>>(in, out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_pair_warp = (double)c / iters / PAIRS;                     // cycles per pair as seen by one warp
  const double per_smsp = per_pair_warp / (warps / 4.0);                       // scheduler cycles per (pair x warp)
  printf("pairs/block %2d bias %d warps/SM %2d (%.0f per scheduler): %.1f cycles per pair per warp, %.1f scheduler cycles per pair-warp (MUFU floor 16)\n",
         PAIRS, BIAS, warps, warps / 4.0, per_pair_warp, per_smsp);
}
int main() {
  float* in; unsigned* out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0x3c, 1 << 20);
  for (int w : {4, 8, 12, 16, 24, 32}) { run<8, 1>(w, in, out, cyc); }
  for (int w : {4, 8, 12, 16, 24, 32}) { run<16, 1>(w, in, out, cyc); }
  for (int w : {8, 16}) { run<8, 0>(w, in, out, cyc); run<16, 0>(w, in, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
