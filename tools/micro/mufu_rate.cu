// Microbenchmark (GPU box): MUFU throughput per scheduler for the ops a GELU/sigmoid epilogue could use.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int OP> __device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) { unsigned u = __float_as_uint(x), v; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 4) { unsigned u = __float_as_uint(x), v; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 5) { unsigned u = __float_as_uint(x), v; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 6) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 7) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int OP> __global__ void k(const float* in, float* out, long long* cyc, int iters) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = in[threadIdx.x + 32 * i];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = op<OP>(v[i]);
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int warps, const float* in, float* out, long long* cyc) {
  const int iters = 256;
  k<OP><<<148, warps * 32>>>(in, out, cyc, iters);
  k<OP><<<148, warps * 32>>>(in, out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-18s warps/scheduler %d: %.2f scheduler cycles per warp-instruction (PTX op; packed ops = 2 values/lane)\n", name, warps / 4,
         (double)c / iters / 16 / (warps / 4.0));
}
int main() {
  float* in; float* out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0x3c, 1 << 20);
  for (int w : {8, 16}) {
    run<0>("tanh.f32", w, in, out, cyc); run<1>("ex2.f32", w, in, out, cyc); run<2>("rcp.f32", w, in, out, cyc);
    run<3>("tanh.f16x2", w, in, out, cyc); run<4>("ex2.f16x2", w, in, out, cyc); run<5>("tanh.bf16x2", w, in, out, cyc);
    run<6>("rsqrt.f32", w, in, out, cyc); run<7>("lg2.f32", w, in, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
