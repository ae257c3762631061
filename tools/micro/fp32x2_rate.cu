// Microbenchmark (GPU box): issue rate of the packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2) and of scalar FFMA, per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(const float* in, float* out, long long* cyc, int iters) {
  float2 v[8], a = make_float2(in[threadIdx.x], in[threadIdx.x + 32]), b = make_float2(in[threadIdx.x + 64], in[threadIdx.x + 96]);
  for (int i = 0; i < 8; ++i) v[i] = make_float2(in[threadIdx.x + 128 + i], in[threadIdx.x + 256 + i]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) v[i] = __fadd2_rn(v[i], a);
      if (OP == 1) v[i] = __fmul2_rn(v[i], a);
      if (OP == 2) v[i] = __ffma2_rn(v[i], a, b);
      if (OP == 3) { v[i].x = fmaf(v[i].x, a.x, b.x); v[i].y = fmaf(v[i].y, a.y, b.y); }
      if (OP == 4) { v[i].x = v[i].x + a.x; v[i].y = v[i].y + a.y; }
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int warps, const float* in, float* out, long long* cyc) {
  const int iters = 512;
  k<OP><<<148, warps * 32>>>(in, out, cyc, iters); k<OP><<<148, warps * 32>>>(in, out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const int per_it = OP >= 3 ? 16 : 8;
  printf("%-12s warps/scheduler %d: %.2f scheduler cycles per warp instruction\n", name, warps / 4, (double)c / iters / per_it / (warps / 4.0));
}
int main() {
  float* in; float* out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0x3c, 1 << 20);
  for (int w : {8, 16, 32}) { run<0>("FADD2", w, in, out, cyc); run<1>("FMUL2", w, in, out, cyc); run<2>("FFMA2", w, in, out, cyc); run<3>("FFMA scalar", w, in, out, cyc); run<4>("FADD scalar", w, in, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
