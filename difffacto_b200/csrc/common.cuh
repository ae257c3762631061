// Shared host/device helpers for the difffacto_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/difffacto_b200.h"

#ifndef __CUDA_ARCH__
#define DFB200_HOST_ONLY 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "difffacto_b200 targets sm_100a (B200) only"
#endif

namespace dfb200 {

// ---- error reporting -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);

#define DFB_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::dfb200::set_error(__VA_ARGS__); \
      return (code);                   \
    }                                  \
  } while (0)

#define DFB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::dfb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                     \
      return DFB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

// After a <<<>>> launch: count it and surface launch-configuration errors as a status code
// (the reference prints and calls exit(-1), cuda_utils.h:30-39).
#define DFB_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ::dfb200::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::dfb200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                          __FILE__, __LINE__);                                           \
      return DFB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

// Per-device facts (a process may drive several GPUs through `_lib.on(device)`): nothing device-specific is cached in a
// process-wide static.  current_device_sm_count() returns the SM count of the CURRENT device (0 on error);
// DeviceOnce remembers, per device ordinal, whether a one-time setup (cudaFuncSetAttribute is per device) has run.
int current_device_sm_count();
int current_device_ordinal();  // -1 on error
struct DeviceOnce {
  bool done[64] = {};
  // true the first time it is asked for the current device (ordinals >= 64 are never remembered: the setup just repeats)
  bool first_time() {
    const int d = current_device_ordinal();
    if (d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

static inline cudaStream_t as_stream(dfb200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers ------------------------------------------------------------------------
// Squared norm with exactly the rounding sequence nvcc emits for the reference kernels'
// `dx*dx + dy*dy + dz*dz` (fma(dz,dz, fma(dx,dx, mul(dy,dy))), checked in the reference PTX):
// index outputs that depend on distance comparisons are bit-exact only with this order.
__device__ __forceinline__ float sq3(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace dfb200
