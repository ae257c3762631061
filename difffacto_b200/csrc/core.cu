// Library-wide state of the C ABI: status messages, launch counter, ABI version.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace dfb200 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int current_device_ordinal() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int current_device_sm_count() {
  static int cache[64] = {};
  const int dev = current_device_ordinal();
  if (dev < 0) return 0;
  if (dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev < 64) cache[dev] = n;
  return n;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dfb200

extern "C" int dfb200_abi_version(void) { return 1; }
extern "C" const char* dfb200_last_error(void) { return dfb200::g_err; }
extern "C" unsigned long long dfb200_launch_count(void) {
  return dfb200::g_launches.load(std::memory_order_relaxed);
}
