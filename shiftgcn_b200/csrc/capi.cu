// capi.cu -- library-wide pieces of the C ABI declared in include/shiftgcn_b200.h.
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "capi_internal.h"

namespace sgcn {

static thread_local char g_err[512] = "";

int set_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}

int set_cuda_error(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return -2;
}

int check_launch(const char* kernel_name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(kernel_name, e);
  return 0;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// process-wide (autograd runs the backward kernels from its own host thread); a race between two launching threads
// can only pick a less favourable order, never a wrong result
static std::atomic<int> g_snake{0}, g_last_rev{0};

int next_direction() {
  if (!g_snake.load(std::memory_order_relaxed)) return 0;
  return g_last_rev.fetch_xor(1, std::memory_order_relaxed) ^ 1;
}

void mark_forward() { g_last_rev.store(0, std::memory_order_relaxed); }

}  // namespace sgcn

extern "C" int sgcn_set_traversal(int snake) {
  sgcn::g_last_rev.store(snake == 2 ? 1 : 0);                  // 1: the next kernel descends; 2: it ascends
  return sgcn::g_snake.exchange(snake ? 1 : 0);
}

extern "C" const char* sgcn_last_error(void) { return sgcn::g_err; }

extern "C" int sgcn_abi_version(void) { return 1; }

extern "C" int sgcn_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return sgcn::set_cuda_error("cudaGetDevice", e);
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    char buf[128];
    snprintf(buf, sizeof(buf), "shiftgcn_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return sgcn::set_error(buf);
  }
  return 0;
}
