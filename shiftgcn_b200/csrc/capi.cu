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

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  const int dev = current_device();
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

// Persistent-grid cap (sgcn_set_max_ctas; tests): with a few CTAs even a small tensor gives every CTA many tiles, i.e.
// the multi-tile steady state of the pipelines (second accumulator, ring wrap, phase parities) the full-size runs use.
static std::atomic<int> g_max_ctas{0};

int tile_ctas() {
  const int n = num_sms(), cap = g_max_ctas.load(std::memory_order_relaxed);
  return (cap > 0 && cap < n) ? cap : n;
}

// The on/off setting is process-wide; the alternation itself is kept PER DEVICE: autograd runs the backward kernels of
// a device from its own host thread, so the state cannot be per host thread, while threads that drive different
// devices (nn.DataParallel replicas) never share an entry.  Two threads launching on the same device can only pick a
// less favourable order, never a wrong result.
constexpr int kMaxDevices = 64;
static std::atomic<int> g_snake{0}, g_last_rev[kMaxDevices];

static int dev_slot() { return current_device() % kMaxDevices; }

int next_direction() {
  if (!g_snake.load(std::memory_order_relaxed)) return 0;
  return g_last_rev[dev_slot()].fetch_xor(1, std::memory_order_relaxed) ^ 1;
}

void mark_forward() { g_last_rev[dev_slot()].store(0, std::memory_order_relaxed); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE property of a kernel: remember which devices a
// kernel has been configured on (one bit each) instead of a per-thread flag that a second GPU would inherit.
bool needs_configure(std::atomic<unsigned long long>& done_mask) {
  return !(done_mask.load(std::memory_order_acquire) >> dev_slot() & 1ull);
}
void mark_configured(std::atomic<unsigned long long>& done_mask) {
  done_mask.fetch_or(1ull << dev_slot(), std::memory_order_release);
}

}  // namespace sgcn

extern "C" int sgcn_set_traversal(int snake) {
  sgcn::g_last_rev[sgcn::dev_slot()].store(snake == 2 ? 1 : 0);   // 1: the next kernel descends; 2: it ascends
  return sgcn::g_snake.exchange(snake ? 1 : 0);
}

extern "C" int sgcn_set_max_ctas(int n) { return sgcn::g_max_ctas.exchange(n > 0 ? n : 0); }

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
