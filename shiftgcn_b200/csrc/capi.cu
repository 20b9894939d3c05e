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

// ------------------------------------------------------------------------------------------------ TMA tensor maps
#include <cudaTypedefs.h>

#include "tensormap.h"

namespace sgcn {

static PFN_cuTensorMapEncodeTiled encode_fn() {
  static std::atomic<void*> cached{nullptr};
  void* fn = cached.load(std::memory_order_acquire);
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    cached.store(fn, std::memory_order_release);
  }
  return (PFN_cuTensorMapEncodeTiled)fn;
}

static int encode(CUtensorMap* m, int rank, const float* base, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapSwizzle sw) {
  PFN_cuTensorMapEncodeTiled fn = encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled is not available from this driver");
  if (((uintptr_t)base & 15u) != 0) return set_error("TMA source tensors must be 16-byte aligned");
  const cuuint32_t ones[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return set_error(buf);
  }
  return 0;
}

int make_rows_map(CUtensorMap* m, const float* base, long long rows, int pitch, int box_ch, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_ch, (cuuint32_t)box_rows};
  return encode(m, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int make_frames_map(CUtensorMap* m, const float* base, long long frames, int V, int pitch, int box_ch, int box_v,
                    int box_frames) {
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)V, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)V * pitch * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_ch, (cuuint32_t)box_v, (cuuint32_t)box_frames};
  return encode(m, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int make_groups_map_sw128(CUtensorMap* m, const float* base, long long groups, long long gs, int V, int pitch,
                          int box_groups) {
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)V, (cuuint64_t)groups};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)(gs > 0 ? gs : 1) * V * pitch * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)V, (cuuint32_t)box_groups};
  return encode(m, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sgcn
