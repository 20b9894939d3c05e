// capi_internal.h -- error plumbing shared by the extern "C" entry points (include/shiftgcn_b200.h).
// Every entry point returns 0 on success or a negative code; the message is kept per host thread and
// read back with sgcn_last_error().  Kernels are launched on the caller's stream, never synchronise,
// and are therefore CUDA-graph capturable.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace sgcn {
int set_error(const char* msg);                              // returns -1
int set_cuda_error(const char* where, cudaError_t e);        // returns -2
int check_launch(const char* kernel_name);                   // cudaGetLastError() -> 0 / -2
int num_sms();                                               // SM count of the current device (cached)
int tile_ctas();                                             // grid of the persistent tile kernels: num_sms() or the test cap
// per-device "dynamic shared memory attribute set" bookkeeping of one kernel instantiation (a static per launch function)
bool needs_configure(std::atomic<unsigned long long>& done_mask);
void mark_configured(std::atomic<unsigned long long>& done_mask);
// "Snake" traversal (sgcn_set_traversal): every full-tensor kernel walks its tiles in the opposite order of the kernel
// launched before it on this host thread, so it starts on the ~100 MB that are still resident in the 126 MB L2.
int next_direction();                                        // 0 = ascending, 1 = descending; flips when enabled
void mark_forward();                                         // a kernel without a reversed order was launched
}  // namespace sgcn
