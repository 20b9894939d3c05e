// capi_internal.h -- error plumbing shared by the extern "C" entry points (include/shiftgcn_b200.h).
// Every entry point returns 0 on success or a negative code; the message is kept per host thread and
// read back with sgcn_last_error().  Kernels are launched on the caller's stream, never synchronise,
// and are therefore CUDA-graph capturable.
#pragma once
#include <cuda_runtime.h>

namespace sgcn {
int set_error(const char* msg);                              // returns -1
int set_cuda_error(const char* where, cudaError_t e);        // returns -2
int check_launch(const char* kernel_name);                   // cudaGetLastError() -> 0 / -2
int num_sms();                                               // SM count of the current device (cached)
}  // namespace sgcn
