// head.cu -- output side of the path: global pooling over (T, V) and the persons, and the classifier
// (model/shift_gcn.py:212-216:  x.view(N, M, C, -1).mean(3).mean(1);  fc(x)).
//
// The pooled SUMS arrive from the kernel that produces the last unit's output (sgcn_tshift_fwd, mode 1 with a stats
// pointer): the [n*M, C] fp64 buffer holds sum over (t, v) of y.  Forward = one block per sample: mean over persons and
// rows, then K dot products of length C.  Backward = one kernel for fc.weight.grad, fc.bias.grad and the gradient of the
// pooled features, which sgcn_bcast_rows then spreads over the rows of the last unit's output.
#include "capi_internal.h"
#include "shiftgcn_b200.h"

namespace sgcn {

__global__ void __launch_bounds__(256) head_fwd_kernel(double* __restrict__ pool_sums, const float* __restrict__ W,
                                                       const float* __restrict__ b, float* __restrict__ pooled,
                                                       float* __restrict__ logits, int M, int C, int K, double inv_count) {
  extern __shared__ float feat[];                                  // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;
    for (int m = 0; m < M; ++m) {
      double* q = pool_sums + ((size_t)n * M + m) * C + c;
      s += *q;
      *q = 0.0;                                                    // handed back zeroed for the next step
    }
    const float f = (float)(s * inv_count);
    feat[c] = f;
    pooled[(size_t)n * C + c] = f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = warp; k < K; k += nw) {                             // one warp per class: coalesced rows of W
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(W + (size_t)k * C + c), feat[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[(size_t)n * K + k] = acc + (b ? __ldg(b + k) : 0.f);
  }
}

// thread i < K*C: dW[k, c] = sum_n dl[n, k] * pooled[n, c];  i < K: db[k] = sum_n dl[n, k];
// thread j < N*M*C: gpool[(n, m), c] = scale * sum_k dl[n, k] * W[k, c]
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ pooled,
                                                       const float* __restrict__ W, float* __restrict__ dW,
                                                       float* __restrict__ db, float* __restrict__ gpool, int N, int M,
                                                       int C, int K, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nw = (long long)K * C, ng = (long long)N * M * C;
  if (i < nw) {
    const int k = (int)(i / C), c = (int)(i % C);
    float acc = 0.f, accb = 0.f;
    for (int n = 0; n < N; ++n) {
      const float d = __ldg(dl + (size_t)n * K + k);
      acc = fmaf(d, __ldg(pooled + (size_t)n * C + c), acc);
      accb += d;
    }
    dW[i] = acc;
    if (c == 0 && db) db[k] = accb;
  } else if (i < nw + ng) {
    const long long j = i - nw;
    const int c = (int)(j % C);
    const long long n = j / ((long long)C * M);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__ldg(dl + (size_t)n * K + k), __ldg(W + (size_t)k * C + c), acc);
    gpool[j] = acc * scale;
  }
}

}  // namespace sgcn

extern "C" int sgcn_head_fwd(double* pool_sums, const float* W, const float* b, float* pooled, float* logits, int N, int M,
                             int C, int K, double inv_count, void* stream) {
  using namespace sgcn;
  if (N < 0 || M < 1 || C < 1 || K < 1 || C > 4096) return set_error("sgcn_head_fwd: bad shape");
  if (N == 0) return 0;
  if (!pool_sums || !W || !pooled || !logits) return set_error("sgcn_head_fwd: null pointer");
  head_fwd_kernel<<<N, 256, C * sizeof(float), (cudaStream_t)stream>>>(pool_sums, W, b, pooled, logits, M, C, K, inv_count);
  return check_launch("head_fwd_kernel");
}

extern "C" int sgcn_head_bwd(const float* dlogits, const float* pooled, const float* W, float* dW, float* db, float* gpool,
                             int N, int M, int C, int K, float scale, void* stream) {
  using namespace sgcn;
  if (N < 0 || M < 1 || C < 1 || K < 1) return set_error("sgcn_head_bwd: bad shape");
  if (!dlogits || !pooled || !W || !dW || !gpool) return set_error("sgcn_head_bwd: null pointer");
  const long long total = (long long)K * C + (long long)N * M * C;
  head_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dlogits, pooled, W, dW, db, gpool, N, M, C,
                                                                                      K, scale);
  return check_launch("head_bwd_kernel");
}
