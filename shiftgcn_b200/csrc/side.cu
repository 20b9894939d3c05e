// side.cu -- parameter-sized arithmetic of the conv + BatchNorm side branches (`down` of Shift_gcn,
// model/shift_gcn.py:82-86; the strided `tcn` residual, :31-45, 157-158) in fp64, as three small kernels.
//
// The full-size work of those branches is done by the tensor-core kernels (one Gram matrix X^T X, one correlation
// X^T G, one folded forward GEMM and one [G | x] input-gradient GEMM, see shiftgcn_b200/functional.py:side_forward /
// side_backward).  What is left is C x C x D arithmetic on the weights.  As a chain of library calls that was ~80
// launches per branch and step (~320 of the ~600 small launches of a training step); here it is 1 + 2.
//
//   fold (forward):   mu = sx/rows,  mean_r = Wd mu + bd,  var_r[d] = Wd[d] (XX/rows - mu mu^T) Wd[d]^T,
//                     Wf = (gamma*invstd) Wd,  bf = beta + gamma*invstd*(bd - mean_r)        (+ running statistics)
//   coeffs (backward): dgamma, dbeta, the affine BatchNorm-backward coefficients (al, be, ga), dWd, dbd and the first
//                     D rows of the input-gradient weight  Wcat = [ al*Wd ; Wd^T be Wd ]
//   mix (backward):   the last C rows of Wcat and the constant term  kvec = Wd^T (be*bd + ga)
#include "capi_internal.h"
#include "shiftgcn_b200.h"

namespace sgcn {
namespace side {

constexpr int kThreads = 128;
constexpr int kMaxC = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
  // all threads of the block call this; returns the block-wide sum to every thread
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kThreads / 32; ++w) s += red[w];
  return s;
}

// sum_k v[k] * M[k * ld + col] in fp64 with four independent accumulators: a thread walks one COLUMN of the row-major
// fp32 matrix, so a warp reads consecutive addresses, and the dependent-FMA chain is a quarter as long (the serial
// version of these parameter-sized loops took 18 / 23 / 53 us per launch at C = 128, D = 256)
__device__ __forceinline__ double col_dot(const double* __restrict__ v, const float* __restrict__ M, int n, int ld, int col) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const float* m = M + col;
  int k = 0;
  for (; k + 15 < n; k += 16) {                                     // 16 independent loads in flight (L2 latency bound otherwise)
    float mv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) mv[j] = __ldg(m + (size_t)(k + j) * ld);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      a0 = fma(v[k + j], (double)mv[j], a0);
      a1 = fma(v[k + j + 1], (double)mv[j + 1], a1);
      a2 = fma(v[k + j + 2], (double)mv[j + 2], a2);
      a3 = fma(v[k + j + 3], (double)mv[j + 3], a3);
    }
  }
  for (; k < n; ++k) a0 = fma(v[k], (double)__ldg(m + (size_t)k * ld), a0);
  return (a0 + a1) + (a2 + a3);
}

// one block per output channel d
__global__ void __launch_bounds__(kThreads) fold_kernel(const SgcnSideFold p) {
  __shared__ double mu[kMaxC], wrow[kMaxC], red[kThreads / 32];
  const int d = blockIdx.x, C = p.C, tid = threadIdx.x;
  const double bd = p.bd ? (double)p.bd[d] : 0.0;
  double mean_r, var_r;
  if (p.training) {
    for (int c = tid; c < C; c += kThreads) {
      mu[c] = p.sx_sums[2 * c] / p.rows;
      wrow[c] = (double)p.Wd[(size_t)d * C + c];
    }
    __syncthreads();
    double a = 0.0, q = 0.0;
    for (int c = tid; c < C; c += kThreads) {
      a += wrow[c] * mu[c];
      const double t = col_dot(wrow, p.XX, C, C, c);              // (Wd[d] XX)[c]
      q += t * wrow[c];
    }
    const double wmu = block_sum(a, red);
    const double quad = block_sum(q, red);
    mean_r = wmu + bd;
    var_r = quad / p.rows - wmu * wmu;
    if (var_r < 0.0) var_r = 0.0;
    if (tid == 0 && p.running_mean) {
      const double unbiased = var_r * (p.rows / (p.rows > 1.0 ? p.rows - 1.0 : 1.0));
      // momentum < 0 = nn.BatchNorm(momentum=None): cumulative average, counter already incremented by the caller
      const bool cma = p.momentum < 0.0;
      const double mom = !cma ? p.momentum
                              : ((p.num_batches_tracked && *p.num_batches_tracked > 0) ? 1.0 / (double)*p.num_batches_tracked : 0.0);
      p.running_mean[d] = (float)((1.0 - mom) * (double)p.running_mean[d]) + (float)(mom * mean_r);
      p.running_var[d] = (float)((1.0 - mom) * (double)p.running_var[d]) + (float)(mom * unbiased);
      if (!cma && d == 0 && p.num_batches_tracked) *p.num_batches_tracked += 1;
    }
    if (d == 0)
      for (int c = tid; c < C; c += kThreads) p.sx[c] = p.sx_sums[2 * c];
  } else {
    mean_r = (double)p.running_mean[d];
    var_r = (double)p.running_var[d];
  }
  const double invstd = 1.0 / sqrt(var_r + p.eps);
  const double sc = (double)p.gamma[d] * invstd;
  for (int c = tid; c < C; c += kThreads) p.Wf[(size_t)d * C + c] = (float)((double)p.Wd[(size_t)d * C + c] * sc);
  if (tid == 0) {
    p.bf[d] = (float)((double)p.beta[d] + sc * (bd - mean_r));
    p.mean_r[d] = mean_r;
    p.invstd[d] = invstd;
  }
  if (p.training) {   // the last block to finish hands the channel sums back zeroed
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(p.counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {
      for (int c = tid; c < 2 * C; c += kThreads) p.sx_sums[c] = 0.0;
      if (tid == 0) *p.counter = 0;
    }
  }
}

// one block per output channel d
__global__ void __launch_bounds__(kThreads) coeffs_kernel(const SgcnSideBwd p) {
  __shared__ double wrow[kMaxC], red[kThreads / 32];
  const int d = blockIdx.x, C = p.C, D = p.D, tid = threadIdx.x;
  const double bd = p.bd ? (double)p.bd[d] : 0.0;
  const double invstd = p.invstd[d], mean_r = p.mean_r[d], sg = (double)p.sg[d];
  for (int c = tid; c < C; c += kThreads) wrow[c] = (double)p.Wd[(size_t)d * C + c];
  __syncthreads();
  double a = 0.0, b = 0.0;
  for (int c = tid; c < C; c += kThreads) {
    a += wrow[c] * (double)p.P[(size_t)c * D + d];
    if (p.training) b += wrow[c] * p.sx[c];
  }
  const double dot = block_sum(a, red);
  const double wsx = block_sum(b, red);
  const double dgamma = invstd * (dot + (bd - mean_r) * sg);
  const double k = (double)p.gamma[d] * invstd;
  const double m1 = p.training ? sg / p.rows : 0.0, m2 = p.training ? dgamma / p.rows : 0.0;
  const double al = k, be = -k * m2 * invstd, ga = -k * m1 + k * m2 * invstd * mean_r;
  for (int c = tid; c < C; c += kThreads) {
    double wxx = 0.0, sxc = 0.0;
    if (p.training) {
      wxx = col_dot(wrow, p.XX, C, C, c);                         // (Wd[d] XX)[c]
      sxc = p.sx[c];
    }
    p.dWd[(size_t)d * C + c] = (float)(al * (double)p.P[(size_t)c * D + d] + be * (wxx + bd * sxc) + ga * sxc);
    p.Wcat[(size_t)d * C + c] = (float)(al * wrow[c]);
  }
  if (tid == 0) {
    p.dgamma[d] = (float)dgamma;
    p.dbeta[d] = (float)sg;
    p.dbd[d] = (float)(al * sg + be * (wsx + p.rows * bd) + p.rows * ga);
    p.coef[2 * d] = be;
    p.coef[2 * d + 1] = be * bd + ga;
  }
}

// one block per input channel c: Wcat[D + c][c2] = sum_d Wd[d,c] be[d] Wd[d,c2];  kvec[c] = sum_d Wd[d,c] (be*bd + ga)[d]
__global__ void __launch_bounds__(kThreads) mix_kernel(const SgcnSideBwd p) {
  __shared__ double wcol[kMaxC], red[kThreads / 32];
  const int c = blockIdx.x, C = p.C, D = p.D, tid = threadIdx.x;
  double kv = 0.0;
  for (int d = tid; d < D; d += kThreads) {
    const double w = (double)p.Wd[(size_t)d * C + c];
    wcol[d] = w * p.coef[2 * d];
    kv += w * p.coef[2 * d + 1];
  }
  __syncthreads();
  for (int c2 = tid; c2 < C; c2 += kThreads) {
    p.Wcat[(size_t)(D + c) * C + c2] = (float)col_dot(wcol, p.Wd, D, C, c2);
  }
  const double ks = block_sum(kv, red);
  if (tid == 0) p.kvec[c] = (float)ks;
}

static int check_cd(int C, int D) {
  if (C < 1 || C > kMaxC || D < 1 || D > kMaxC) return set_error("sgcn_side: channel counts must be in [1, 256]");
  return 0;
}

}  // namespace side
}  // namespace sgcn

using namespace sgcn;

extern "C" int sgcn_side_fold(const SgcnSideFold* p, void* stream) {
  if (!p || !p->Wd || !p->gamma || !p->beta || !p->Wf || !p->bf || !p->mean_r || !p->invstd)
    return set_error("sgcn_side_fold: null pointer");
  if (int rc = side::check_cd(p->C, p->D)) return rc;
  if (p->training) {
    if (!p->sx_sums || !p->XX || !p->sx || !p->counter || !(p->rows > 0.0))
      return set_error("sgcn_side_fold: batch statistics need sx_sums, XX, sx, counter and rows > 0");
  } else if (!p->running_mean || !p->running_var) {
    return set_error("sgcn_side_fold: inference needs the running statistics");
  }
  side::fold_kernel<<<p->D, side::kThreads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("side fold_kernel");
}

extern "C" int sgcn_side_bwd(const SgcnSideBwd* p, void* stream) {
  if (!p || !p->P || !p->sg || !p->Wd || !p->gamma || !p->invstd || !p->mean_r || !p->dgamma || !p->dbeta || !p->dWd ||
      !p->dbd || !p->Wcat || !p->kvec || !p->coef)
    return set_error("sgcn_side_bwd: null pointer");
  if (int rc = side::check_cd(p->C, p->D)) return rc;
  if (p->training && (!p->sx || !p->XX || !(p->rows > 0.0)))
    return set_error("sgcn_side_bwd: batch statistics need sx, XX and rows > 0");
  side::coeffs_kernel<<<p->D, side::kThreads, 0, (cudaStream_t)stream>>>(*p);
  if (int rc = check_launch("side coeffs_kernel")) return rc;
  side::mix_kernel<<<p->C, side::kThreads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("side mix_kernel");
}
