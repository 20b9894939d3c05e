// prep.cu -- small per-feature kernels: BatchNorm finalisation (forward and backward), the K5 shift-position
// constraint applied to in-kernel reduced sums, mask / weight preparation.  One thread per feature; all
// arithmetic in double, results stored as fp32.
#include <math.h>

#include "capi_internal.h"
#include "common.cuh"
#include "pointwise.h"

namespace sgcn {

// BatchNorm (train): torch semantics -- biased variance for normalisation, unbiased for the running estimate,
// running = (1-momentum)*running + momentum*batch  (model/shift_gcn.py:55-56,99; nn.BatchNorm defaults).
__global__ void bn_fwd_finalize_kernel(double* __restrict__ stats, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, long long* __restrict__ nbt,
                                       float* __restrict__ mean_o, float* __restrict__ invstd_o,
                                       float* __restrict__ scale_o, float* __restrict__ shift_o, int F, double count,
                                       double momentum, double eps, int training) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  // momentum < 0 = nn.BatchNorm(momentum=None): cumulative moving average with factor 1 / num_batches_tracked; the
  // caller has already incremented the counter (every thread reads it here, so it cannot also be written here)
  if (momentum < 0.0) momentum = (training && nbt && *nbt > 0) ? 1.0 / (double)*nbt : 0.0;
  else if (f == 0 && training && nbt) *nbt += 1;
  if (f >= F) return;
  double mean, var;
  if (training) {
    mean = stats[2 * f] / count;
    var = stats[2 * f + 1] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * f] = 0.0;
    stats[2 * f + 1] = 0.0;
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[f] = (float)((1.0 - momentum) * (double)running_mean[f] + momentum * mean);
      running_var[f] = (float)((1.0 - momentum) * (double)running_var[f] + momentum * unbiased);
    }
  } else {
    mean = (double)running_mean[f];
    var = (double)running_var[f];
  }
  const double invstd = 1.0 / sqrt(var + eps);
  const double sc = (gamma ? (double)gamma[f] : 1.0) * invstd;
  mean_o[f] = (float)mean;
  invstd_o[f] = (float)invstd;
  scale_o[f] = (float)sc;
  shift_o[f] = (float)((beta ? (double)beta[f] : 0.0) - mean * sc);
}

// K5 (shift_cuda_kernel.cu:371-395) on one reduced value
__device__ __forceinline__ void shift_constraint(float raw_y, float* gx, float* gy) {
  const float dr = sqrtf(raw_y * raw_y);
  if (dr != 0.f) {
    *gx = 0.f;
    *gy = (float)((double)(raw_y / dr) * 0.01);
  } else {
    *gx = 0.f;
    *gy = (float)0.0001;
  }
}

__global__ void tshift_bwd_finalize_kernel(double* __restrict__ sums, const float* __restrict__ gamma,
                                           const float* __restrict__ invstd, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta, float* __restrict__ k1o, float* __restrict__ m1o,
                                           float* __restrict__ m2o, float* __restrict__ gx, float* __restrict__ gy,
                                           float* __restrict__ raw_out, int C, double count, double n_batch,
                                           int training, int nsum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < nsum; ++i) {
    S[i] = sums[(size_t)c * nsum + i];
    sums[(size_t)c * nsum + i] = 0.0;
  }
  const double k1 = (double)gamma[c] * (double)invstd[c];
  const double m1 = training ? S[0] / count : 0.0, m2 = training ? S[1] / count : 0.0;
  dbeta[c] = (float)S[0];
  dgamma[c] = (float)S[1];
  k1o[c] = (float)k1;
  m1o[c] = (float)m1;
  m2o[c] = (float)m2;
  // nsum == 5: output shift (S2 = sum g*dq, S3 = sum dq, S4 = sum shat*dq); nsum == 3: input shift (S2 = sum dp*dU)
  const double raw = nsum == 5 ? k1 * (S[2] - m1 * S[3] - m2 * S[4]) / n_batch : S[2] / n_batch;
  if (raw_out) raw_out[c] = (float)raw;
  shift_constraint((float)raw, gx + c, gy + c);
}

__global__ void shift_pos_finalize_kernel(double* __restrict__ pos, float* __restrict__ gx, float* __restrict__ gy,
                                          float* __restrict__ raw_out, int C, double n_batch) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double raw = pos[c] / n_batch;
  pos[c] = 0.0;
  if (raw_out) raw_out[c] = (float)raw;
  shift_constraint((float)raw, gx + c, gy + c);
}

__global__ void bn1d_bwd_finalize_kernel(double* __restrict__ vd, const float* __restrict__ gamma,
                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ gam,
                                         float* __restrict__ dbias, int V, int D, double count, int training) {
  // block = 32 channels x 8 joint lanes: the V joints of a channel are spread over threadIdx.y (the loop used to be
  // serial per channel: V dependent round trips to L2 for a [V*D] table)
  __shared__ double part[8][32];
  const int d = blockIdx.x * 32 + threadIdx.x;
  double db = 0.0;
  if (d < D)
    for (int v = threadIdx.y; v < V; v += 8) {
      const size_t f = (size_t)v * D + d;
      const double S0 = vd[2 * f], S1 = vd[2 * f + 1];
      vd[2 * f] = 0.0;
      vd[2 * f + 1] = 0.0;
      const double is = (double)invstd[f], k = (double)gamma[f] * is;
      const double m1 = training ? S0 / count : 0.0, m2 = training ? S1 / count : 0.0;
      dbeta[f] = (float)S0;
      dgamma[f] = (float)S1;
      alpha[f] = (float)k;
      beta[f] = (float)(-k * m2 * is);
      gam[f] = (float)(-k * m1 + k * m2 * is * (double)mean[f]);
      db += k * (S0 - count * m1);
    }
  part[threadIdx.y][threadIdx.x] = db;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x];
    dbias[d] = (float)s;
  }
}

__global__ void mask_prepare_kernel(const float* __restrict__ mask, float* __restrict__ mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = tanhf(mask[i]) + 1.f;
}

__global__ void mask_prepare_rot_kernel(const float* __restrict__ mask, float* __restrict__ mm, float* __restrict__ mm_rot,
                                        int V, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * C) return;
  const int u = i / C, c = i - u * C;
  const float m = tanhf(mask[i]) + 1.f;
  mm[i] = m;
  mm_rot[((u + c) % V) * C + c] = m;                       // source joint (u + c) mod V feeds operand row u
}

__global__ void mask_grad_finalize_kernel(double* __restrict__ raw, const float* __restrict__ mask,
                                          float* __restrict__ dmask, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = tanh((double)mask[i]);
  dmask[i] = (float)(raw[i] * (1.0 - t * t));
  raw[i] = 0.0;
}

__global__ void prep_weight_image_kernel(const float* __restrict__ src, long long ld_n, long long ld_k, int N, int K,
                                         float* __restrict__ image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const size_t off = (size_t)(k >> 6) * N * 256 + (size_t)((k >> 5) & 1) * N * 128 + canon_off(n, k & 31);
  *(float*)((uint8_t*)image + off) = to_tf32(src[(size_t)n * ld_n + (size_t)k * ld_k]);
}

__global__ void prep_weight_image_split_kernel(const float* __restrict__ src, long long ld_n, long long ld_k, int N, int K,
                                               float* __restrict__ image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const size_t off = (size_t)(k >> 6) * N * 256 + (size_t)((k >> 5) & 1) * N * 128 + canon_off(n, k & 31);
  const float w = src[(size_t)n * ld_n + (size_t)k * ld_k];
  const float hi = to_tf32(w);
  *(float*)((uint8_t*)image + off) = hi;
  *(float*)((uint8_t*)image + (size_t)N * K * 4 + off) = to_tf32(w - hi);
}

__global__ void reduce_export_kernel(double* __restrict__ src, float* __restrict__ dst, int n, double scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = (float)(src[i] * scale);
  src[i] = 0.0;
}

}  // namespace sgcn

using namespace sgcn;

extern "C" int sgcn_bn_fwd_finalize(double* stats, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, long long* num_batches_tracked, float* mean, float* invstd,
                                    float* scale, float* shift, int features, double count, double momentum, double eps,
                                    int training, void* stream) {
  if (!mean || !invstd || !scale || !shift) return set_error("sgcn_bn_fwd_finalize: null output");
  if (training && !stats) return set_error("sgcn_bn_fwd_finalize: training needs stats");
  if (!training && (!running_mean || !running_var)) return set_error("sgcn_bn_fwd_finalize: eval needs running stats");
  if (features <= 0) return 0;
  bn_fwd_finalize_kernel<<<(features + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      stats, gamma, beta, running_mean, running_var, num_batches_tracked, mean, invstd, scale, shift, features, count,
      momentum, eps, training);
  return check_launch("bn_fwd_finalize_kernel");
}

extern "C" int sgcn_tshift_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma,
                                        float* dbeta, float* k1, float* m1, float* m2, float* grad_xpos,
                                        float* grad_ypos, float* raw_out, int C, double count, double n_batch,
                                        int training, void* stream) {
  if (!sums || !gamma || !invstd || !dgamma || !dbeta || !k1 || !m1 || !m2 || !grad_xpos || !grad_ypos)
    return set_error("sgcn_tshift_bwd_finalize: null pointer");
  tshift_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sums, gamma, invstd, dgamma, dbeta, k1, m1, m2, grad_xpos, grad_ypos, raw_out, C, count, n_batch, training, 5);
  return check_launch("tshift_bwd_finalize_kernel");
}

extern "C" int sgcn_tshift_in_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma,
                                           float* dbeta, float* k1, float* m1, float* m2, float* grad_xpos,
                                           float* grad_ypos, float* raw_out, int C, double count, double n_batch,
                                           int training, void* stream) {
  if (!sums || !gamma || !invstd || !dgamma || !dbeta || !k1 || !m1 || !m2 || !grad_xpos || !grad_ypos)
    return set_error("sgcn_tshift_in_bwd_finalize: null pointer");
  tshift_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sums, gamma, invstd, dgamma, dbeta, k1, m1, m2, grad_xpos, grad_ypos, raw_out, C, count, n_batch, training, 3);
  return check_launch("tshift_in_bwd_finalize_kernel");
}

extern "C" int sgcn_shift_pos_finalize(double* pos_sums, float* grad_xpos, float* grad_ypos, float* raw_out, int C,
                                       double n_batch, void* stream) {
  if (!pos_sums || !grad_xpos || !grad_ypos) return set_error("sgcn_shift_pos_finalize: null pointer");
  shift_pos_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pos_sums, grad_xpos, grad_ypos, raw_out, C,
                                                                              n_batch);
  return check_launch("shift_pos_finalize_kernel");
}

extern "C" int sgcn_bn1d_bwd_finalize(double* vd_sums, const float* gamma, const float* mean, const float* invstd,
                                      float* dgamma, float* dbeta, float* alpha, float* beta, float* gam, float* dbias,
                                      int V, int D, double count, int training, void* stream) {
  if (!vd_sums || !gamma || !mean || !invstd || !dgamma || !dbeta || !alpha || !beta || !gam || !dbias)
    return set_error("sgcn_bn1d_bwd_finalize: null pointer");
  bn1d_bwd_finalize_kernel<<<(D + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(vd_sums, gamma, mean, invstd, dgamma, dbeta,
                                                                          alpha, beta, gam, dbias, V, D, count,
                                                                          training);
  return check_launch("bn1d_bwd_finalize_kernel");
}

extern "C" int sgcn_mask_prepare(const float* mask, float* maskmul, int n, void* stream) {
  if (!mask || !maskmul) return set_error("sgcn_mask_prepare: null pointer");
  mask_prepare_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mask, maskmul, n);
  return check_launch("mask_prepare_kernel");
}

extern "C" int sgcn_mask_prepare_rot(const float* mask, float* maskmul, float* maskmul_rot, int V, int C, void* stream) {
  if (!mask || !maskmul || !maskmul_rot) return set_error("sgcn_mask_prepare_rot: null pointer");
  if (V < 1 || C < 1) return set_error("sgcn_mask_prepare_rot: bad shape");
  mask_prepare_rot_kernel<<<(V * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mask, maskmul, maskmul_rot, V, C);
  return check_launch("mask_prepare_rot_kernel");
}

extern "C" int sgcn_mask_grad_finalize(double* raw, const float* mask, float* dmask, int n, void* stream) {
  if (!raw || !mask || !dmask) return set_error("sgcn_mask_grad_finalize: null pointer");
  mask_grad_finalize_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(raw, mask, dmask, n);
  return check_launch("mask_grad_finalize_kernel");
}

extern "C" int sgcn_prep_weight_image(const float* src, long long ld_n, long long ld_k, int N, int K, float* image,
                                      void* stream) {
  if (!src || !image) return set_error("sgcn_prep_weight_image: null pointer");
  if (N % 8 != 0 || K % 64 != 0) return set_error("sgcn_prep_weight_image: need N % 8 == 0 and K % 64 == 0");
  prep_weight_image_kernel<<<(N * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, ld_n, ld_k, N, K, image);
  return check_launch("prep_weight_image_kernel");
}

extern "C" int sgcn_prep_weight_image_split(const float* src, long long ld_n, long long ld_k, int N, int K, float* image,
                                            void* stream) {
  if (!src || !image) return set_error("sgcn_prep_weight_image_split: null pointer");
  if (N % 8 != 0 || K % 64 != 0) return set_error("sgcn_prep_weight_image_split: need N % 8 == 0 and K % 64 == 0");
  prep_weight_image_split_kernel<<<(N * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, ld_n, ld_k, N, K, image);
  return check_launch("prep_weight_image_split_kernel");
}

extern "C" int sgcn_reduce_export(double* src, float* dst, int n, double scale, void* stream) {
  if (!src || !dst) return set_error("sgcn_reduce_export: null pointer");
  reduce_export_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, dst, n, scale);
  return check_launch("reduce_export_kernel");
}
