// pointwise.h -- parameter blocks of the SIMT kernels in pointwise.cu / prep.cu (plain C, mirrored by ctypes).
// All tensors are channels-last rows [(n, t, v), C] in fp32; reduction buffers are double and are ACCUMULATED
// (the finalize kernels clear them again), so a freshly zeroed workspace stays reusable step after step.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

/* forward temporal shift with stride: s = Shift(q); mode 0 -> stats[c] += {sum, sumsq}; mode 1 -> out = [relu](s*scale+shift+res) */
typedef struct SgcnTShift {
  const float* q;         /* [n, T_in, V, C]                       */
  const float* res;       /* [n, T_out, V, C] or NULL (mode 1)     */
  float* out;             /* [n, T_out, V, C] (mode 1)             */
  const float* ypos_eff;  /* [C] ypos (+0.5 when stride != 1)      */
  const float* scale;     /* [C] folded BN scale (mode 1)          */
  const float* shift;     /* [C] folded BN shift (mode 1)          */
  double* stats;          /* [C][2] (mode 0)                       */
  long long n_samples;
  int T_in, T_out, V, C, stride, relu;
} SgcnTShift;

/* backward of  out = [relu](BN(Shift(q)) + res):
 * mode 0: sums[c][5] += { g, g*shat, g*dq, dq, shat*dq };  mode 1: dpre = [q>0] * Shift^T(k1*(g - m1 - shat*m2)), dbias[c] += dpre */
typedef struct SgcnTShiftBwd {
  const float* q;
  const float* gy;        /* grad wrt out [n, T_out, V, C]                      */
  const float* y;         /* out (ReLU mask source), needed when relu != 0      */
  const float* ypos_eff;
  const float* mean;      /* [C] BN mean used in the forward                    */
  const float* invstd;    /* [C]                                                */
  const float* k1;        /* [C] gamma*invstd        (mode 1)                   */
  const float* m1;        /* [C] sum(g)/count  or 0  (mode 1)                   */
  const float* m2;        /* [C] sum(g*shat)/count or 0 (mode 1)                */
  double* sums;           /* [C][5] (mode 0)                                    */
  float* dpre;            /* [n, T_in, V, C] (mode 1)                           */
  double* dbias;          /* [C] (mode 1)                                       */
  long long n_samples;
  int T_in, T_out, V, C, stride, relu;
} SgcnTShiftBwd;

/* backward of  p = Shift_1(BN(h)):  du = Shift^T(dp)
 * mode 0: sums[c][3] += { du, du*hhat, dp*dU };
 * mode 1: gh = [h>0] * k1*(du - m1 - hhat*m2);  vd_sums[v,c][2] += { gh, gh*zhat } when z != NULL */
typedef struct SgcnTShiftInBwd {
  const float* dp;        /* grad wrt p [n, T, V, C]                            */
  const float* h;         /* tcn input (gcn output)                             */
  const float* z;         /* pre-BN spatial output or NULL                      */
  const float* ypos_eff;
  const float* mean;      /* [C] BN(h) statistics                               */
  const float* invstd;
  const float* scale;     /* [C] folded BN(h) scale / shift (mode 0)            */
  const float* shift;
  const float* k1;
  const float* m1;
  const float* m2;
  const float* zmean;     /* [V, C] BN1d statistics of z (mode 1, z != NULL)    */
  const float* zinvstd;
  double* sums;           /* [C][3] (mode 0)                                    */
  double* vd_sums;        /* [V, C][2] (mode 1)                                 */
  float* gh;              /* [n, T, V, C] (mode 1)                              */
  long long n_samples;
  int T, V, C, relu_h;
} SgcnTShiftInBwd;

int sgcn_bn_res_relu_fwd(const float* z, const float* res, float* h, const float* scale, const float* shift,
                         double* stats_out, long long rows, int V, int D, int relu, void* stream);
int sgcn_tshift_fwd(const SgcnTShift* p, int mode, void* stream);
int sgcn_tshift_bwd(const SgcnTShiftBwd* p, int mode, void* stream);
int sgcn_tshift_in_bwd(const SgcnTShiftInBwd* p, int mode, void* stream);

/* stats[c][2] += {sum, sumsq} over rows */
int sgcn_channel_stats(const float* x, double* stats, long long rows, int C, void* stream);
/* gh = g*[h>0]; vd_sums[v,c][2] += {gh, gh*zhat} */
int sgcn_relu_bn1d_bwd_stats(const float* g, const float* h, const float* z, const float* zmean, const float* zinvstd,
                             float* gh, double* vd_sums, long long groups, int V, int C, void* stream);
/* out = g*[y>0] */
int sgcn_relu_mask_grad(const float* g, const float* y, float* out, long long numel, void* stream);

/* ---- small per-feature kernels (prep.cu) ---- */

/* batch-norm forward finalize: {sum, sumsq} -> mean/invstd/scale/shift, running-stat update (momentum, unbiased var),
 * num_batches_tracked += 1, stats cleared.  training == 0: scale/shift from the running statistics, nothing updated. */
int sgcn_bn_fwd_finalize(double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                         long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                         int features, double count, double momentum, double eps, int training, void* stream);

/* backward finalize of the output shift + BN: from sums[c][5] -> dgamma, dbeta, k1, m1, m2 and the K5-constrained
 * position gradient (raw = k1*(S2 - m1*S3 - m2*S4) / n_batch).  raw_out (optional) receives the raw means. */
int sgcn_tshift_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma, float* dbeta,
                             float* k1, float* m1, float* m2, float* grad_xpos, float* grad_ypos, float* raw_out,
                             int C, double count, double n_batch, int training, void* stream);

/* backward finalize of BN + input shift: sums[c][3] -> dgamma, dbeta, k1, m1, m2, K5-constrained position gradient */
int sgcn_tshift_in_bwd_finalize(double* sums, const float* gamma, const float* invstd, float* dgamma, float* dbeta,
                                float* k1, float* m1, float* m2, float* grad_xpos, float* grad_ypos, float* raw_out,
                                int C, double count, double n_batch, int training, void* stream);

/* backward finalize of the BN1d over (v,d): vd_sums[f][2] -> dgamma, dbeta, alpha/beta/gamma tables of
 * dz = alpha*gh + beta*z + gamma, and the Linear_bias gradient dbias[d] = sum_v k*(S_g - count*m1). */
int sgcn_bn1d_bwd_finalize(double* vd_sums, const float* gamma, const float* mean, const float* invstd, float* dgamma,
                           float* dbeta, float* alpha, float* beta, float* gam, float* dbias, int V, int D,
                           double count, int training, void* stream);

/* maskmul = tanh(mask) + 1 */
int sgcn_mask_prepare(const float* mask, float* maskmul, int n, void* stream);
/* dmask = raw * (1 - tanh(mask)^2); raw cleared */
int sgcn_mask_grad_finalize(double* raw, const float* mask, float* dmask, int n, void* stream);

/* canonical (SWIZZLE_128B, TF32-rounded) image of B[n][k] = src[n*ld_n + k*ld_k], chunked by 64 k */
int sgcn_prep_weight_image(const float* src, long long ld_n, long long ld_k, int N, int K, float* image, void* stream);

/* double -> float copy of a reduction buffer (+ clear) */
int sgcn_reduce_export(double* src, float* dst, int n, double scale, void* stream);

#ifdef __cplusplus
}
#endif
