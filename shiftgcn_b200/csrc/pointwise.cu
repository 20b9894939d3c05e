// pointwise.cu -- bandwidth-bound SIMT kernels of the Shift-GCN hot path (channels-last rows [(n,t,v), C]).
//
// Thread mapping: threadIdx % C <-> channel (coalesced 128-byte warp accesses), threadIdx / C <-> row slot.
// Cross-row reductions (BatchNorm statistics, position / bias gradients) are accumulated in registers,
// combined across the block's row slots in shared memory and published with one double atomic per value.
//
//   bn_res_relu_fwd     h = relu(BN1d(z) + res)  (+ per-channel stats of h)           model/shift_gcn.py:137-141, :66
//   tshift_fwd          s = Shift_stride(q); stats of s  |  y = [relu](BN(s) + res)    :72-73, :161-162, K1
//   tshift_bwd_stats    BN backward sums + position-gradient sums of shift_out          autograd of :72-73, K4
//   tshift_bwd_apply    dpre = [q>0] * Shift^T(BN-bwd(g))  (+ conv-bias gradient)       K2/K3, autograd of :70-73
//   tshift_in_bwd_stats du = Shift^T(dp); BN backward sums; position-gradient sums      K2, K4, autograd of :66-68
//   tshift_in_bwd_apply gh = [h>0] * BN-bwd(du)  (+ per-(v,d) BN1d backward sums)        autograd of :66, :137-141
#include "capi_internal.h"
#include "common.cuh"
#include "pointwise.h"

namespace sgcn {

constexpr int kPwThreads = 256;

// combine per-thread partials over the row slots of a block, then one double atomic per channel
template <int NV>
__device__ __forceinline__ void block_reduce_channels(float (&v)[NV], double* __restrict__ dst, int dst_stride, int C,
                                                      float* scratch /* [kPwThreads * NV] */) {
  const int tid = threadIdx.x;
  const int slots = kPwThreads / C;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) scratch[k * kPwThreads + tid] = v[k];
  __syncthreads();
  if (tid < C) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int sl = 0; sl < slots; ++sl) s += (double)scratch[k * kPwThreads + sl * C + tid];
      atomicAdd(dst + (size_t)tid * dst_stride + k, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// h = relu(z * sc[v,d] + sh[v,d] + res)          stats_out[d] += {sum h, sum h^2}
__global__ void __launch_bounds__(kPwThreads) bn_res_relu_fwd_kernel(const float* __restrict__ z,
                                                                     const float* __restrict__ res,
                                                                     float* __restrict__ h,
                                                                     const float* __restrict__ sc,
                                                                     const float* __restrict__ sh,
                                                                     double* __restrict__ stats_out, long long rows,
                                                                     int V, int D, int relu) {
  __shared__ float scratch[kPwThreads * 2];
  const int tid = threadIdx.x, d = tid % D, slot = tid / D, slots = kPwThreads / D;
  float acc[2] = {0.f, 0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const int v = (int)(r % V);
    const size_t o = (size_t)r * D + d;
    float y = fmaf(z[o], __ldg(sc + v * D + d), __ldg(sh + v * D + d));
    if (res) y += res[o];
    if (relu) y = fmaxf(y, 0.f);
    h[o] = y;
    acc[0] += y;
    acc[1] = fmaf(y, y, acc[1]);
  }
  if (stats_out) block_reduce_channels<2>(acc, stats_out, 2, D, scratch);
}

// ------------------------------------------------------------------------------------------------ temporal shift
struct LerpCh {
  int y1;
  float fy;
};
__device__ __forceinline__ LerpCh lerp_of(float ypos_eff) {
  const float f = floorf(ypos_eff);
  return {(int)f, ypos_eff - f};
}

// value of the zero-padded row `t` of sample-plane `plane` ((n, t=0) row base), channel offset included in ptr
__device__ __forceinline__ float tap_row(const float* __restrict__ base, int t, int T, size_t row_pitch) {
  return (t >= 0 && t < T) ? __ldg(base + (size_t)t * row_pitch) : 0.f;
}

// MODE 0: stats[c] += {sum s, sum s^2};  MODE 1: out = [relu](s*sc + sh + res)
template <int MODE>
__global__ void __launch_bounds__(kPwThreads) tshift_fwd_kernel(const SgcnTShift p) {
  __shared__ float scratch[kPwThreads * 2];
  const int C = p.C, V = p.V, tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  const LerpCh L = lerp_of(p.ypos_eff[c]);
  const float sc = MODE == 1 ? p.scale[c] : 0.f, sh = MODE == 1 ? p.shift[c] : 0.f;
  const size_t pitch = (size_t)V * C;
  const long long rows = p.n_samples * p.T_out * V;
  float acc[2] = {0.f, 0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const long long grp = r / V;
    const int v = (int)(r - grp * V);
    const long long n = grp / p.T_out;
    const int to = (int)(grp - n * p.T_out);
    const float* base = p.q + ((size_t)n * p.T_in * V + v) * C + c;
    const int ta = to * p.stride + L.y1;
    const float s = tap_row(base, ta, p.T_in, pitch) * (1.f - L.fy) + tap_row(base, ta + 1, p.T_in, pitch) * L.fy;
    if (MODE == 0) {
      acc[0] += s;
      acc[1] = fmaf(s, s, acc[1]);
    } else {
      const size_t o = (size_t)r * C + c;
      float y = fmaf(s, sc, sh);
      if (p.res) y += p.res[o];
      p.out[o] = p.relu ? fmaxf(y, 0.f) : y;
    }
  }
  if (MODE == 0) block_reduce_channels<2>(acc, p.stats, 2, C, scratch);
}

// sums5[c] += { g, g*shat, g*dq, dq, shat*dq }   with g = gy*[y>0] (if relu), s = Shift(q), shat = (s-mean)*invstd,
// dq = Q(ta+1) - Q(ta)  (d s / d ypos, K4 with xpos = 0)
__global__ void __launch_bounds__(kPwThreads) tshift_bwd_stats_kernel(const SgcnTShiftBwd p) {
  __shared__ float scratch[kPwThreads * 5];
  const int C = p.C, V = p.V, tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  const LerpCh L = lerp_of(p.ypos_eff[c]);
  const float mean = p.mean[c], invstd = p.invstd[c];
  const size_t pitch = (size_t)V * C;
  const long long rows = p.n_samples * p.T_out * V;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const long long grp = r / V;
    const int v = (int)(r - grp * V);
    const long long n = grp / p.T_out;
    const int to = (int)(grp - n * p.T_out);
    const float* base = p.q + ((size_t)n * p.T_in * V + v) * C + c;
    const int ta = to * p.stride + L.y1;
    const float q0 = tap_row(base, ta, p.T_in, pitch), q1 = tap_row(base, ta + 1, p.T_in, pitch);
    const float s = q0 * (1.f - L.fy) + q1 * L.fy;
    const float shat = (s - mean) * invstd;
    const float dq = q1 - q0;
    const size_t o = (size_t)r * C + c;
    float g = p.gy[o];
    if (p.relu && !(p.y[o] > 0.f)) g = 0.f;
    acc[0] += g;
    acc[1] = fmaf(g, shat, acc[1]);
    acc[2] = fmaf(g, dq, acc[2]);
    acc[3] += dq;
    acc[4] = fmaf(shat, dq, acc[4]);
  }
  block_reduce_channels<5>(acc, p.sums, 5, C, scratch);
}

// ds(n,t',v,c) = k1*(g - m1 - shat*m2)
__device__ __forceinline__ float ds_at(const SgcnTShiftBwd& p, const float* __restrict__ qbase, long long n, int to,
                                       int v, int c, const LerpCh& L, float mean, float invstd, float k1, float m1,
                                       float m2, size_t pitch) {
  const size_t o = (((size_t)n * p.T_out + to) * p.V + v) * p.C + c;
  float g = __ldg(p.gy + o);
  if (p.relu && !(__ldg(p.y + o) > 0.f)) g = 0.f;
  const int ta = to * p.stride + L.y1;
  const float s = tap_row(qbase, ta, p.T_in, pitch) * (1.f - L.fy) + tap_row(qbase, ta + 1, p.T_in, pitch) * L.fy;
  return k1 * (g - m1 - (s - mean) * invstd * m2);
}

// dpre(n,t,v,c) = [q > 0] * ( (1-fy)*ds[(t-y1)/stride] + fy*ds[(t-y1-1)/stride] )   (taps exist only when divisible)
// dbias[c] += dpre
__global__ void __launch_bounds__(kPwThreads) tshift_bwd_apply_kernel(const SgcnTShiftBwd p) {
  __shared__ float scratch[kPwThreads];
  const int C = p.C, V = p.V, tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  const LerpCh L = lerp_of(p.ypos_eff[c]);
  const float mean = p.mean[c], invstd = p.invstd[c], k1 = p.k1[c], m1 = p.m1[c], m2 = p.m2[c];
  const size_t pitch = (size_t)V * C;
  const long long rows = p.n_samples * p.T_in * V;
  const int st = p.stride;
  float acc[1] = {0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const long long grp = r / V;
    const int v = (int)(r - grp * V);
    const long long n = grp / p.T_in;
    const int t = (int)(grp - n * p.T_in);
    const float* qbase = p.q + ((size_t)n * p.T_in * V + v) * C + c;
    const size_t o = (size_t)r * C + c;
    float d = 0.f;
    if (p.q[o] > 0.f) {
      const int r0 = t - L.y1, r1 = t - L.y1 - 1;     // output-row coordinates times stride
      if (r0 >= 0 && r0 % st == 0 && r0 / st < p.T_out)
        d = (1.f - L.fy) * ds_at(p, qbase, n, r0 / st, v, c, L, mean, invstd, k1, m1, m2, pitch);
      if (r1 >= 0 && r1 % st == 0 && r1 / st < p.T_out)
        d = fmaf(L.fy, ds_at(p, qbase, n, r1 / st, v, c, L, mean, invstd, k1, m1, m2, pitch), d);
    }
    p.dpre[o] = d;
    acc[0] += d;
  }
  block_reduce_channels<1>(acc, p.dbias, 1, C, scratch);
}

// du = Shift_1^T(dp);  sums3[c] += { du, du*hhat, dp * dU }  with dU = U(t+y1+1) - U(t+y1), U = BN(h) zero padded
__global__ void __launch_bounds__(kPwThreads) tshift_in_bwd_stats_kernel(const SgcnTShiftInBwd p) {
  __shared__ float scratch[kPwThreads * 3];
  const int C = p.C, V = p.V, T = p.T, tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  const LerpCh L = lerp_of(p.ypos_eff[c]);
  const float mean = p.mean[c], invstd = p.invstd[c], sc = p.scale[c], sh = p.shift[c];
  const size_t pitch = (size_t)V * C;
  const long long rows = p.n_samples * T * V;
  float acc[3] = {0.f, 0.f, 0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const long long grp = r / V;
    const int v = (int)(r - grp * V);
    const long long n = grp / T;
    const int t = (int)(grp - n * T);
    const size_t pb = ((size_t)n * T * V + v) * C + c;
    const float* dpb = p.dp + pb;
    const float* hb = p.h + pb;
    const float du = (1.f - L.fy) * tap_row(dpb, t - L.y1, T, pitch) + L.fy * tap_row(dpb, t - L.y1 - 1, T, pitch);
    const size_t o = (size_t)r * C + c;
    const float hv = p.h[o];
    const int ta = t + L.y1;
    const float u0 = (ta >= 0 && ta < T) ? fmaf(sc, __ldg(hb + (size_t)ta * pitch), sh) : 0.f;
    const float u1 = (ta + 1 >= 0 && ta + 1 < T) ? fmaf(sc, __ldg(hb + (size_t)(ta + 1) * pitch), sh) : 0.f;
    acc[0] += du;
    acc[1] = fmaf(du, (hv - mean) * invstd, acc[1]);
    acc[2] = fmaf(p.dp[o], u1 - u0, acc[2]);
  }
  block_reduce_channels<3>(acc, p.sums, 3, C, scratch);
}

// gh = [h > 0] * k*(du - m1 - hhat*m2);   per-(v,d) sums for the BN1d backward: { gh, gh * zhat }
__global__ void __launch_bounds__(kPwThreads) tshift_in_bwd_apply_kernel(const SgcnTShiftInBwd p) {
  const int C = p.C, V = p.V, T = p.T, tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  const LerpCh L = lerp_of(p.ypos_eff[c]);
  const float mean = p.mean[c], invstd = p.invstd[c], k = p.k1[c], m1 = p.m1[c], m2 = p.m2[c];
  const size_t pitch = (size_t)V * C;
  const long long groups = p.n_samples * T;
  // joints are the outer loop so that the per-(v,c) sums live in two registers
  for (int v = slot; v < V; v += slots) {
    const float zmean = p.z ? __ldg(p.zmean + v * C + c) : 0.f, zinv = p.z ? __ldg(p.zinvstd + v * C + c) : 0.f;
    float s0 = 0.f, s1 = 0.f;
    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
      const long long n = grp / T;
      const int t = (int)(grp - n * T);
      const float* dpb = p.dp + ((size_t)n * T * V + v) * C + c;
      const float du = (1.f - L.fy) * tap_row(dpb, t - L.y1, T, pitch) + L.fy * tap_row(dpb, t - L.y1 - 1, T, pitch);
      const size_t o = ((size_t)grp * V + v) * C + c;
      const float hv = p.h[o];
      float g = k * (du - m1 - (hv - mean) * invstd * m2);
      if (p.relu_h && !(hv > 0.f)) g = 0.f;
      p.gh[o] = g;
      if (p.z) {
        s0 += g;
        s1 = fmaf(g, (p.z[o] - zmean) * zinv, s1);
      }
    }
    if (p.z) {
      atomicAdd(p.vd_sums + 2 * ((size_t)v * C + c), (double)s0);
      atomicAdd(p.vd_sums + 2 * ((size_t)v * C + c) + 1, (double)s1);
    }
  }
}

// stats[c] += { sum x, sum x^2 }   (BatchNorm2d statistics of a stand-alone Shift_tcn input, model/shift_gcn.py:66)
__global__ void __launch_bounds__(kPwThreads) channel_stats_kernel(const float* __restrict__ x,
                                                                   double* __restrict__ stats, long long rows, int C) {
  __shared__ float scratch[kPwThreads * 2];
  const int tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  float acc[2] = {0.f, 0.f};
  for (long long r = (long long)blockIdx.x * slots + slot; r < rows; r += (long long)gridDim.x * slots) {
    const float v = x[(size_t)r * C + c];
    acc[0] += v;
    acc[1] = fmaf(v, v, acc[1]);
  }
  block_reduce_channels<2>(acc, stats, 2, C, scratch);
}

// gh = g * [h > 0];  vd_sums[v,c] += { gh, gh * zhat }      (stand-alone Shift_gcn backward, model/shift_gcn.py:137-141)
__global__ void __launch_bounds__(kPwThreads) relu_bn1d_bwd_stats_kernel(const float* __restrict__ g,
                                                                         const float* __restrict__ h,
                                                                         const float* __restrict__ z,
                                                                         const float* __restrict__ zmean,
                                                                         const float* __restrict__ zinvstd,
                                                                         float* __restrict__ gh,
                                                                         double* __restrict__ vd_sums, long long groups,
                                                                         int V, int C) {
  const int tid = threadIdx.x, c = tid % C, slot = tid / C, slots = kPwThreads / C;
  for (int v = slot; v < V; v += slots) {
    const float zm = __ldg(zmean + v * C + c), zi = __ldg(zinvstd + v * C + c);
    float s0 = 0.f, s1 = 0.f;
    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
      const size_t o = ((size_t)grp * V + v) * C + c;
      const float gv = (h[o] > 0.f) ? g[o] : 0.f;
      gh[o] = gv;
      s0 += gv;
      s1 = fmaf(gv, (z[o] - zm) * zi, s1);
    }
    atomicAdd(vd_sums + 2 * ((size_t)v * C + c), (double)s0);
    atomicAdd(vd_sums + 2 * ((size_t)v * C + c) + 1, (double)s1);
  }
}

// out = g * [y > 0]
__global__ void __launch_bounds__(kPwThreads) relu_mask_grad_kernel(const float4* __restrict__ g,
                                                                    const float4* __restrict__ y,
                                                                    float4* __restrict__ out, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = g[i], b = y[i];
    out[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
  }
}

static int pw_grid(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int check_c(int C) {
  if (C != 64 && C != 128 && C != 256) return set_error("pointwise: channel count must be 64, 128 or 256");
  return 0;
}

}  // namespace sgcn

using namespace sgcn;

extern "C" int sgcn_bn_res_relu_fwd(const float* z, const float* res, float* h, const float* scale, const float* shift,
                                    double* stats_out, long long rows, int V, int D, int relu, void* stream) {
  if (!z || !h || !scale || !shift) return set_error("sgcn_bn_res_relu_fwd: null pointer");
  if (int rc = check_c(D)) return rc;
  if (rows <= 0) return 0;
  bn_res_relu_fwd_kernel<<<pw_grid(rows, kPwThreads / D * 8), kPwThreads, 0, (cudaStream_t)stream>>>(
      z, res, h, scale, shift, stats_out, rows, V, D, relu);
  return check_launch("bn_res_relu_fwd_kernel");
}

extern "C" int sgcn_tshift_fwd(const SgcnTShift* p, int mode, void* stream) {
  if (!p || !p->q || !p->ypos_eff) return set_error("sgcn_tshift_fwd: null pointer");
  if (int rc = check_c(p->C)) return rc;
  if (p->stride < 1 || p->T_out != p->T_in / p->stride) return set_error("sgcn_tshift_fwd: T_out must be T_in / stride");
  const long long rows = p->n_samples * p->T_out * p->V;
  if (rows <= 0) return 0;
  const int grid = pw_grid(rows, kPwThreads / p->C * 8);
  if (mode == 0) {
    if (!p->stats) return set_error("sgcn_tshift_fwd(stats): null stats");
    tshift_fwd_kernel<0><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>(*p);
  } else {
    if (!p->out || !p->scale || !p->shift) return set_error("sgcn_tshift_fwd(apply): null pointer");
    tshift_fwd_kernel<1><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>(*p);
  }
  return check_launch("tshift_fwd_kernel");
}

extern "C" int sgcn_tshift_bwd(const SgcnTShiftBwd* p, int mode, void* stream) {
  if (!p || !p->q || !p->gy || !p->ypos_eff || !p->mean || !p->invstd) return set_error("sgcn_tshift_bwd: null pointer");
  if (p->relu && !p->y) return set_error("sgcn_tshift_bwd: relu mask needs y");
  if (int rc = check_c(p->C)) return rc;
  if (p->stride < 1 || p->T_out != p->T_in / p->stride) return set_error("sgcn_tshift_bwd: T_out must be T_in / stride");
  if (mode == 0) {
    if (!p->sums) return set_error("sgcn_tshift_bwd(stats): null sums");
    const long long rows = p->n_samples * p->T_out * p->V;
    if (rows <= 0) return 0;
    tshift_bwd_stats_kernel<<<pw_grid(rows, kPwThreads / p->C * 8), kPwThreads, 0, (cudaStream_t)stream>>>(*p);
    return check_launch("tshift_bwd_stats_kernel");
  }
  if (!p->dpre || !p->dbias || !p->k1 || !p->m1 || !p->m2) return set_error("sgcn_tshift_bwd(apply): null pointer");
  const long long rows = p->n_samples * p->T_in * p->V;
  if (rows <= 0) return 0;
  tshift_bwd_apply_kernel<<<pw_grid(rows, kPwThreads / p->C * 8), kPwThreads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("tshift_bwd_apply_kernel");
}

extern "C" int sgcn_tshift_in_bwd(const SgcnTShiftInBwd* p, int mode, void* stream) {
  if (!p || !p->dp || !p->h || !p->ypos_eff || !p->mean || !p->invstd) return set_error("sgcn_tshift_in_bwd: null pointer");
  if (int rc = check_c(p->C)) return rc;
  const long long rows = p->n_samples * p->T * p->V;
  if (rows <= 0) return 0;
  if (mode == 0) {
    if (!p->sums || !p->scale || !p->shift) return set_error("sgcn_tshift_in_bwd(stats): null pointer");
    tshift_in_bwd_stats_kernel<<<pw_grid(rows, kPwThreads / p->C * 8), kPwThreads, 0, (cudaStream_t)stream>>>(*p);
    return check_launch("tshift_in_bwd_stats_kernel");
  }
  if (!p->gh || !p->k1 || !p->m1 || !p->m2) return set_error("sgcn_tshift_in_bwd(apply): null pointer");
  if (p->z && (!p->zmean || !p->zinvstd || !p->vd_sums)) return set_error("sgcn_tshift_in_bwd(apply): null BN1d tables");
  const long long groups = p->n_samples * p->T;
  long long grid = groups < (long long)num_sms() * 4 ? groups : (long long)num_sms() * 4;
  tshift_in_bwd_apply_kernel<<<(unsigned)grid, kPwThreads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("tshift_in_bwd_apply_kernel");
}

extern "C" int sgcn_channel_stats(const float* x, double* stats, long long rows, int C, void* stream) {
  if (!x || !stats) return set_error("sgcn_channel_stats: null pointer");
  if (int rc = check_c(C)) return rc;
  if (rows <= 0) return 0;
  channel_stats_kernel<<<pw_grid(rows, kPwThreads / C * 8), kPwThreads, 0, (cudaStream_t)stream>>>(x, stats, rows, C);
  return check_launch("channel_stats_kernel");
}

extern "C" int sgcn_relu_bn1d_bwd_stats(const float* g, const float* h, const float* z, const float* zmean,
                                        const float* zinvstd, float* gh, double* vd_sums, long long groups, int V,
                                        int C, void* stream) {
  if (!g || !h || !z || !zmean || !zinvstd || !gh || !vd_sums) return set_error("sgcn_relu_bn1d_bwd_stats: null pointer");
  if (int rc = check_c(C)) return rc;
  if (groups <= 0) return 0;
  long long grid = groups < (long long)num_sms() * 4 ? groups : (long long)num_sms() * 4;
  relu_bn1d_bwd_stats_kernel<<<(unsigned)grid, kPwThreads, 0, (cudaStream_t)stream>>>(g, h, z, zmean, zinvstd, gh,
                                                                                     vd_sums, groups, V, C);
  return check_launch("relu_bn1d_bwd_stats_kernel");
}

extern "C" int sgcn_relu_mask_grad(const float* g, const float* y, float* out, long long numel, void* stream) {
  if (!g || !y || !out) return set_error("sgcn_relu_mask_grad: null pointer");
  if (numel % 4 != 0) return set_error("sgcn_relu_mask_grad: numel must be a multiple of 4");
  if (numel <= 0) return 0;
  relu_mask_grad_kernel<<<pw_grid(numel / 4, kPwThreads * 4), kPwThreads, 0, (cudaStream_t)stream>>>(
      (const float4*)g, (const float4*)y, (float4*)out, numel / 4);
  return check_launch("relu_mask_grad_kernel");
}
