// stem.cu -- the 3-channel first spatial unit (l1.gcn1 = Shift_gcn(3, 64), model/shift_gcn.py:178, 121-142) as four
// SIMT kernels.  With three input channels the "contraction" is three FMAs per output, so nothing is worth storing:
// z (the pre-BatchNorm spatial output) and the 1x1-conv residual r are RECOMPUTED from the tiny input wherever they
// are needed, and the only full-size tensors that cross HBM are h (written once, forward) and g, h (read twice,
// backward) -- 1a + 4a instead of the ~30a of the generic path.
//
//   z[g,v,d]  = b[d]  + sum_c x[g,(u+c)%V,c] * m[u,c] * W[c,d],   u = (v-d) mod V        (:127-136)
//   r[g,v,d]  = bd[d] + sum_c x[g,v,c] * Wd[d,c]                                          (down conv, :82-86)
//   h         = relu(BN1d_{v,d}(z) + BN2d_d(r))                                           (:137-141)
//
// Thread = (joints v, v+nwj, ..; channel d); block = all 64 channels x all joints; the block walks a chunk of row
// groups (frames), staging 16 groups of x (16*V*3 floats) in shared memory at a time.
//   fwd  mode 0  batch statistics of z per (v,d) and of r per d            (no HBM traffic beyond x)
//        mode 1  h, and the per-channel statistics of h the temporal unit's first BatchNorm needs
//   bwd  mode 0  gm = g*[h>0]; BatchNorm backward sums per (v,d) and per d
//        mode 1  dz, dr from the finalised tables; dW, dWd, dMask accumulate in registers over the whole walk, dx
//                in a second phase per chunk from dz / dr left in shared memory (no atomics)
#include "capi_internal.h"
#include "common.cuh"
#include "pointwise.h"

namespace sgcn {
namespace stem {

constexpr int D = 64;              // output channels of the first unit (model/shift_gcn.py:178)
constexpr int kGS = 16;            // groups staged per step
constexpr int kUn = 4;             // groups in flight per thread (global loads of the backward)

struct Lay {
  int nwj, jp;                     // warps per channel block, joints per thread
};
__host__ __device__ inline Lay layout(int V) {
  Lay l;
  l.jp = (V + 12) / 13;
  l.nwj = (V + l.jp - 1) / l.jp;
  return l;
}

template <int kMaxJ>               // joints per thread: 2 (V <= 26) or 3 (V <= 39)
struct Thread {
  int lane, d, jw, nj;             // channel, first joint, joints owned
  int v[kMaxJ], u[kMaxJ];          // output joint, source row of the rotation u = (v-d) mod V
  int xo[kMaxJ][3];                // float offset of x[(u+c)%V, c] inside a group
  float mw[kMaxJ][3];              // m[u,c] * W[c,d]
  float wd[3];                     // Wd[d,c]
  float b, bd;
};

template <int kMaxJ>
__device__ __forceinline__ Thread<kMaxJ> setup(const SgcnStem& p) {
  Thread<kMaxJ> t;
  const Lay l = layout(p.V);
  const int warp = threadIdx.x >> 5;
  t.lane = threadIdx.x & 31;
  const int cb = warp / l.nwj;
  t.jw = warp - cb * l.nwj;
  t.d = cb * 32 + t.lane;
  t.nj = 0;
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    const int v = t.jw + j * l.nwj;
    const bool ok = j < l.jp && v < p.V;
    if (ok) t.nj = j + 1;
    t.v[j] = ok ? v : 0;
    int u = t.v[j] - t.d % p.V;
    if (u < 0) u += p.V;
    t.u[j] = u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      t.xo[j][c] = ((u + c) % p.V) * 3 + c;
      t.mw[j][c] = __ldg(p.maskmul + u * 3 + c) * __ldg(p.W + c * D + t.d);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) t.wd[c] = __ldg(p.Wd + t.d * 3 + c);
  t.b = p.bias ? __ldg(p.bias + t.d) : 0.f;
  t.bd = p.bd ? __ldg(p.bd + t.d) : 0.f;
  return t;
}

// stage groups [g, g+n) of x into shared memory (coalesced), n <= kGS
__device__ __forceinline__ void stage_x(const SgcnStem& p, float* sx, long long g, int n) {
  const int cnt = n * p.V * 3;
  const float* src = p.x + (size_t)g * p.V * 3;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sx[i] = __ldg(src + i);
}

template <int NV>
__device__ __forceinline__ void reduce_channels(const float (&v)[NV], double* __restrict__ dst, int d, float* scratch,
                                                int nwj) {
  // scratch [nwj][NV][64]: combine the joints (warps) of each channel, one fp64 atomic per value
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = warp / nwj, jw = warp - cb * nwj;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) scratch[(jw * NV + k) * D + cb * 32 + lane] = v[k];
  __syncthreads();
  for (int i = threadIdx.x; i < NV * D; i += blockDim.x) {
    const int k = i / D, dd = i - k * D;
    double s = 0.0;
    for (int w = 0; w < nwj; ++w) s += (double)scratch[(w * NV + k) * D + dd];
    atomicAdd(dst + (size_t)dd * NV + k, s);
  }
  (void)d;
}

// ------------------------------------------------------------------------------------------------ forward
template <int MODE, int kMaxJ>
__global__ void __launch_bounds__(832, 1) stem_fwd_kernel(const SgcnStem p, int gper, int rev) {
  __shared__ float sx[kGS * 40 * 3];
  __shared__ float scratch[16 * 2 * D];
  const Thread<kMaxJ> t = setup<kMaxJ>(p);
  const Lay l = layout(p.V);
  const int V = p.V;
  const long long g0 = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * gper;   // snake traversal
  const int ng = (int)((p.groups - g0) < gper ? (p.groups - g0) : gper);
  float sc1[kMaxJ], sh1[kMaxJ], sc2 = 0.f, sh2 = 0.f;
  float sz[kMaxJ], szz[kMaxJ], acc[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    sz[j] = szz[j] = 0.f;
    sc1[j] = sh1[j] = 0.f;
    if (MODE == 1 && j < t.nj) {
      sc1[j] = __ldg(p.sc1 + t.v[j] * D + t.d);
      sh1[j] = __ldg(p.sh1 + t.v[j] * D + t.d);
    }
  }
  if (MODE == 1) {
    sc2 = __ldg(p.sc2 + t.d);
    sh2 = __ldg(p.sh2 + t.d) + t.bd * sc2;                       // conv bias folded into the BN shift
  }
  for (int gs = 0; gs < ng; gs += kGS) {
    const int n = min(kGS, ng - gs);
    __syncthreads();
    stage_x(p, sx, g0 + gs, n);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (j < t.nj) {
        const float* xs = sx + t.xo[j][0];
        const int o1 = t.xo[j][1] - t.xo[j][0], o2 = t.xo[j][2] - t.xo[j][0];
        const float* xp = sx + t.v[j] * 3;
        float* hp = MODE == 1 ? p.h + ((size_t)(g0 + gs) * V + t.v[j]) * D + t.d : nullptr;
#pragma unroll 4
        for (int gi = 0; gi < n; ++gi) {
          const int o = gi * V * 3;
          const float z = fmaf(xs[o], t.mw[j][0], fmaf(xs[o + o1], t.mw[j][1], fmaf(xs[o + o2], t.mw[j][2], t.b)));
          const float r = fmaf(xp[o], t.wd[0], fmaf(xp[o + 1], t.wd[1], xp[o + 2] * t.wd[2]));   // without the conv bias
          if (MODE == 0) {
            sz[j] += z;
            szz[j] = fmaf(z, z, szz[j]);
            const float rb = r + t.bd;
            acc[0] += rb;
            acc[1] = fmaf(rb, rb, acc[1]);
          } else {
            const float hv = fmaxf(fmaf(z, sc1[j], sh1[j]) + fmaf(r, sc2, sh2), 0.f);
            hp[(size_t)gi * V * D] = hv;
            acc[0] += hv;
            acc[1] = fmaf(hv, hv, acc[1]);
          }
        }
      }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (j < t.nj) {
        const size_t f = (size_t)t.v[j] * D + t.d;
        atomicAdd(p.stats_vd + 2 * f, (double)sz[j]);
        atomicAdd(p.stats_vd + 2 * f + 1, (double)szz[j]);
      }
    reduce_channels<2>(acc, p.stats_r, t.d, scratch, l.nwj);
  } else if (p.stats_h) {
    reduce_channels<2>(acc, p.stats_h, t.d, scratch, l.nwj);
  }
}

// ------------------------------------------------------------------------------------------------ backward, statistics
// gm = g*[h>0]; BatchNorm backward sums per (v,d) (of z) and per d (of the conv output r), both recomputed from x.
// One pass over the block's groups PER JOINT the thread owns (x is tiny and is simply staged again).
template <int kMaxJ>
__global__ void __launch_bounds__(832, 1) stem_bwd_stats_kernel(const SgcnStem p, int gper, int rev) {
  __shared__ float sx[kGS * 40 * 3];
  __shared__ float scratch[16 * 2 * D];
  const Lay l = layout(p.V);
  const int V = p.V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = warp / l.nwj, jw = warp - cb * l.nwj, d = cb * 32 + lane;
  const long long g0 = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * gper;   // snake traversal
  const int ng = (int)((p.groups - g0) < gper ? (p.groups - g0) : gper);
  float wc[3], wd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wc[c] = __ldg(p.W + c * D + d);
    wd[c] = __ldg(p.Wd + d * 3 + c);
  }
  const float b = p.bias ? __ldg(p.bias + d) : 0.f, bd = p.bd ? __ldg(p.bd + d) : 0.f;
  const float m2 = __ldg(p.mean2 + d), i2 = __ldg(p.invstd2 + d);
  float acc[2] = {0.f, 0.f};                                       // {sum gm, sum gm*rhat}

#pragma unroll 1
  for (int j = 0; j < l.jp; ++j) {
    const int vraw = jw + j * l.nwj;
    const bool own = vraw < V;                                     // (warp uniform) the last pass may have no joint
    const int v = own ? vraw : 0;
    int u = v - d % V;
    if (u < 0) u += V;
    int xo[3];
    float mw[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xo[c] = ((u + c) % V) * 3 + c;
      mw[c] = __ldg(p.maskmul + u * 3 + c) * wc[c];
    }
    const int f = v * D + d;
    const float k0 = __ldg(p.mean1 + f), k1 = __ldg(p.invstd1 + f);   // mean / invstd of z
    float s0 = 0.f, s1 = 0.f;
    for (int gs = 0; gs < ng; gs += kGS) {
      const int n = min(kGS, ng - gs);
      __syncthreads();
      stage_x(p, sx, g0 + gs, n);
      __syncthreads();
      if (own) {
        const float* xs = sx + xo[0];
        const int o1 = xo[1] - xo[0], o2 = xo[2] - xo[0];
        const float* xp = sx + v * 3;
        const size_t ob = ((size_t)(g0 + gs) * V + v) * D + d;
        for (int gb = 0; gb < n; gb += kUn) {
          float gv[kUn], hv[kUn];
#pragma unroll
          for (int q = 0; q < kUn; ++q) {
            const size_t o = ob + (size_t)min(gb + q, n - 1) * V * D;
            gv[q] = __ldg(p.g + o);
            hv[q] = __ldg(p.h + o);
          }
#pragma unroll
          for (int q = 0; q < kUn; ++q)
            if (gb + q < n) {
              const int o = (gb + q) * V * 3;
              const float z = fmaf(xs[o], mw[0], fmaf(xs[o + o1], mw[1], fmaf(xs[o + o2], mw[2], b)));
              const float r = fmaf(xp[o], wd[0], fmaf(xp[o + 1], wd[1], fmaf(xp[o + 2], wd[2], bd)));
              const float gm = hv[q] > 0.f ? gv[q] : 0.f;
              s0 += gm;
              s1 = fmaf(gm, (z - k0) * k1, s1);
              acc[0] += gm;
              acc[1] = fmaf(gm, (r - m2) * i2, acc[1]);
            }
        }
      }
    }
    if (own) {
      atomicAdd(p.vd_sums + 2 * (size_t)f, (double)s0);
      atomicAdd(p.vd_sums + 2 * (size_t)f + 1, (double)s1);
    }
  }
  reduce_channels<2>(acc, p.r_sums, d, scratch, l.nwj);
}

// ------------------------------------------------------------------------------------------------ backward, apply
// dz, dr from the finalised BatchNorm tables; dW, dWd, dbd, dMask accumulate in registers over the whole walk; dx in
// two phases per chunk of kGB groups (no shared-memory atomics, no per-element warp reductions):
//   phase A  thread = (joint v, channel d): loads g and h (coalesced), recomputes z and r from the staged x, keeps the
//            parameter gradients in registers and leaves dz, dr [g][v][d] in shared memory
//   phase B  thread = one output value (group g, joint w, channel c), serial over the 64 channels d:
//            dx[g,w,c] = m[u,c] * sum_d dz[g,(u+d)%V,d] W[c,d]  +  sum_d dr[g,w,d] Wd[d,c],   u = (w-c) mod V
//            -- every lane starts its d loop at d = lane, so the row gathers are bank-conflict free (bank = d mod 32)
constexpr int kGB = 8;

template <int kMaxJ>
__global__ void __launch_bounds__(832, 1) stem_bwd_apply_kernel(const SgcnStem p, int gper, int rev) {
  extern __shared__ __align__(16) float dyn[];                     // dz [kGB][V][64], dr [kGB][V][64]
  __shared__ float sx[kGB * 40 * 3];
  __shared__ float smk[40 * 3];
  __shared__ float sW[3 * D], sWd[3 * D];                          // W[c][d] and Wd[d][c] transposed to [c][d]
  __shared__ float scratch[16 * 8 * D];
  const Lay l = layout(p.V);
  const int V = p.V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = warp / l.nwj, jw = warp - cb * l.nwj, d = cb * 32 + lane;
  float* sDz = dyn;
  float* sDr = dyn + (size_t)kGB * V * D;
  const long long g0 = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * gper;   // snake traversal
  const int ng = (int)((p.groups - g0) < gper ? (p.groups - g0) : gper);
  float wc[3], wd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wc[c] = __ldg(p.W + c * D + d);
    wd[c] = __ldg(p.Wd + d * 3 + c);
  }
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) smk[i] = __ldg(p.maskmul + i);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    sW[i] = __ldg(p.W + i);
    sWd[i] = __ldg(p.Wd + (i % D) * 3 + i / D);
  }
  const float b = p.bias ? __ldg(p.bias + d) : 0.f, bd = p.bd ? __ldg(p.bd + d) : 0.f;
  const float a2 = __ldg(p.a2 + d), b2 = __ldg(p.b2 + d), c2 = __ldg(p.c2 + d);
  const int lm0 = lane % V;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};         // dW[3], dWd[3], dbd, unused
  float dM[kMaxJ][3];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) dM[j][0] = dM[j][1] = dM[j][2] = 0.f;

  for (int gs = 0; gs < ng; gs += kGB) {
    const int n = min(kGB, ng - gs);
    __syncthreads();                                               // phase B of the previous chunk is done
    stage_x(p, sx, g0 + gs, n);
    __syncthreads();
    // ---------------------------------------------------------------- phase A
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
      const int v = jw + j * l.nwj;
      if (j < l.jp && v < V) {                                     // warp uniform
        int u = v - d % V;
        if (u < 0) u += V;
        int xo[3];
        float mk[3], mw[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int uc = u + c;
          if (uc >= V) uc -= V;
          xo[c] = uc * 3 + c;
          mk[c] = smk[u * 3 + c];
          mw[c] = mk[c] * wc[c];
        }
        const int f = v * D + d;
        const float k0 = __ldg(p.al + f), k1 = __ldg(p.be + f), k2 = __ldg(p.ga + f);
        const float* xp = sx + v * 3;
        const size_t ob = ((size_t)(g0 + gs) * V + v) * D + d;
        float* dzp = sDz + (size_t)v * D + d;
        float* drp = sDr + (size_t)v * D + d;
        for (int gb = 0; gb < n; gb += kUn) {
          float gv[kUn], hv[kUn];
#pragma unroll
          for (int q = 0; q < kUn; ++q) {
            const size_t o = ob + (size_t)min(gb + q, n - 1) * V * D;
            gv[q] = __ldg(p.g + o);
            hv[q] = __ldg(p.h + o);
          }
#pragma unroll
          for (int q = 0; q < kUn; ++q)
            if (gb + q < n) {
              const int o = (gb + q) * V * 3;
              const float x0 = sx[o + xo[0]], x1 = sx[o + xo[1]], x2 = sx[o + xo[2]];
              const float p0 = xp[o], p1 = xp[o + 1], p2 = xp[o + 2];
              const float z = fmaf(x0, mw[0], fmaf(x1, mw[1], fmaf(x2, mw[2], b)));
              const float r = fmaf(p0, wd[0], fmaf(p1, wd[1], fmaf(p2, wd[2], bd)));
              const float gm = hv[q] > 0.f ? gv[q] : 0.f;
              const float dz = fmaf(k0, gm, fmaf(k1, z, k2));
              const float dr = fmaf(a2, gm, fmaf(b2, r, c2));
              // Linear_weight / Feature_Mask gradients (autograd of :128-131)
              acc[0] = fmaf(x0 * mk[0], dz, acc[0]);
              acc[1] = fmaf(x1 * mk[1], dz, acc[1]);
              acc[2] = fmaf(x2 * mk[2], dz, acc[2]);
              dM[j][0] = fmaf(dz * wc[0], x0, dM[j][0]);
              dM[j][1] = fmaf(dz * wc[1], x1, dM[j][1]);
              dM[j][2] = fmaf(dz * wc[2], x2, dM[j][2]);
              // down conv weight / bias gradients
              acc[3] = fmaf(p0, dr, acc[3]);
              acc[4] = fmaf(p1, dr, acc[4]);
              acc[5] = fmaf(p2, dr, acc[5]);
              acc[6] += dr;
              dzp[(size_t)(gb + q) * V * D] = dz;
              drp[(size_t)(gb + q) * V * D] = dr;
            }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase B
    float* dxo = p.dx + (size_t)(g0 + gs) * V * 3;
    for (int o = threadIdx.x; o < n * V * 3; o += blockDim.x) {
      const int pair = o / 3, c = o - pair * 3;
      const int g = pair / V, w = pair - g * V;
      int u = w - c;
      if (u < 0) u += V;
      const float* dzg = sDz + (size_t)g * V * D;
      const float* drw = sDr + (size_t)pair * D;
      const float* wrow = sW + c * D;
      const float* wdrow = sWd + c * D;
      float s_shift = 0.f, s_conv = 0.f;
      int dd = lane, r = u + lm0;                                  // r = (u + dd) mod V
      if (r >= V) r -= V;
#pragma unroll 4
      for (int i = 0; i < D; ++i) {
        s_shift = fmaf(dzg[r * D + dd], wrow[dd], s_shift);
        s_conv = fmaf(drw[dd], wdrow[dd], s_conv);
        ++dd;
        ++r;
        if (r == V) r = 0;
        if (dd == D) {
          dd = 0;
          r = u;
        }
      }
      dxo[o] = fmaf(s_shift, smk[u * 3 + c], s_conv);
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    const int v = jw + j * l.nwj;
    if (j < l.jp && v < V) {
      int u = v - d % V;
      if (u < 0) u += V;
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(p.dmask_raw + u * 3 + c, (double)dM[j][c]);
    }
  }
  reduce_channels<8>(acc, p.dw_raw, d, scratch, l.nwj);            // [d][8]: dW[0..2][d], dWd[d][0..2], dbd[d], unused
}

static int check(const SgcnStem* p) {
  if (!p || !p->x || !p->maskmul || !p->W || !p->Wd) return set_error("sgcn_stem: null pointer");
  if (p->D != D) return set_error("sgcn_stem: the first unit has 64 output channels");
  if (p->V < 1 || p->V > 39) return set_error("sgcn_stem: num_point must be in [1, 39]");
  return 0;
}

static int per_block(long long groups) {
  const long long want = (long long)num_sms();                  // one block (all 64 channels x all joints) per SM
  long long per = (groups + want - 1) / want;
  per = (per + kGS - 1) / kGS * kGS;
  return (int)(per < kGS ? kGS : per);
}

}  // namespace stem
}  // namespace sgcn

using namespace sgcn;

extern "C" int sgcn_stem_fwd(const SgcnStem* p, int mode, void* stream) {
  if (int rc = stem::check(p)) return rc;
  if (p->groups <= 0) return 0;
  const stem::Lay l = stem::layout(p->V);
  const int threads = 2 * l.nwj * 32, per = stem::per_block(p->groups);
  const unsigned grid = (unsigned)((p->groups + per - 1) / per);
  const int rev = next_direction();
  if (mode == 0) {
    if (!p->stats_vd || !p->stats_r) return set_error("sgcn_stem_fwd(stats): null statistics buffer");
    if (l.jp <= 2) stem::stem_fwd_kernel<0, 2><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
    else stem::stem_fwd_kernel<0, 3><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
  } else {
    if (!p->h || !p->sc1 || !p->sh1 || !p->sc2 || !p->sh2) return set_error("sgcn_stem_fwd(apply): null pointer");
    if (l.jp <= 2) stem::stem_fwd_kernel<1, 2><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
    else stem::stem_fwd_kernel<1, 3><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
  }
  return check_launch("stem_fwd_kernel");
}

extern "C" int sgcn_stem_bwd(const SgcnStem* p, int mode, void* stream) {
  if (int rc = stem::check(p)) return rc;
  if (!p->g || !p->h) return set_error("sgcn_stem_bwd: null gradient / activation");
  if (p->groups <= 0) return 0;
  const stem::Lay l = stem::layout(p->V);
  const int threads = 2 * l.nwj * 32, per = stem::per_block(p->groups);
  const unsigned grid = (unsigned)((p->groups + per - 1) / per);
  const int rev = next_direction();
  if (mode == 0) {
    if (!p->mean1 || !p->invstd1 || !p->mean2 || !p->invstd2 || !p->vd_sums || !p->r_sums)
      return set_error("sgcn_stem_bwd(stats): null pointer");
    if (l.jp <= 2) stem::stem_bwd_stats_kernel<2><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
    else stem::stem_bwd_stats_kernel<3><<<grid, threads, 0, (cudaStream_t)stream>>>(*p, per, rev);
  } else {
    if (!p->al || !p->be || !p->ga || !p->a2 || !p->b2 || !p->c2 || !p->dw_raw || !p->dmask_raw || !p->dx)
      return set_error("sgcn_stem_bwd(apply): null pointer");
    const size_t smem = (size_t)2 * stem::kGB * p->V * stem::D * sizeof(float);
    static std::atomic<unsigned long long> configured{0};         // one bit per device (the attribute is per device)
    if (needs_configure(configured)) {
      const int cap = 2 * stem::kGB * 39 * stem::D * (int)sizeof(float);
      cudaError_t e = cudaFuncSetAttribute(stem::stem_bwd_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(stem::stem_bwd_apply_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
      if (e != cudaSuccess) return set_cuda_error("stem_bwd_apply smem attribute", e);
      mark_configured(configured);
    }
    if (l.jp <= 2) stem::stem_bwd_apply_kernel<2><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev);
    else stem::stem_bwd_apply_kernel<3><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev);
  }
  return check_launch("stem_bwd_kernel");
}
