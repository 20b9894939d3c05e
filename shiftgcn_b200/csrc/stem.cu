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
//        both backward kernels receive g and h through a two-stage cp.async.bulk ring (chunks of <= 8 groups)
#include "capi_internal.h"
#include "common.cuh"
#include "pointwise.h"

namespace sgcn {
namespace stem {

constexpr int D = 64;              // output channels of the first unit (model/shift_gcn.py:178)
constexpr int kGS = 16;            // groups staged per step

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

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// all cp.async of this thread, committed or not (cp.async.wait_group alone ignores the copies that were never committed)
__device__ __forceinline__ void cp_async_drain() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// request groups [g, g+n) of x into shared memory (coalesced 4-byte cp.async: x has no 16-byte alignment per group),
// n <= kGS; the caller waits with cp_async_drain() + __syncthreads() one chunk later
__device__ __forceinline__ void prefetch_x(const SgcnStem& p, float* sx, long long g, int n) {
  const int cnt = n * p.V * 3;
  const float* src = p.x + (size_t)g * p.V * 3;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) cp_async4(sx + i, src + i);
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
  __shared__ float sxb[2][kGS * 40 * 3];                          // x of this chunk / of the next one (in flight)
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
  prefetch_x(p, sxb[0], g0, min(kGS, ng));
  for (int gs = 0, ci = 0; gs < ng; gs += kGS, ++ci) {
    const int n = min(kGS, ng - gs);
    const float* sx = sxb[ci & 1];
    cp_async_drain();
    __syncthreads();                                               // this chunk's x has landed, the other buffer is free
    if (gs + kGS < ng) prefetch_x(p, sxb[(ci & 1) ^ 1], g0 + gs + kGS, min(kGS, ng - gs - kGS));
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
// ---- the two full-size streams of the backward (g and h, 2a) reach shared memory by bulk copies: a chunk of kgb row
// groups of either tensor is one contiguous span of kgb*V*64 floats, so one elected thread requests both spans with
// cp.async.bulk onto an mbarrier and the copy of chunk i+2 / i+1 runs under the arithmetic of chunk i (two stages).
// Before: every thread kept 8 of its own global loads in flight and waited for them 4 times per chunk behind two
// table loads (26 KB in flight per SM, ~7 serialised memory round trips per chunk of 8 groups: 585 us at NTU batch 64).
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// request groups [g, g+n) of p.g and p.h into one stage: g at stage[0 ..), h at stage[half ..)
__device__ __forceinline__ void issue_chunk(const SgcnStem& p, float* stage, int half, uint64_t* bar, long long g, int n) {
  const uint32_t bytes = (uint32_t)n * (uint32_t)p.V * D * 4u;
  mbar_expect_tx(bar, 2u * bytes);
  bulk_load(smem_u32(stage), p.g + (size_t)g * p.V * D, bytes, bar);
  bulk_load(smem_u32(stage + half), p.h + (size_t)g * p.V * D, bytes, bar);
}
constexpr int kRedPitch = 65, kRedFloats = 40 * 3 * kRedPitch + 8;   // mask-gradient staging (V <= 39), conflict-free pitch
constexpr int kSxFloats = 624;     // kgb * V * 3 <= 609 for every V <= 39 (host: chunk_groups)

// ONE pass over the block's groups: chunks of kgb groups arrive in shared memory, a thread visits its joints per chunk.
template <int kMaxJ>
__global__ void __launch_bounds__(832, 1) stem_bwd_stats_kernel(const SgcnStem p, int gper, int rev, int kgb) {
  extern __shared__ __align__(128) float dyn[];                    // 2 stages x { g [kgb][V][64], h [kgb][V][64] }
  __shared__ float sx[2][kSxFloats];
  __shared__ uint64_t full[2];
  const Lay l = layout(p.V);
  const int V = p.V, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cb = warp / l.nwj, jw = warp - cb * l.nwj, d = cb * 32 + lane;
  const long long g0 = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * gper;   // snake traversal
  const int ng = (int)((p.groups - g0) < gper ? (p.groups - g0) : gper);
  const int half = kgb * V * D, nchunks = (ng + kgb - 1) / kgb;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  float wc[3], wd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wc[c] = __ldg(p.W + c * D + d);
    wd[c] = __ldg(p.Wd + d * 3 + c);
  }
  const float b = p.bias ? __ldg(p.bias + d) : 0.f, bd = p.bd ? __ldg(p.bd + d) : 0.f;
  const float m2 = __ldg(p.mean2 + d), i2 = __ldg(p.invstd2 + d);
  float acc[2] = {0.f, 0.f};                                       // {sum gm, sum gm*rhat}
  int xo[kMaxJ][3];
  float mw[kMaxJ][3], k0[kMaxJ], k1[kMaxJ], s0[kMaxJ], s1[kMaxJ];
  bool own[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    const int vraw = jw + j * l.nwj;
    own[j] = j < l.jp && vraw < V;                                 // warp uniform
    const int v = own[j] ? vraw : 0;
    int u = v - d % V;
    if (u < 0) u += V;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xo[j][c] = ((u + c) % V) * 3 + c;
      mw[j][c] = __ldg(p.maskmul + u * 3 + c) * wc[c];
    }
    k0[j] = __ldg(p.mean1 + v * D + d);                            // mean / invstd of z
    k1[j] = __ldg(p.invstd1 + v * D + d);
    s0[j] = s1[j] = 0.f;
  }
  __syncthreads();                                                 // barriers initialised
  if (tid == 0)
    for (int s = 0; s < 2 && s < nchunks; ++s)
      issue_chunk(p, dyn + (size_t)s * 2 * half, half, &full[s], g0 + (long long)s * kgb, min(kgb, ng - s * kgb));
  for (int i = tid; i < min(kgb, ng) * V * 3; i += blockDim.x) sx[0][i] = __ldg(p.x + (size_t)g0 * V * 3 + i);
  __syncthreads();
  for (int ci = 0; ci < nchunks; ++ci) {
    const int s = ci & 1, gs = ci * kgb, n = min(kgb, ng - gs);
    const float* sG = dyn + (size_t)s * 2 * half;
    const float* sH = sG + half;
    const float* sxc = sx[s];
    if (ci + 1 < nchunks) {                                        // x of the next chunk: one value per thread, no register
      const int cntn = min(kgb, ng - gs - kgb) * V * 3;
      if (tid < cntn) cp_async4(&sx[s ^ 1][tid], p.x + (size_t)(g0 + gs + kgb) * V * 3 + tid);
    }
    mbar_wait(&full[s], (uint32_t)((ci >> 1) & 1));
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (own[j]) {
        const int v = jw + j * l.nwj;
        const float* xs = sxc + xo[j][0];
        const int o1 = xo[j][1] - xo[j][0], o2 = xo[j][2] - xo[j][0];
        const float* xp = sxc + v * 3;
        const int e0 = v * D + d;
#pragma unroll 4
        for (int q = 0; q < n; ++q) {
          const int o = q * V * 3, e = e0 + q * V * D;
          const float z = fmaf(xs[o], mw[j][0], fmaf(xs[o + o1], mw[j][1], fmaf(xs[o + o2], mw[j][2], b)));
          const float r = fmaf(xp[o], wd[0], fmaf(xp[o + 1], wd[1], fmaf(xp[o + 2], wd[2], bd)));
          const float gm = sH[e] > 0.f ? sG[e] : 0.f;
          s0[j] += gm;
          s1[j] = fmaf(gm, (z - k0[j]) * k1[j], s1[j]);
          acc[0] += gm;
          acc[1] = fmaf(gm, (r - m2) * i2, acc[1]);
        }
      }
    cp_async_drain();
    __syncthreads();                                               // stage s and sx[s] are free, sx[s^1] is written
    if (tid == 0 && ci + 2 < nchunks)
      issue_chunk(p, dyn + (size_t)s * 2 * half, half, &full[s], g0 + gs + 2 * kgb, min(kgb, ng - gs - 2 * kgb));
  }
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j)
    if (own[j]) {
      const size_t f = (size_t)(jw + j * l.nwj) * D + d;
      atomicAdd(p.vd_sums + 2 * f, (double)s0[j]);
      atomicAdd(p.vd_sums + 2 * f + 1, (double)s1[j]);
    }
  reduce_channels<2>(acc, p.r_sums, d, dyn, l.nwj);                // the stages are idle: every copy has been consumed
}

// ------------------------------------------------------------------------------------------------ backward, apply
// dz, dr from the finalised BatchNorm tables; dW, dWd, dbd, dMask accumulate in registers over the whole walk; dx in
// two phases per chunk of kgb groups (no shared-memory atomics, no per-element warp reductions):
//   phase A  thread = (joint v, channel d): reads g and h from the stage the bulk copies filled, recomputes z and r
//            from the staged x, keeps the parameter gradients in registers and leaves dz, dr [g][v][d] IN PLACE of g, h
//   phase B  warp = one (group g, joint w) pair per turn, lane = channels d = lane and lane + 32, three outputs c:
//            dx[g,w,c] = m[u,c] * sum_d dz[g,(u+d)%V,d] W[c,d]  +  sum_d dr[g,w,d] Wd[d,c],   u = (w-c) mod V
//            -- lane = channel makes the row gathers bank-conflict free; the sums are combined with shuffles
//            (the serial 64-channel loop per output it replaces spent ~20 instructions per channel on index wraps)
template <int kMaxJ>
__global__ void __launch_bounds__(832, 1) stem_bwd_apply_kernel(const SgcnStem p, int gper, int rev, int kgb) {
  extern __shared__ __align__(128) float dyn[];                    // 2 stages x { g -> dz [kgb][V][64], h -> dr [kgb][V][64] }
  __shared__ float sx[2][kSxFloats];
  __shared__ float smk[40 * 3];
  __shared__ float sW[3 * D], sWd[3 * D];                          // W[c][d] and Wd[d][c] transposed to [c][d]
  __shared__ uint64_t full[2];
  const Lay l = layout(p.V);
  const int V = p.V, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cb = warp / l.nwj, jw = warp - cb * l.nwj, d = cb * 32 + lane;
  const long long g0 = (long long)(rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * gper;   // snake traversal
  const int ng = (int)((p.groups - g0) < gper ? (p.groups - g0) : gper);
  const int half = kgb * V * D, nchunks = (ng + kgb - 1) / kgb;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  float wc[3], wd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wc[c] = __ldg(p.W + c * D + d);
    wd[c] = __ldg(p.Wd + d * 3 + c);
  }
  for (int i = tid; i < V * 3; i += blockDim.x) smk[i] = __ldg(p.maskmul + i);
  for (int i = tid; i < 3 * D; i += blockDim.x) {
    sW[i] = __ldg(p.W + i);
    sWd[i] = __ldg(p.Wd + (i % D) * 3 + i / D);
  }
  const float b = p.bias ? __ldg(p.bias + d) : 0.f, bd = p.bd ? __ldg(p.bd + d) : 0.f;
  const float a2 = __ldg(p.a2 + d), b2 = __ldg(p.b2 + d), c2 = __ldg(p.c2 + d);
  const int lmA = lane % V, lmB = (lane + 32) % V, nwarps = (int)(blockDim.x >> 5);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};         // dW[3], dWd[3], dbd, unused
  float dM[kMaxJ][3], kt[kMaxJ][3];                                // dM: sum x_c * dz per joint; kt: the per-(v,d) BatchNorm backward coefficients
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    dM[j][0] = dM[j][1] = dM[j][2] = 0.f;
    const int v = jw + j * l.nwj;
    const int f = (j < l.jp && v < V ? v : 0) * D + d;
    kt[j][0] = __ldg(p.al + f);
    kt[j][1] = __ldg(p.be + f);
    kt[j][2] = __ldg(p.ga + f);
  }
  __syncthreads();                                                 // barriers initialised, tables staged
  if (tid == 0)
    for (int s = 0; s < 2 && s < nchunks; ++s)
      issue_chunk(p, dyn + (size_t)s * 2 * half, half, &full[s], g0 + (long long)s * kgb, min(kgb, ng - s * kgb));
  for (int i = tid; i < min(kgb, ng) * V * 3; i += blockDim.x) sx[0][i] = __ldg(p.x + (size_t)g0 * V * 3 + i);
  __syncthreads();

  for (int ci = 0; ci < nchunks; ++ci) {
    const int s = ci & 1, gs = ci * kgb, n = min(kgb, ng - gs);
    float* sDz = dyn + (size_t)s * 2 * half;
    float* sDr = sDz + half;
    const float* sxc = sx[s];
    if (ci + 1 < nchunks) {                                        // x of the next chunk: one value per thread, no register
      const int cntn = min(kgb, ng - gs - kgb) * V * 3;
      if (tid < cntn) cp_async4(&sx[s ^ 1][tid], p.x + (size_t)(g0 + gs + kgb) * V * 3 + tid);
    }
    mbar_wait(&full[s], (uint32_t)((ci >> 1) & 1));
    // ---------------------------------------------------------------- phase A
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
      const int v = jw + j * l.nwj;
      if (j < l.jp && v < V) {                                     // warp uniform
        int u = v - d % V;
        if (u < 0) u += V;
        int xo[3];
        float mw[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int uc = u + c;
          if (uc >= V) uc -= V;
          xo[c] = uc * 3 + c;
          mw[c] = smk[u * 3 + c] * wc[c];
        }
        const float k0 = kt[j][0], k1 = kt[j][1], k2 = kt[j][2];
        const float* xp = sxc + v * 3;
        float* dzp = sDz + v * D + d;
        float* drp = sDr + v * D + d;
#pragma unroll 4
        for (int q = 0; q < n; ++q) {
          const int o = q * V * 3, e = q * V * D;
          const float gv = dzp[e], hv = drp[e];
          const float x0 = sxc[o + xo[0]], x1 = sxc[o + xo[1]], x2 = sxc[o + xo[2]];
          const float p0 = xp[o], p1 = xp[o + 1], p2 = xp[o + 2];
          const float z = fmaf(x0, mw[0], fmaf(x1, mw[1], fmaf(x2, mw[2], b)));
          const float r = fmaf(p0, wd[0], fmaf(p1, wd[1], fmaf(p2, wd[2], bd)));
          const float gm = hv > 0.f ? gv : 0.f;
          const float dz = fmaf(k0, gm, fmaf(k1, z, k2));
          const float dr = fmaf(a2, gm, fmaf(b2, r, c2));
          // Linear_weight / Feature_Mask gradients (autograd of :128-131) share t = sum x_c * dz: dM = t * W[c,d] and
          // dW[c,d] = sum over the joints of t * m[u,c] are formed once, after the walk
          dM[j][0] = fmaf(x0, dz, dM[j][0]);
          dM[j][1] = fmaf(x1, dz, dM[j][1]);
          dM[j][2] = fmaf(x2, dz, dM[j][2]);
          // down conv weight / bias gradients
          acc[3] = fmaf(p0, dr, acc[3]);
          acc[4] = fmaf(p1, dr, acc[4]);
          acc[5] = fmaf(p2, dr, acc[5]);
          acc[6] += dr;
          dzp[e] = dz;
          drp[e] = dr;
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase B
    {
      float wq[2][3], wdq[2][3];                                   // W[c][d], Wd[d][c] of this lane's two channels
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wq[hh][c] = sW[c * D + lane + 32 * hh];
          wdq[hh][c] = sWd[c * D + lane + 32 * hh];
        }
      float* dxo = p.dx + (size_t)(g0 + gs) * V * 3;
      for (int w = warp; w < V; w += nwarps) {                     // V <= warps at V = 25: one joint per warp, no division
        int offA[3], offB[3];
        float mkc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int u = w - c;
          while (u < 0) u += V;
          int ra = u + lmA, rb = u + lmB;                          // (u + d) mod V for d = lane, lane + 32
          if (ra >= V) ra -= V;
          if (rb >= V) rb -= V;
          offA[c] = ra * D;
          offB[c] = rb * D + 32;
          mkc[c] = smk[u * 3 + c];
        }
        const float* dzg = sDz + lane;
        const float* drw = sDr + w * D + lane;
        auto sums = [&](const float* zg, const float* rw, float (&sc)[3]) {
          const float r0 = rw[0], r1 = rw[32];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float zs = fmaf(zg[offA[c]], wq[0][c], zg[offB[c]] * wq[1][c]);
            sc[c] = fmaf(zs, mkc[c], fmaf(r0, wdq[0][c], r1 * wdq[1][c]));
          }
        };
        // two groups per turn: their six sums over the 32 lanes share one packed butterfly (8 shuffles instead of 30):
        // after the xor-16 step the low half-warp carries the even group and the high one the odd group, after xor-8 /
        // xor-4 bit 3 / bit 2 of the lane select the output channel -- the results end in lanes 0 / 8 / 4 (+16)
        const bool hi = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        const bool writer = (lane & 3) == 0 && (lane & 12) != 12;
        float* dxw = dxo + w * 3 + (hi ? V * 3 : 0) + (b2 ? 2 : (b3 ? 1 : 0));
        for (int g = 0; g < n; g += 2) {
          const bool two = g + 1 < n;                              // warp uniform
          float sa[3], sb[3] = {0.f, 0.f, 0.f};
          sums(dzg, drw, sa);
          if (two) sums(dzg + V * D, drw + V * D, sb);
          const float t0 = (hi ? sb[0] : sa[0]) + __shfl_xor_sync(0xffffffffu, hi ? sa[0] : sb[0], 16);
          const float t1 = (hi ? sb[1] : sa[1]) + __shfl_xor_sync(0xffffffffu, hi ? sa[1] : sb[1], 16);
          const float t2 = (hi ? sb[2] : sa[2]) + __shfl_xor_sync(0xffffffffu, hi ? sa[2] : sb[2], 16);
          const float s01 = (b3 ? t1 : t0) + __shfl_xor_sync(0xffffffffu, b3 ? t0 : t1, 8);
          const float s2 = t2 + __shfl_xor_sync(0xffffffffu, t2, 8);
          float r = (b2 ? s2 : s01) + __shfl_xor_sync(0xffffffffu, b2 ? s01 : s2, 4);
          r += __shfl_xor_sync(0xffffffffu, r, 2);
          r += __shfl_xor_sync(0xffffffffu, r, 1);
          if (writer && (!hi || two)) *dxw = r;
          dzg += 2 * V * D;
          drw += 2 * V * D;
          dxw += 2 * V * 3;
        }
      }
    }
    cp_async_drain();
    fence_proxy_async();                                           // this thread's dz / dr stores before the next bulk copy
    __syncthreads();                                               // phase B is done with stage s
    if (tid == 0 && ci + 2 < nchunks)
      issue_chunk(p, dyn + (size_t)s * 2 * half, half, &full[s], g0 + gs + 2 * kgb, min(kgb, ng - gs - 2 * kgb));
  }
  // Feature_Mask gradient: every (u, c) receives exactly one term per channel d (from the thread that owns joint
  // (u + d) mod V); they are combined in shared memory in a fixed order and leave as ONE fp64 atomic per (u, c) and
  // block (a per-thread atomic put 4992 of them per block on 75 addresses: a third of the kernel's time)
  __syncthreads();                                                 // the stages are idle: every copy has been consumed
  float* red = dyn;                                                // [V*3][kRedPitch]
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    const int v = jw + j * l.nwj;
    if (j < l.jp && v < V) {
      int u = v - d % V;
      if (u < 0) u += V;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] = fmaf(dM[j][c], smk[u * 3 + c], acc[c]);
        red[(u * 3 + c) * kRedPitch + d] = dM[j][c] * wc[c];
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < V * 3; i += blockDim.x) {
    double sum = 0.0;
    for (int k = 0; k < D; ++k) sum += (double)red[i * kRedPitch + k];
    atomicAdd(p.dmask_raw + i, sum);
  }
  reduce_channels<8>(acc, p.dw_raw, d, dyn + kRedFloats, l.nwj);   // [d][8]: dW[0..2][d], dWd[d][0..2], dbd[d], unused
}

// groups per bulk-copied chunk: two stages of { g, h } next to ~10 KB of static shared memory, one x value per thread
static int chunk_groups(int V, int threads) {
  int k = 104000 / (V * D * 8);
  k = k > 8 ? 8 : (k < 1 ? 1 : k);
  while (k > 1 && k * V * 3 > threads) --k;
  return k;
}
static size_t bwd_smem(int V, int kgb) {                          // the cross-warp reduction scratch aliases the stages
  const size_t stages = (size_t)2 * 2 * kgb * V * D * sizeof(float), scratch = (size_t)(kRedFloats + 16 * 8 * D) * sizeof(float);
  return stages > scratch ? stages : scratch;
}

static int check(const SgcnStem* p) {
  if (!p || !p->x || !p->maskmul || !p->W || !p->Wd) return set_error("sgcn_stem: null pointer");
  if (p->D != D) return set_error("sgcn_stem: the first unit has 64 output channels");
  if (p->V < 1 || p->V > 39) return set_error("sgcn_stem: num_point must be in [1, 39]");
  return 0;
}

static int per_block(long long groups) {
  const long long want = (long long)tile_ctas();                // one block (all 64 channels x all joints) per SM (or the test cap)
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
  if ((reinterpret_cast<uintptr_t>(p->g) | reinterpret_cast<uintptr_t>(p->h)) & 15)
    return set_error("sgcn_stem_bwd: g and h must be 16-byte aligned (bulk copies)");
  if (p->groups <= 0) return 0;
  const stem::Lay l = stem::layout(p->V);
  const int threads = 2 * l.nwj * 32, per = stem::per_block(p->groups);
  const unsigned grid = (unsigned)((p->groups + per - 1) / per);
  const int rev = next_direction();
  const int kgb = stem::chunk_groups(p->V, threads);
  const size_t smem = stem::bwd_smem(p->V, kgb);
  static std::atomic<unsigned long long> configured{0};           // one bit per device (the attribute is per device)
  if (needs_configure(configured)) {
    const int cap = 214 * 1024;                                   // >= bwd_smem of every V <= 39
    cudaError_t e = cudaFuncSetAttribute(stem::stem_bwd_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem::stem_bwd_apply_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem::stem_bwd_stats_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem::stem_bwd_stats_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e != cudaSuccess) return set_cuda_error("stem_bwd smem attribute", e);
    mark_configured(configured);
  }
  if (mode == 0) {
    if (!p->mean1 || !p->invstd1 || !p->mean2 || !p->invstd2 || !p->vd_sums || !p->r_sums)
      return set_error("sgcn_stem_bwd(stats): null pointer");
    if (l.jp <= 2) stem::stem_bwd_stats_kernel<2><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev, kgb);
    else stem::stem_bwd_stats_kernel<3><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev, kgb);
  } else {
    if (!p->al || !p->be || !p->ga || !p->a2 || !p->b2 || !p->c2 || !p->dw_raw || !p->dmask_raw || !p->dx)
      return set_error("sgcn_stem_bwd(apply): null pointer");
    if (l.jp <= 2) stem::stem_bwd_apply_kernel<2><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev, kgb);
    else stem::stem_bwd_apply_kernel<3><<<grid, threads, smem, (cudaStream_t)stream>>>(*p, per, rev, kgb);
  }
  return check_launch("stem_bwd_kernel");
}
