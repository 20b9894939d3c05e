// pointwise.cu -- bandwidth-bound SIMT kernels of the Shift-GCN hot path (channels-last rows [(n,t,v), C]).
//
// "Walker" kernels: a thread owns ONE column (joint v, channel c) of one sample and walks along the frame
// axis t over a chunk of frames.  Consequences:
//   * lane <-> channel: every warp access is one coalesced 128-byte line, no index arithmetic per element
//     (the row pointer advances by the frame pitch V*C), no integer division anywhere in the loops;
//   * the fractional temporal shift is a two-tap filter along the walk direction, so the second tap of
//     frame t is the first tap of frame t+1 and stays in a register: every tensor crosses L1 exactly once;
//   * per-channel parameters (shift position, BatchNorm constants) live in registers for the whole walk;
//   * cross-row reductions (BatchNorm statistics, position / bias gradients) accumulate in registers, are
//     combined over the block's warps (= joints) in shared memory and published with one fp64 atomic per
//     value and block.
// Block = 32 channels x ceil(V/2) warps (each warp walks one or two joints); grid = channel blocks x samples x
// frame chunks.
//
//   bn_res_relu_fwd     h = relu(BN1d(z) + res)  (+ per-channel stats of h)           model/shift_gcn.py:137-141, :66
//   tshift_fwd          s = Shift_stride(q); stats of s  |  y = [relu](BN(s) + res)    :72-73, :161-162, K1
//   tshift_bwd_stats    BN backward sums + position-gradient sums of shift_out          autograd of :72-73, K4
//   tshift_bwd_apply    dpre = [q>0] * Shift^T(BN-bwd(g))  (+ conv-bias gradient)       K2/K3, autograd of :70-73
//   tshift_in_bwd_stats BN backward sums + position-gradient sums of shift_in           K2, K4, autograd of :66-68
//   tshift_in_bwd_apply gh = [h>0] * BN-bwd(Shift^T dp)  (+ per-(v,d) BN1d backward sums) autograd of :66, :137-141
// (K1..K5 = model/Temporal_shift/cuda/shift_cuda_kernel.cu:12-76, 79-152, 156-256, 278-363, 371-395.)
#include "capi_internal.h"
#include "common.cuh"
#include "pointwise.h"

namespace sgcn {

constexpr int kMaxWarps = 20;     // ceil(V/2) warps, V <= 40
#ifndef SGCN_WALKER_MINBLOCKS
#define SGCN_WALKER_MINBLOCKS 2   // register cap of the walkers: 65536 / (640 * MINBLOCKS)
#endif
#ifndef SGCN_UNROLL
#define SGCN_UNROLL 4
#endif
#ifndef SGCN_UNROLL_WIDE
#define SGCN_UNROLL_WIDE 8
#endif
#ifndef SGCN_UNROLL_MAX
#define SGCN_UNROLL_MAX 16
#endif
#ifndef SGCN_GEOM_MULT
#define SGCN_GEOM_MULT 8          // grid ~ this many blocks per SM (about two waves of the resident blocks)
#endif
constexpr int kUnroll = SGCN_UNROLL;           // frames in flight per thread (kernels with >= 3 loads per frame)
constexpr int kUnrollWide = SGCN_UNROLL_WIDE;  // ... with 2 loads per frame
constexpr int kUnrollMax = SGCN_UNROLL_MAX;    // ... with a single load per frame (read-only statistics passes are latency bound)

// ------------------------------------------------------------------------------------------------ helpers
struct Col {            // the walker's identity
  int lane, warp, nw;   // lane = channel inside the 32-channel block
  int c;                // channel
  long long n;          // sample (or group-chunk index)
  int chunk;            // frame chunk
};

// rev: walk the grid from its last block to its first ("snake" traversal, capi_internal.h:next_direction) so that
// the kernel starts on the part of its inputs that the previous kernel touched last and that is still in L2
__device__ __forceinline__ Col col_of(int C, int nchunks, int rev) {
  Col k;
  k.lane = threadIdx.x & 31;
  k.warp = threadIdx.x >> 5;
  k.nw = blockDim.x >> 5;
  const int cblocks = C >> 5;
  unsigned b = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned cb = b % cblocks;
  b /= cblocks;
  k.chunk = (int)(b % (unsigned)nchunks);
  k.n = b / (unsigned)nchunks;
  k.c = (int)cb * 32 + k.lane;
  return k;
}

// combine per-thread partials over the block's warps, then one double atomic per (channel, value)
template <int NV>
__device__ __forceinline__ void block_reduce_channels(const float (&v)[NV], double* __restrict__ dst, int c,
                                                      float* scratch /* [kMaxWarps * 32 * NV] */, int stride = NV) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) scratch[(warp * NV + k) * 32 + lane] = v[k];
  __syncthreads();
  for (int k = warp; k < NV; k += nw) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += (double)scratch[(w * NV + k) * 32 + lane];
    atomicAdd(dst + (size_t)c * stride + k, s);
  }
}

// zero-padded frame tap: value of frame tt of the column whose frame 0 is at `base`.
// The load itself is UNCONDITIONAL (row index clamped into [0, T)) and the padding is applied with a select:
// a predicated / branched load makes ptxas consume each result right behind its LDG, which serialises the
// loads of an unrolled batch (measured: 1 load in flight per warp instead of kUnroll).
__device__ __forceinline__ float ldrow(const float* __restrict__ base, int tt, int T, int pitch) {
  const int c = min(max(tt, 0), T - 1);
  return __ldg(base + (size_t)c * (size_t)pitch);
}
__device__ __forceinline__ bool inside(int tt, int T) { return (unsigned)tt < (unsigned)T; }
__device__ __forceinline__ float tap(const float* __restrict__ base, int tt, int T, int pitch) {
  const float v = ldrow(base, tt, T, pitch);
  return inside(tt, T) ? v : 0.f;
}

// U consecutive frames [t, t+U) of a column, zero padded outside [0, T).  Fast path (all U frames inside): one base
// address and U loads at compile-time offsets u*PITCH (PITCH = V*C floats, 0 = run-time pitch) -- no clamping, no
// per-load address arithmetic; this took the walkers from ~21 to ~8 issued instructions per 128-byte line.  The branch
// is per thread (the shift differs per channel) but both sides fill the same registers with independent loads.
template <int U, int PITCH>
__device__ __forceinline__ void load_frames(float (&d)[U], const float* __restrict__ base, int t, int T, int pitch) {
  if (t >= 0 && t + U <= T) {
    const float* __restrict__ q = base + (size_t)t * (size_t)(PITCH ? PITCH : pitch);
#pragma unroll
    for (int u = 0; u < U; ++u) d[u] = __ldg(q + (size_t)u * (size_t)(PITCH ? PITCH : pitch));
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) d[u] = tap(base, t + u, T, pitch);
  }
}

struct LerpCh {
  int y1;
  float f, g;    // weights of tap y1+1 and of tap y1  (g = 1 - f)
};
__device__ __forceinline__ LerpCh lerp_of(float ypos_eff) {
  const float fl = floorf(ypos_eff);
  const float f = ypos_eff - fl;
  return {(int)fl, f, 1.f - f};
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// h = relu(z * sc[v,d] + sh[v,d] + res)          stats_out[d] += {sum h, sum h^2}
// walks row GROUPS (frames of any sample): groups [g0, g1) of the chunk
template <int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) bn_res_relu_fwd_kernel(const float* __restrict__ z,
                                                                         const float* __restrict__ res,
                                                                         float* __restrict__ h,
                                                                         const float* __restrict__ sc,
                                                                         const float* __restrict__ sh,
                                                                         double* __restrict__ stats_out,
                                                                         long long groups, int gper, int nchunks, int V,
                                                                         int D, int relu, int rev) {
  __shared__ float scratch[kMaxWarps * 32 * 2];
  const Col k = col_of(D, nchunks, rev);
  const long long g0 = (long long)k.chunk * gper;
  const int ng = (int)((groups - g0) < gper ? (groups - g0) : gper);
  const int pitch = V * D;
  float acc[2] = {0.f, 0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const float a = __ldg(sc + v * D + k.c), b = __ldg(sh + v * D + k.c);
    const size_t o0 = ((size_t)g0 * V + v) * D + k.c;
    const float* zp = z + o0;
    const float* rp = res ? res + o0 : zp;                 // no residual: read z twice (L1 hit) and ignore it
    const float rsel = res ? 1.f : 0.f;
    float* hp = h + o0;
    for (int g = 0; g < ng; g += kUnrollWide) {
      float zv[kUnrollWide], rv[kUnrollWide];
      load_frames<kUnrollWide, PITCH>(zv, zp, g, ng, pitch);
      load_frames<kUnrollWide, PITCH>(rv, rp, g, ng, pitch);
#pragma unroll
      for (int u = 0; u < kUnrollWide; ++u)
        if (g + u < ng) {
          float y = fmaf(rv[u], rsel, fmaf(zv[u], a, b));
          if (relu) y = fmaxf(y, 0.f);
          hp[(size_t)(g + u) * pitch] = y;
          acc[0] += y;
          acc[1] = fmaf(y, y, acc[1]);
        }
    }
  }
  if (stats_out) block_reduce_channels<2>(acc, stats_out, k.c, scratch);
}

// The same op with 128-bit accesses: no shift is involved, so a thread can own FOUR consecutive channels of one joint
// (float4 column j of the V*D floats of a row group) and stream the groups of its chunk: whole 16-byte requests, a
// quarter of the load / store instructions.  Block = a divisor of the column count (<= 256); the per-channel statistics of the block are combined
// in shared memory (fp32 atomics on D*2 words) and published with one fp64 atomic per value and block.
#ifndef SGCN_BRR_U
#define SGCN_BRR_U 4
#endif
__global__ void __launch_bounds__(256, 4) bn_res_relu_fwd4_kernel(const float4* __restrict__ z, const float4* __restrict__ res,
                                                                  float4* __restrict__ h, const float4* __restrict__ sc,
                                                                  const float4* __restrict__ sh, double* __restrict__ stats_out,
                                                                  long long groups, int gper, int cols, int D, int relu,
                                                                  int rev) {
  extern __shared__ float sstat[];                                 // [D][2]
  const unsigned cblocks = cols / blockDim.x;               // blockDim.x divides cols
  unsigned b = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int j = (int)((b % cblocks) * blockDim.x + threadIdx.x);            // float4 column inside a row group
  const long long g0 = (long long)(b / cblocks) * gper;
  const int ng = (int)((groups - g0) < gper ? (groups - g0) : gper);
  if (stats_out) {
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sstat[i] = 0.f;
    __syncthreads();
  }
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (j < cols) {
    const float4 a = __ldg(sc + j), bb = __ldg(sh + j);
    const size_t o0 = (size_t)g0 * cols + j;
    const float4* zp = z + o0;
    const float4* rp = res ? res + o0 : zp;
    const float rsel = res ? 1.f : 0.f;
    float4* hp = h + o0;
    for (int g = 0; g < ng; g += SGCN_BRR_U) {
      float4 zv[SGCN_BRR_U], rv[SGCN_BRR_U];
#pragma unroll
      for (int u = 0; u < SGCN_BRR_U; ++u) {
        const size_t o = (size_t)min(g + u, ng - 1) * cols;
        zv[u] = __ldg(zp + o);
        rv[u] = __ldg(rp + o);
      }
#pragma unroll
      for (int u = 0; u < SGCN_BRR_U; ++u)
        if (g + u < ng) {
          float4 y;
          y.x = fmaf(rv[u].x, rsel, fmaf(zv[u].x, a.x, bb.x));
          y.y = fmaf(rv[u].y, rsel, fmaf(zv[u].y, a.y, bb.y));
          y.z = fmaf(rv[u].z, rsel, fmaf(zv[u].z, a.z, bb.z));
          y.w = fmaf(rv[u].w, rsel, fmaf(zv[u].w, a.w, bb.w));
          if (relu) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
          hp[(size_t)(g + u) * cols] = y;
          s1.x += y.x, s1.y += y.y, s1.z += y.z, s1.w += y.w;
          s2.x = fmaf(y.x, y.x, s2.x), s2.y = fmaf(y.y, y.y, s2.y), s2.z = fmaf(y.z, y.z, s2.z), s2.w = fmaf(y.w, y.w, s2.w);
        }
    }
  }
  if (stats_out) {
    if (j < cols) {
      const int d = (j * 4) % D;                                   // first of this thread's four channels
      atomicAdd(&sstat[2 * d + 0], s1.x), atomicAdd(&sstat[2 * d + 1], s2.x);
      atomicAdd(&sstat[2 * d + 2], s1.y), atomicAdd(&sstat[2 * d + 3], s2.y);
      atomicAdd(&sstat[2 * d + 4], s1.z), atomicAdd(&sstat[2 * d + 5], s2.z);
      atomicAdd(&sstat[2 * d + 6], s1.w), atomicAdd(&sstat[2 * d + 7], s2.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) atomicAdd(stats_out + i, (double)sstat[i]);
  }
}

// ------------------------------------------------------------------------------------------------ temporal shift, forward
// s(to) = g * Q(to*stride + y1) + f * Q(to*stride + y1 + 1), Q zero padded           (K1 with xpos = 0)
// MODE 0: stats[c] += {sum s, sum s^2};  MODE 1: out = [relu](s*sc + sh + res), and with a stats pointer the global
// pooling of the model head rides along: stats[n*C + c] += sum over (t, v) of out (model/shift_gcn.py:212-214); out may
// then be NULL (inference: the last unit's output is only ever pooled)
template <int MODE, bool S1, int PITCH, bool POOL = false>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_fwd_kernel(const SgcnTShift p, int tper, int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32 * 2];
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, Ti = p.T_in, To = p.T_out, st = S1 ? 1 : p.stride;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const float sc = MODE == 1 ? __ldg(p.scale + k.c) : 0.f, sh = MODE == 1 ? __ldg(p.shift + k.c) : 0.f;
  const int to0 = k.chunk * tper, to1 = min(To, to0 + tper);
  // loads per frame: statistics pass 1 (16 frames in flight), stride-1 apply 2 (8), strided apply 3 (4)
  constexpr int U = MODE == 0 ? (S1 ? kUnrollMax : kUnrollWide) : (S1 ? kUnrollWide : kUnroll);   // strided: 2U taps per batch
  float acc[2] = {0.f, 0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const float* qb = p.q + ((size_t)k.n * Ti * V + v) * C + k.c;
    const size_t ob = ((size_t)k.n * To * V + v) * C + k.c;
    float qa = S1 ? tap(qb, to0 + L.y1, Ti, pitch) : 0.f;
    for (int to = to0; to < to1; to += U) {
      float q0[S1 ? 1 : U], q1[U], rv[MODE == 1 ? U : 1];
      if (S1) {
        load_frames<U, PITCH>(q1, qb, to + L.y1 + 1, Ti, pitch);
      } else if (st == 2) {
        // stride 2: output frame to + u taps input frames 2(to + u) + y1 and + 1 -- 2U consecutive frames, one base
        // address and immediate offsets like the stride-1 walk (frames past the chunk are loaded but never used)
        float qq[2 * U];
        load_frames<2 * U, PITCH>(qq, qb, 2 * to + L.y1, Ti, pitch);
#pragma unroll
        for (int u = 0; u < U; ++u) q0[S1 ? 0 : u] = qq[2 * u], q1[u] = qq[2 * u + 1];
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ta = min(to + u, to1 - 1) * st + L.y1;
          q1[u] = tap(qb, ta + 1, Ti, pitch);
          q0[S1 ? 0 : u] = tap(qb, ta, Ti, pitch);
        }
      }
      if (MODE == 1) load_frames<(MODE == 1 ? U : 1), PITCH>(rv, MODE == 1 && p.res ? p.res + ob : qb, to, MODE == 1 && p.res ? To : Ti, pitch);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (to + u < to1) {
          const float a = S1 ? qa : q0[S1 ? 0 : u];
          const float s = fmaf(L.f, q1[u], L.g * a);
          qa = q1[u];
          if (MODE == 0) {
            acc[0] += s;
            acc[1] = fmaf(s, s, acc[1]);
          } else {
            float y = fmaf(s, sc, sh);
            if (p.res) y += rv[MODE == 1 ? u : 0];
            if (p.relu) y = fmaxf(y, 0.f);
            if constexpr (POOL) {                                  // head of the model: pooled sums, output optional
              if (p.out) p.out[ob + (size_t)(to + u) * pitch] = y;
              acc[0] += y;
            } else {
              p.out[ob + (size_t)(to + u) * pitch] = y;
            }
          }
        }
    }
  }
  if (MODE == 0) block_reduce_channels<2>(acc, p.stats, k.c, scratch);
  if constexpr (MODE == 1 && POOL) {                               // pooled sums of sample k.n
    const float a1[1] = {acc[0]};
    block_reduce_channels<1>(a1, p.stats + (size_t)k.n * C, k.c, scratch, 1);
  }
}

// ------------------------------------------------------------------------------------------------ output shift + BN, backward
// sums5[c] += { g, g*shat, g*dq, dq, shat*dq }   with g = gy*[y>0] (if relu), s = Shift(q), shat = (s-mean)*invstd,
// dq = Q(ta+1) - Q(ta)  (d s / d ypos, K4 with xpos = 0)
template <bool S1, bool RELU, int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_bwd_stats_kernel(const SgcnTShiftBwd p, int tper, int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32 * 5];
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, Ti = p.T_in, To = p.T_out, st = S1 ? 1 : p.stride;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const int to0 = k.chunk * tper, to1 = min(To, to0 + tper);
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};                       // { g, g*s, g*dq, dq, s*dq }: shat applied after the walk
  for (int v = k.warp; v < V; v += k.nw) {
    const float* qb = p.q + ((size_t)k.n * Ti * V + v) * C + k.c;
    const size_t ob = ((size_t)k.n * To * V + v) * C + k.c;
    const float* yb = p.relu ? p.y + ob : p.gy + ob;            // without ReLU the mask source is irrelevant
    float qa = S1 ? tap(qb, to0 + L.y1, Ti, pitch) : 0.f;
    for (int to = to0; to < to1; to += kUnrollWide) {
      float q0[S1 ? 1 : kUnrollWide], q1[kUnrollWide], gv[kUnrollWide], yv[kUnrollWide];
      if (S1) {
        load_frames<kUnrollWide, PITCH>(q1, qb, to + L.y1 + 1, Ti, pitch);
      } else if (st == 2) {                                       // 2U consecutive frames (see tshift_fwd_kernel)
        float qq[2 * kUnrollWide];
        load_frames<2 * kUnrollWide, PITCH>(qq, qb, 2 * to + L.y1, Ti, pitch);
#pragma unroll
        for (int u = 0; u < kUnrollWide; ++u) q0[S1 ? 0 : u] = qq[2 * u], q1[u] = qq[2 * u + 1];
      } else {
#pragma unroll
        for (int u = 0; u < kUnrollWide; ++u) {
          const int ta = min(to + u, to1 - 1) * st + L.y1;
          q1[u] = tap(qb, ta + 1, Ti, pitch);
          q0[S1 ? 0 : u] = tap(qb, ta, Ti, pitch);
        }
      }
      load_frames<kUnrollWide, PITCH>(gv, p.gy + ob, to, To, pitch);
      if (RELU) {                                          // compile time: a pre-masked g_y has no third stream
        load_frames<kUnrollWide, PITCH>(yv, yb, to, To, pitch);
      } else {
#pragma unroll
        for (int u = 0; u < kUnrollWide; ++u) yv[u] = 1.f;
      }
#pragma unroll
      for (int u = 0; u < kUnrollWide; ++u)
        if (to + u < to1) {
          const float a = S1 ? qa : q0[S1 ? 0 : u];
          const float s = fmaf(L.f, q1[u], L.g * a);
          const float dq = q1[u] - a;
          qa = q1[u];
          const float g = yv[u] > 0.f ? gv[u] : 0.f;
          acc[0] += g;
          acc[1] = fmaf(g, s, acc[1]);
          acc[2] = fmaf(g, dq, acc[2]);
          acc[3] += dq;
          acc[4] = fmaf(s, dq, acc[4]);
        }
    }
  }
  {   // sum x*shat = invstd * (sum x*s - mean * sum x)
    const float mean = __ldg(p.mean + k.c), invstd = __ldg(p.invstd + k.c);
    acc[1] = invstd * (acc[1] - mean * acc[0]);
    acc[4] = invstd * (acc[4] - mean * acc[3]);
  }
  block_reduce_channels<5>(acc, p.sums, k.c, scratch);
}

struct BwdCh {
  float k1, ka, kb;      // ds = k1*(g - m1 - shat*m2) = k1*g + ka*s + kb
};
__device__ __forceinline__ BwdCh bwd_of(const SgcnTShiftBwd& p, int c) {
  const float mean = __ldg(p.mean + c), invstd = __ldg(p.invstd + c), k1 = __ldg(p.k1 + c), m1 = __ldg(p.m1 + c),
              m2 = __ldg(p.m2 + c);
  const float ka = -k1 * invstd * m2;
  return {k1, ka, -k1 * m1 - ka * mean};
}
// ds(to) = k1*(g - m1 - shat*m2) for an output frame inside [0, T_out), else 0
__device__ __forceinline__ float ds_eval(bool valid, float g, float y, int relu, float qa, float qb, const LerpCh& L,
                                         const BwdCh& B) {
  if (!valid) return 0.f;
  if (relu && !(y > 0.f)) g = 0.f;
  const float s = fmaf(L.f, qb, L.g * qa);
  return fmaf(B.k1, g, fmaf(B.ka, s, B.kb));
}

// stride 1:  dpre(t) = [Q(t) > 0] * ( g*ds(t-y1) + f*ds(t-y1-1) ),  ds(to) uses s(to) = g*Q(to+y1) + f*Q(to+y1+1)
// walk over input frames t; to = t - y1; the previous ds and the tap Q(t+1) slide along in registers
template <bool RELU, int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_bwd_apply_s1_kernel(const SgcnTShiftBwd p, int tper,
                                                                            int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32];
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, T = p.T_in;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const BwdCh B = bwd_of(p, k.c);
  const int relu = RELU ? 1 : 0;
  constexpr int U = RELU ? kUnroll : kUnrollWide;             // 3 or 2 loads per frame
  const int t0 = k.chunk * tper, t1 = min(T, t0 + tper);
  float acc[1] = {0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const size_t cb = ((size_t)k.n * T * V + v) * C + k.c;
    const float* qb = p.q + cb;
    const float* gb = p.gy + cb;
    const float* yb = relu ? p.y + cb : gb;
    float* db = p.dpre + cb;
    // warm-up: ds of output frame to = t0 - y1 - 1
    float qa = tap(qb, t0, T, pitch);
    float ds_prev;
    {
      const int to = t0 - L.y1 - 1;
      ds_prev = ds_eval(inside(to, T), ldrow(gb, to, T, pitch), ldrow(yb, to, T, pitch), relu,
                        tap(qb, t0 - 1, T, pitch), qa, L, B);
    }
    for (int t = t0; t < t1; t += U) {
      float q1[U], gv[U], yv[U];
      load_frames<U, PITCH>(q1, qb, t + 1, T, pitch);
      load_frames<U, PITCH>(gv, gb, t - L.y1, T, pitch);      // frames outside [0, T) are masked by ds_eval
      if (RELU) {                                                   // compile time, see tshift_bwd_stats_kernel
        load_frames<U, PITCH>(yv, yb, t - L.y1, T, pitch);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) yv[u] = 1.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (t + u < t1) {
          const int to = t + u - L.y1;
          const float ds = ds_eval((unsigned)to < (unsigned)T, gv[u], yv[u], relu, qa, q1[u], L, B);
          const float d = (qa > 0.f) ? fmaf(L.f, ds_prev, L.g * ds) : 0.f;
          db[(size_t)(t + u) * pitch] = d;
          acc[0] += d;
          ds_prev = ds;
          qa = q1[u];
        }
    }
  }
  block_reduce_channels<1>(acc, p.dbias, k.c, scratch);
}

// stride 2 (K3): output frame to feeds input frames t = 2*to + y1 (weight g) and t + 1 (weight f); frames that no
// output frame touches get 0.  Walk over output frames.
template <bool RELU, int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_bwd_apply_s2_kernel(const SgcnTShiftBwd p, int tper,
                                                                            int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32];
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, Ti = p.T_in, To = p.T_out;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const BwdCh B = bwd_of(p, k.c);
  const int relu = RELU ? 1 : 0;
  const int to0 = k.chunk * tper, to1 = min(To, to0 + tper);
  float acc[1] = {0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const float* qb = p.q + ((size_t)k.n * Ti * V + v) * C + k.c;
    float* db = p.dpre + ((size_t)k.n * Ti * V + v) * C + k.c;
    const size_t ob = ((size_t)k.n * To * V + v) * C + k.c;
    const float* yb = relu ? p.y + ob : p.gy + ob;
    if (k.chunk == 0)                                     // frames before the first touched one
      for (int t = 0; t < min(Ti, L.y1); ++t) db[(size_t)t * pitch] = 0.f;
    if (k.chunk == nchunks - 1)                           // frames after the last touched one
      for (int t = max(0, 2 * To + L.y1); t < Ti; ++t) db[(size_t)t * pitch] = 0.f;
    for (int to = to0; to < to1; to += kUnroll) {
      // output frames to .. to+U-1 tap (and feed) the 2U CONSECUTIVE input frames 2*to + y1 ..: one base address and
      // immediate offsets for the loads, and for the stores of a batch that lies inside the chunk and the sequence
      float qq[2 * kUnroll], gv[kUnroll], yv[kUnroll];
      const int ta0 = 2 * to + L.y1;
      load_frames<2 * kUnroll, PITCH>(qq, qb, ta0, Ti, pitch);
      load_frames<kUnroll, PITCH>(gv, p.gy + ob, to, To, pitch);
      if (RELU) {
        load_frames<kUnroll, PITCH>(yv, yb, to, To, pitch);
      } else {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) yv[u] = 1.f;
      }
      const bool whole = to + kUnroll <= to1 && ta0 >= 0 && ta0 + 2 * kUnroll <= Ti;
      float* dq = db + (size_t)ta0 * (size_t)(PITCH ? PITCH : pitch);      // (only dereferenced inside the sequence)
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (whole || to + u < to1) {
          const int ta = ta0 + 2 * u;
          const float ds = ds_eval(true, gv[u], yv[u], relu, qq[2 * u], qq[2 * u + 1], L, B);
          if (whole || (unsigned)ta < (unsigned)Ti) {
            const float d = (qq[2 * u] > 0.f) ? L.g * ds : 0.f;
            dq[(size_t)(2 * u) * (size_t)(PITCH ? PITCH : pitch)] = d;
            acc[0] += d;
          }
          if (whole || (unsigned)(ta + 1) < (unsigned)Ti) {
            const float d = (qq[2 * u + 1] > 0.f) ? L.f * ds : 0.f;
            dq[(size_t)(2 * u + 1) * (size_t)(PITCH ? PITCH : pitch)] = d;
            acc[0] += d;
          }
        }
    }
  }
  block_reduce_channels<1>(acc, p.dbias, k.c, scratch);
}

// ------------------------------------------------------------------------------------------------ BN + input shift, backward
// p = Shift_1(U), U = BN(h) zero padded; du = Shift_1^T(dp).  Re-indexed over the frame r of dp:
//   sum_t du            = sum_r dp(r) * ( g*[r+y1 in range] + f*[r+y1+1 in range] )
//   sum_t du * hhat(t)  = sum_r dp(r) * ( g*hhat(r+y1) + f*hhat(r+y1+1) )          (hhat zero padded)
//   d/dypos             = sum_r dp(r) * ( U(r+y1+1) - U(r+y1) )
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_in_bwd_stats_kernel(const SgcnTShiftInBwd p, int tper,
                                                                            int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32 * 3];
  if (p.gate && *p.gate == 0) return;                          // sgcn_tshift_in_bwd_sums already produced the sums
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, T = p.T;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const float mean = __ldg(p.mean + k.c), invstd = __ldg(p.invstd + k.c), sc = __ldg(p.scale + k.c),
              sh = __ldg(p.shift + k.c);
  const int t0 = k.chunk * tper, t1 = min(T, t0 + tper);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const size_t cb = ((size_t)k.n * T * V + v) * C + k.c;
    const float* dpb = p.dp + cb;
    const float* hb = p.h + cb;
    float ha = ldrow(hb, t0 + L.y1, T, pitch);               // validity is applied per use below
    for (int r = t0; r < t1; r += kUnrollWide) {
      float dv[kUnrollWide], h1[kUnrollWide];
#pragma unroll
      for (int u = 0; u < kUnrollWide; ++u) {
        const int rc = min(r + u, t1 - 1);
        dv[u] = __ldg(dpb + (size_t)rc * pitch);
        h1[u] = ldrow(hb, rc + L.y1 + 1, T, pitch);
      }
#pragma unroll
      for (int u = 0; u < kUnrollWide; ++u)
        if (r + u < t1) {
          const int ta = r + u + L.y1;
          const bool ia = (unsigned)ta < (unsigned)T, ib = (unsigned)(ta + 1) < (unsigned)T;
          const float ua = ia ? fmaf(sc, ha, sh) : 0.f, ub = ib ? fmaf(sc, h1[u], sh) : 0.f;
          const float na = ia ? (ha - mean) * invstd : 0.f, nb = ib ? (h1[u] - mean) * invstd : 0.f;
          acc[0] = fmaf(dv[u], (ia ? L.g : 0.f) + (ib ? L.f : 0.f), acc[0]);
          acc[1] = fmaf(dv[u], fmaf(L.f, nb, L.g * na), acc[1]);
          acc[2] = fmaf(dv[u], ub - ua, acc[2]);
          ha = h1[u];
        }
    }
  }
  block_reduce_channels<3>(acc, p.sums, k.c, scratch);
}

// gh(t) = [h(t) > 0] * k*(du(t) - m1 - hhat(t)*m2),  du(t) = g*dp(t-y1) + f*dp(t-y1-1);
// per-(v,c) sums for the BN1d backward of the spatial unit: { gh, gh * zhat }
template <int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_in_bwd_apply_kernel(const SgcnTShiftInBwd p, int tper,
                                                                            int nchunks, int rev) {
  __shared__ float scratch[kMaxWarps * 32];
  const Col k = col_of(p.C, nchunks, rev);
  const int C = p.C, V = p.V, T = p.T;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  // k*(du - m1 - hhat*m2) as the affine map  kk*du + ka*h + kb  (three registers instead of five)
  const float kk = __ldg(p.k1 + k.c);
  float ka, kb;
  {
    const float mean = __ldg(p.mean + k.c), invstd = __ldg(p.invstd + k.c), m1 = __ldg(p.m1 + k.c), m2 = __ldg(p.m2 + k.c);
    ka = -kk * invstd * m2;
    kb = -kk * m1 - ka * mean;
  }
  const int t0 = k.chunk * tper, t1 = min(T, t0 + tper);
  const bool has_z = p.z != nullptr;
  // position gradient (K4): sum_r dp(r) * (U(r+y1+1) - U(r+y1)), U = sc*h + sh zero padded, re-indexed over
  // t = r+y1+1 so that it needs only the walker's own taps:  sc * sum_t dp(t-y1-1) * (h(t) - h(t-1))  [h(-1) = h(T) = 0]
  // + sh * (dp(-y1-1) - dp(T-y1-1)); the term of t = T is added by the chunk that ends the sequence
  float acc[1] = {0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const size_t cb = ((size_t)k.n * T * V + v) * C + k.c;
    const float* dpb = p.dp + cb;
    const float* hb = p.h + cb;
    const float* zb = has_z ? p.z + cb : hb;
    float* gb = p.gh + cb;
    const float zinv = has_z ? __ldg(p.zinvstd + v * C + k.c) : 0.f;
    const float zoff = has_z ? -__ldg(p.zmean + v * C + k.c) * zinv : 0.f;      // zhat = z*zinv + zoff
    float s0 = 0.f, s1 = 0.f;
    float dprev = tap(dpb, t0 - L.y1 - 1, T, pitch);
    float hprev = t0 > 0 ? ldrow(hb, t0 - 1, T, pitch) : 0.f;
    float ah = 0.f;
    for (int t = t0; t < t1; t += kUnroll) {
      float d0[kUnroll], hv[kUnroll], zv[kUnroll];
      load_frames<kUnroll, PITCH>(d0, dpb, t - L.y1, T, pitch);
      load_frames<kUnroll, PITCH>(hv, hb, t, T, pitch);
      load_frames<kUnroll, PITCH>(zv, zb, t, T, pitch);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (t + u < t1) {
          const float du = fmaf(L.f, dprev, L.g * d0[u]);
          ah = fmaf(dprev, hv[u] - hprev, ah);
          hprev = hv[u];
          dprev = d0[u];
          float g = fmaf(kk, du, fmaf(ka, hv[u], kb));
          if (p.relu_h && !(hv[u] > 0.f)) g = 0.f;
          gb[(size_t)(t + u) * pitch] = g;
          s0 += g;
          s1 = fmaf(g, fmaf(zv[u], zinv, zoff), s1);
        }
    }
    {
      float edge = t0 == 0 ? tap(dpb, -L.y1 - 1, T, pitch) : 0.f;   // re-loaded: not worth a register across the walk
      if (t1 == T) {                                              // t = T: h(T) = 0, dprev = dp(T-1-y1)
        ah = fmaf(dprev, -hprev, ah);
        edge -= dprev;
      }
      acc[0] += fmaf(__ldg(p.scale + k.c), ah, __ldg(p.shift + k.c) * edge);
    }
    if (has_z) {
      atomicAdd(p.vd_sums + 2 * ((size_t)v * C + k.c), (double)s0);
      atomicAdd(p.vd_sums + 2 * ((size_t)v * C + k.c) + 1, (double)s1);
    }
  }
  block_reduce_channels<1>(acc, p.pos_sums, k.c, scratch);
}

// ------------------------------------------------------------------------------------------------ algebraic BN(h) backward sums
// gate = 1 when a BatchNorm weight is too small for the division below (the exact pass runs instead)
__global__ void tshift_in_gate_kernel(const float* __restrict__ scale, const float* __restrict__ invstd, int* gate, int C) {
  const int c = threadIdx.x;
  const int bad = (c < C) && !(fabsf(__ldg(scale + c)) >= 1e-3f * fabsf(__ldg(invstd + c)));
  const int any = __syncthreads_or(bad);
  if (threadIdx.x == 0) *gate = any;
}

// sums[c][0] -= sum over the frames r whose shift taps leave the sequence of dp(r) * (1 - e_c(r)),
// e_c(r) = g*[0 <= r+y1 < T] + f*[0 <= r+y1+1 < T]:  at most |y1|+1 frames at one end of every sample
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) tshift_in_boundary_kernel(const SgcnTShiftInSums p) {
  __shared__ float scratch[kMaxWarps * 32];
  if (*p.gate) return;
  const Col k = col_of(p.C, 1, 0);
  const int C = p.C, V = p.V, T = p.T;
  const int pitch = V * C;
  const LerpCh L = lerp_of(__ldg(p.ypos_eff + k.c));
  const int r0 = L.y1 >= 0 ? max(0, T - L.y1 - 1) : 0, r1 = L.y1 >= 0 ? T : min(T, -L.y1);
  float acc[1] = {0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const float* dpb = p.dp + ((size_t)k.n * T * V + v) * C + k.c;
    for (int r = r0; r < r1; ++r) {
      const float w = (inside(r + L.y1, T) ? 0.f : L.g) + (inside(r + L.y1 + 1, T) ? 0.f : L.f);
      acc[0] = fmaf(__ldg(dpb + (size_t)r * pitch), -w, acc[0]);
    }
  }
  block_reduce_channels<1>(acc, p.sums, k.c, scratch, 3);       // sums[c][0]
}

__global__ void __launch_bounds__(1024) tshift_in_combine_kernel(const SgcnTShiftInSums p) {
  __shared__ double part[2][16][64];
  const int cl = threadIdx.x & 63, dg = threadIdx.x >> 6;         // 64 channels x 16 interleaved row groups
  const int c = blockIdx.x * 64 + cl, C = p.C;
  if (*p.gate) return;
  double s0 = 0.0, sw = 0.0;
  for (int d = dg; d < C; d += 16) {
    const double w = (double)__ldg(p.Wt + (size_t)d * C + c);
    s0 = fma(w, (double)__ldg(p.dbt + d), s0);
    sw = fma(w, (double)__ldg(p.dWt + (size_t)d * C + c), sw);
  }
  part[0][dg][cl] = s0;
  part[1][dg][cl] = sw;
  __syncthreads();
  if (dg != 0) return;
  s0 = sw = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    s0 += part[0][j][cl];
    sw += part[1][j][cl];
  }
  s0 += p.sums[3 * (size_t)c];                                   // minus the boundary frames (tshift_in_boundary_kernel)
  const double duh = (sw - (double)p.shift[c] * s0) / (double)p.scale[c];   // sum du * h
  p.sums[3 * (size_t)c] = s0;
  p.sums[3 * (size_t)c + 1] = (double)p.invstd[c] * (duh - (double)p.mean[c] * s0);
}

// ------------------------------------------------------------------------------------------------ group walkers
// stats[c] += { sum x, sum x^2 }   (BatchNorm2d statistics of a stand-alone Shift_tcn input, model/shift_gcn.py:66)
template <int PITCH>
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) channel_stats_kernel(const float* __restrict__ x,
                                                                       double* __restrict__ stats, long long groups,
                                                                       int gper, int nchunks, int V, int C, int gs, int rev) {
  __shared__ float scratch[kMaxWarps * 32 * 2];
  const Col k = col_of(C, nchunks, rev);
  const long long g0 = (long long)k.chunk * gper;
  const int ng = (int)((groups - g0) < gper ? (groups - g0) : gper);
  const int pitch = V * C * gs;                                  // gs > 1: every gs-th group (frame) only
  float acc[2] = {0.f, 0.f};
  for (int v = k.warp; v < V; v += k.nw) {
    const float* xp = x + ((size_t)g0 * gs * V + v) * C + k.c;
    for (int g = 0; g < ng; g += kUnrollMax) {
      float xv[kUnrollMax];
      load_frames<kUnrollMax, PITCH>(xv, xp, g, ng, pitch);        // zero padded past the chunk
#pragma unroll
      for (int u = 0; u < kUnrollMax; ++u) {
        acc[0] += xv[u];
        acc[1] = fmaf(xv[u], xv[u], acc[1]);
      }
    }
  }
  block_reduce_channels<2>(acc, stats, k.c, scratch);
}

// gh = g * [h > 0];  vd_sums[v,c] += { gh, gh * zhat }      (stand-alone Shift_gcn backward, model/shift_gcn.py:137-141)
__global__ void __launch_bounds__(kMaxWarps * 32, SGCN_WALKER_MINBLOCKS) relu_bn1d_bwd_stats_kernel(const float* __restrict__ g,
                                                                             const float* __restrict__ h,
                                                                             const float* __restrict__ z,
                                                                             const float* __restrict__ zmean,
                                                                             const float* __restrict__ zinvstd,
                                                                             float* __restrict__ gh,
                                                                             double* __restrict__ vd_sums,
                                                                             long long groups, int gper, int nchunks,
                                                                             int V, int C, int rev) {
  const Col k = col_of(C, nchunks, rev);
  const long long g0 = (long long)k.chunk * gper;
  const int ng = (int)((groups - g0) < gper ? (groups - g0) : gper);
  const int pitch = V * C;
  for (int v = k.warp; v < V; v += k.nw) {
    const float zm = __ldg(zmean + v * C + k.c), zi = __ldg(zinvstd + v * C + k.c);
    const size_t o0 = ((size_t)g0 * V + v) * C + k.c;
    float s0 = 0.f, s1 = 0.f;
    for (int gi = 0; gi < ng; gi += kUnroll) {
      float gv[kUnroll], hv[kUnroll], zv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t o = o0 + (size_t)min(gi + u, ng - 1) * pitch;
        gv[u] = __ldg(g + o);
        hv[u] = __ldg(h + o);
        zv[u] = __ldg(z + o);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (gi + u < ng) {
          const float val = (hv[u] > 0.f) ? gv[u] : 0.f;
          gh[o0 + (size_t)(gi + u) * pitch] = val;
          s0 += val;
          s1 = fmaf(val, (zv[u] - zm) * zi, s1);
        }
    }
    atomicAdd(vd_sums + 2 * ((size_t)v * C + k.c), (double)s0);
    atomicAdd(vd_sums + 2 * ((size_t)v * C + k.c) + 1, (double)s1);
  }
}

// out = g * [y > 0]
__global__ void __launch_bounds__(256) relu_mask_grad_kernel(const float4* __restrict__ g, const float4* __restrict__ y,
                                                             float4* __restrict__ out, long long n4, int rev) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (long long)gridDim.x * blockDim.x) {
    const long long i = rev ? n4 - 1 - j : j;
    const float4 a = g[i], b = y[i];
    out[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
  }
}

// out[n, r, c] = g[n, c] * scale for r in [0, rows_per_n): gradient of the global mean pooling (model/shift_gcn.py:212-214)
// written straight into the row layout (autograd would expand, scale and copy: three passes over a full tensor)
// mask_y (optional, same shape as out): out *= [mask_y > 0] -- the rows are the output y of a unit that ends in a ReLU, and
// its backward then finds the gradient already masked (functional._links) instead of reading y in three kernels
__global__ void __launch_bounds__(256) bcast_rows_kernel(const float4* __restrict__ g, float4* __restrict__ out,
                                                         const float4* __restrict__ mask_y, long long rows_per_n, int c4,
                                                         float scale, long long total4) {
  const long long per_n = rows_per_n * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / per_n;
    const int c = (int)(i % c4);
    float4 v = __ldg(g + n * c4 + c);
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    if (mask_y) {
      const float4 y = __ldg(mask_y + i);
      v.x = y.x > 0.f ? v.x : 0.f, v.y = y.y > 0.f ? v.y : 0.f, v.z = y.z > 0.f ? v.z : 0.f, v.w = y.w > 0.f ? v.w : 0.f;
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ launch geometry
struct Geo {
  int threads, nchunks, per;
  unsigned grid;
};
// `outer` independent walks (samples) of `len` steps each, split into chunks so that the grid has a few waves
static Geo geometry(int C, int V, long long outer, long long len, int min_per) {
  Geo g;
  g.threads = 32 * ceil_div(V, 2);
  const long long base = (long long)(C / 32) * (outer > 0 ? outer : 1);
  const long long want = (long long)num_sms() * SGCN_GEOM_MULT;   // ~2 waves of 4 resident blocks per SM
  long long nch = (want + base - 1) / base;
  const long long max_ch = (len + min_per - 1) / min_per;
  if (nch > max_ch) nch = max_ch;
  if (nch < 1) nch = 1;
  g.per = (int)((len + nch - 1) / nch);
  g.nchunks = (int)((len + g.per - 1) / g.per);
  g.grid = (unsigned)(base * g.nchunks);
  return g;
}

// compile-time frame pitch (V*C floats) for the shapes of the reference's two skeletons (25 / 33 joints) and of the
// pseudo-groups of sgcn_channel_stats (32 rows); anything else takes the run-time pitch
#define SGCN_PITCH_CASE(N, ...) \
  case N: {                     \
    constexpr int P = N;        \
    __VA_ARGS__;                \
  } break;
#define SGCN_PITCH_DISPATCH(pitch, ...)                                                             \
  switch (pitch) {                                                                                  \
    SGCN_PITCH_CASE(1600, __VA_ARGS__) SGCN_PITCH_CASE(3200, __VA_ARGS__) SGCN_PITCH_CASE(6400, __VA_ARGS__) \
    SGCN_PITCH_CASE(2112, __VA_ARGS__) SGCN_PITCH_CASE(4224, __VA_ARGS__) SGCN_PITCH_CASE(8448, __VA_ARGS__) \
    SGCN_PITCH_CASE(2048, __VA_ARGS__) SGCN_PITCH_CASE(4096, __VA_ARGS__) SGCN_PITCH_CASE(8192, __VA_ARGS__) \
    default: {                                                                                      \
      constexpr int P = 0;                                                                          \
      __VA_ARGS__;                                                                                  \
    }                                                                                               \
  }

static int check_cv(int C, int V) {
  if (C != 64 && C != 128 && C != 256) return set_error("pointwise: channel count must be 64, 128 or 256");
  if (V < 1 || V > 2 * kMaxWarps) return set_error("pointwise: num_point must be in [1, 40]");
  return 0;
}

}  // namespace sgcn

using namespace sgcn;

extern "C" int sgcn_bn_res_relu_fwd(const float* z, const float* res, float* h, const float* scale, const float* shift,
                                    double* stats_out, long long rows, int V, int D, int relu, void* stream) {
  if (!z || !h || !scale || !shift) return set_error("sgcn_bn_res_relu_fwd: null pointer");
  if (int rc = check_cv(D, V)) return rc;
  if (rows <= 0) return 0;
  if (rows % V != 0) return set_error("sgcn_bn_res_relu_fwd: rows must be a multiple of V");
  const long long groups = rows / V;
#ifndef SGCN_BRR_VEC
#define SGCN_BRR_VEC 1
#endif
  if (SGCN_BRR_VEC) {                                              // 128-bit streaming variant
    const int cols = V * D / 4;
    int threads = 256;
    while (cols % threads) --threads;                              // 200 (V = 25) / 176 (V = 33): every thread has a column
    const int cblocks = cols / threads;
    long long nch = ((long long)num_sms() * 16 + cblocks - 1) / cblocks;    // ~4 waves of 4 resident blocks per SM
    if (nch > groups) nch = groups;
    const int gper = (int)((groups + nch - 1) / nch);
    nch = (groups + gper - 1) / gper;
    const int rev = next_direction();
    bn_res_relu_fwd4_kernel<<<(unsigned)(nch * cblocks), threads, 2 * D * sizeof(float), (cudaStream_t)stream>>>(
        (const float4*)z, (const float4*)res, (float4*)h, (const float4*)scale, (const float4*)shift, stats_out, groups, gper,
        cols, D, relu, rev);
    return check_launch("bn_res_relu_fwd4_kernel");
  }
  const Geo g = geometry(D, V, 1, groups, 8);
  const int rev = next_direction();
  SGCN_PITCH_DISPATCH(V * D, (bn_res_relu_fwd_kernel<P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(
                                 z, res, h, scale, shift, stats_out, groups, g.per, g.nchunks, V, D, relu, rev)))
  return check_launch("bn_res_relu_fwd_kernel");
}

extern "C" int sgcn_tshift_fwd(const SgcnTShift* p, int mode, void* stream) {
  if (!p || !p->q || !p->ypos_eff) return set_error("sgcn_tshift_fwd: null pointer");
  if (int rc = check_cv(p->C, p->V)) return rc;
  if (p->stride < 1 || p->T_out != p->T_in / p->stride) return set_error("sgcn_tshift_fwd: T_out must be T_in / stride");
  if (p->n_samples <= 0 || p->T_out <= 0) return 0;
  const Geo g = geometry(p->C, p->V, p->n_samples, p->T_out, 8);
  const int rev = next_direction();
  if (mode == 0) {
    if (!p->stats) return set_error("sgcn_tshift_fwd(stats): null stats");
    if (p->stride == 1) {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_fwd_kernel<0, true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_fwd_kernel<0, false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    }
  } else {
    if ((!p->out && !p->stats) || !p->scale || !p->shift) return set_error("sgcn_tshift_fwd(apply): null pointer");
    if (p->stats) {                                                // with the pooled sums of the model head (last unit)
      if (p->stride != 1) return set_error("sgcn_tshift_fwd(apply): pooled sums need stride 1");
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_fwd_kernel<1, true, P, true><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else if (p->stride == 1) {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_fwd_kernel<1, true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_fwd_kernel<1, false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    }
  }
  return check_launch("tshift_fwd_kernel");
}

extern "C" int sgcn_tshift_bwd(const SgcnTShiftBwd* p, int mode, void* stream) {
  if (!p || !p->q || !p->gy || !p->ypos_eff || !p->mean || !p->invstd) return set_error("sgcn_tshift_bwd: null pointer");
  if (p->relu && !p->y) return set_error("sgcn_tshift_bwd: relu mask needs y");
  if (int rc = check_cv(p->C, p->V)) return rc;
  if (p->stride < 1 || p->T_out != p->T_in / p->stride) return set_error("sgcn_tshift_bwd: T_out must be T_in / stride");
  if (p->n_samples <= 0) return 0;
  const int rev = next_direction();
  if (mode == 0) {
    if (!p->sums) return set_error("sgcn_tshift_bwd(stats): null sums");
    if (p->T_out <= 0) return 0;
    const Geo g = geometry(p->C, p->V, p->n_samples, p->T_out, 8);
    if (p->stride == 1) {
      if (p->relu) {
        SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_stats_kernel<true, true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
      } else {
        SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_stats_kernel<true, false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
      }
    } else if (p->relu) {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_stats_kernel<false, true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_stats_kernel<false, false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    }
    return check_launch("tshift_bwd_stats_kernel");
  }
  if (!p->dpre || !p->dbias || !p->k1 || !p->m1 || !p->m2) return set_error("sgcn_tshift_bwd(apply): null pointer");
  if (p->stride == 1) {
    const Geo g = geometry(p->C, p->V, p->n_samples, p->T_in, 8);
    if (p->relu) {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_apply_s1_kernel<true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_apply_s1_kernel<false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    }
    return check_launch("tshift_bwd_apply_s1_kernel");
  }
  if (p->stride == 2) {   // the reference's backward exists for strides 1 and 2 only (shift_cuda_kernel.cu:156-256)
    if (p->T_out <= 0) return set_error("sgcn_tshift_bwd(apply): empty output");
    const Geo g = geometry(p->C, p->V, p->n_samples, p->T_out, 8);
    if (p->relu) {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_apply_s2_kernel<true, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    } else {
      SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_bwd_apply_s2_kernel<false, P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
    }
    return check_launch("tshift_bwd_apply_s2_kernel");
  }
  return set_error("sgcn_tshift_bwd(apply): stride must be 1 or 2 (as in the reference's backward kernels)");
}

extern "C" int sgcn_tshift_in_bwd(const SgcnTShiftInBwd* p, int mode, void* stream) {
  if (!p || !p->dp || !p->h || !p->ypos_eff || !p->mean || !p->invstd) return set_error("sgcn_tshift_in_bwd: null pointer");
  if (int rc = check_cv(p->C, p->V)) return rc;
  if (p->n_samples <= 0 || p->T <= 0) return 0;
  const int rev = next_direction();
  if (mode == 0) {
    if (!p->sums || !p->scale || !p->shift) return set_error("sgcn_tshift_in_bwd(stats): null pointer");
    const Geo g = geometry(p->C, p->V, p->n_samples, p->T, 8);
    tshift_in_bwd_stats_kernel<<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev);
    return check_launch("tshift_in_bwd_stats_kernel");
  }
  if (!p->gh || !p->k1 || !p->m1 || !p->m2) return set_error("sgcn_tshift_in_bwd(apply): null pointer");
  if (!p->pos_sums || !p->scale || !p->shift) return set_error("sgcn_tshift_in_bwd(apply): null pos_sums / scale / shift");
  if (p->z && (!p->zmean || !p->zinvstd || !p->vd_sums)) return set_error("sgcn_tshift_in_bwd(apply): null BN1d tables");
  const Geo g = geometry(p->C, p->V, p->n_samples, p->T, 16);
  SGCN_PITCH_DISPATCH(p->V * p->C, (tshift_in_bwd_apply_kernel<P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(*p, g.per, g.nchunks, rev)))
  return check_launch("tshift_in_bwd_apply_kernel");
}

extern "C" int sgcn_tshift_in_bwd_sums(const SgcnTShiftInSums* p, void* stream) {
  if (!p || !p->dp || !p->ypos_eff || !p->Wt || !p->dWt || !p->dbt || !p->mean || !p->invstd || !p->scale || !p->shift ||
      !p->sums || !p->gate)
    return set_error("sgcn_tshift_in_bwd_sums: null pointer");
  if (int rc = check_cv(p->C, p->V)) return rc;
  if (p->n_samples <= 0 || p->T <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  tshift_in_gate_kernel<<<1, 256, 0, s>>>(p->scale, p->invstd, p->gate, p->C);
  if (int rc = check_launch("tshift_in_gate_kernel")) return rc;
  const unsigned grid = (unsigned)((p->C / 32) * p->n_samples);
  tshift_in_boundary_kernel<<<grid, 32 * ceil_div(p->V, 2), 0, s>>>(*p);
  if (int rc = check_launch("tshift_in_boundary_kernel")) return rc;
  tshift_in_combine_kernel<<<p->C / 64, 1024, 0, s>>>(*p);
  return check_launch("tshift_in_combine_kernel");
}

extern "C" int sgcn_channel_stats(const float* x, double* stats, long long rows, int C, void* stream) {
  if (!x || !stats) return set_error("sgcn_channel_stats: null pointer");
  if (C != 64 && C != 128 && C != 256) return set_error("pointwise: channel count must be 64, 128 or 256");
  if (rows <= 0) return 0;
  // any row order works for per-channel statistics: walk pseudo-groups of 32 rows
  const int V = 32;
  const long long groups = rows / V;
  if (groups > 0) {
    const Geo g = geometry(C, V, 1, groups, 8);
    const int rev = next_direction();
    SGCN_PITCH_DISPATCH(V * C, (channel_stats_kernel<P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(x, stats, groups, g.per, g.nchunks, V, C, 1, rev)))
    if (int rc = check_launch("channel_stats_kernel")) return rc;
  }
  const int tail = (int)(rows - groups * V);
  if (tail > 0) {
    channel_stats_kernel<0><<<C / 32, 32 * ceil_div(tail, 2), 0, (cudaStream_t)stream>>>(x + (size_t)groups * V * C, stats,
                                                                                     1, 1, 1, tail, C, 1, 0);
    return check_launch("channel_stats_kernel(tail)");
  }
  return 0;
}

extern "C" int sgcn_channel_stats_groups(const float* x, double* stats, long long groups, int V, int C, int gs,
                                         void* stream) {
  if (!x || !stats) return set_error("sgcn_channel_stats_groups: null pointer");
  if (int rc = check_cv(C, V)) return rc;
  if (gs < 1) return set_error("sgcn_channel_stats_groups: group stride must be >= 1");
  if (groups <= 0) return 0;
  const Geo g = geometry(C, V, 1, groups, 8);
  const int rev = next_direction();
  SGCN_PITCH_DISPATCH(V * C * gs, (channel_stats_kernel<P><<<g.grid, g.threads, 0, (cudaStream_t)stream>>>(
                                      x, stats, groups, g.per, g.nchunks, V, C, gs, rev)))
  return check_launch("channel_stats_kernel(groups)");
}

extern "C" int sgcn_relu_bn1d_bwd_stats(const float* g, const float* h, const float* z, const float* zmean,
                                        const float* zinvstd, float* gh, double* vd_sums, long long groups, int V,
                                        int C, void* stream) {
  if (!g || !h || !z || !zmean || !zinvstd || !gh || !vd_sums) return set_error("sgcn_relu_bn1d_bwd_stats: null pointer");
  if (int rc = check_cv(C, V)) return rc;
  if (groups <= 0) return 0;
  const Geo geo = geometry(C, V, 1, groups, 16);
  relu_bn1d_bwd_stats_kernel<<<geo.grid, geo.threads, 0, (cudaStream_t)stream>>>(g, h, z, zmean, zinvstd, gh, vd_sums,
                                                                                groups, geo.per, geo.nchunks, V, C, next_direction());
  return check_launch("relu_bn1d_bwd_stats_kernel");
}

extern "C" int sgcn_bcast_rows(const float* g, float* out, const float* mask_y, long long n, long long rows_per_n, int C,
                               float scale, void* stream) {
  if (!g || !out) return set_error("sgcn_bcast_rows: null pointer");
  if (C < 4 || C % 4 != 0) return set_error("sgcn_bcast_rows: C must be a positive multiple of 4");
  if (n <= 0 || rows_per_n <= 0) return 0;
  const long long total4 = n * rows_per_n * (C / 4);
  long long blocks = (total4 + 1023) / 1024;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  bcast_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)g, (float4*)out, (const float4*)mask_y,
                                                                       rows_per_n, C / 4, scale, total4);
  return check_launch("bcast_rows_kernel");
}

extern "C" int sgcn_relu_mask_grad(const float* g, const float* y, float* out, long long numel, void* stream) {
  if (!g || !y || !out) return set_error("sgcn_relu_mask_grad: null pointer");
  if (numel % 4 != 0) return set_error("sgcn_relu_mask_grad: numel must be a multiple of 4");
  if (numel <= 0) return 0;
  long long blocks = (numel / 4 + 1023) / 1024;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  relu_mask_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)g, (const float4*)y,
                                                                           (float4*)out, numel / 4, next_direction());
  return check_launch("relu_mask_grad_kernel");
}
