// wgrad.cu -- weight-gradient contractions  dW[a, b] = sum_rows  A[row, a] * B[row, b]
// on tcgen05 tensor cores, operands built on the fly from the activations.
//
//   spatial  (autograd of model/shift_gcn.py:131):  A = xm  (gather + mask of x),  B = dy (BN1d-bwd + inverse rotation)
//            -> Linear_weight.grad [C, D]
//   temporal (autograd of model/shift_gcn.py:69):   A = dpre (grad wrt the conv output, pre-ReLU), B = p (shifted BN(h))
//            -> temporal_linear.weight.grad [Cout, Cin]
//
// Warp-specialised, persistent: one CTA per SM walks a strided subset of the row tiles and keeps its partial dW
// -- the WHOLE [CA x CB] matrix, up to 2 x 256 TMEM columns -- in tensor memory for the entire kernel, so every
// activation element is read from HBM exactly once; one coalesced atomic flush at the end.
//   warp 0            issues tcgen05.mma (one lane), owns the TMEM allocation
//   builder groups    of 8 warps; group g builds tiles g, g+NG, ... of this CTA into operand stage g:
//                     global loads (lane <-> channel, 128-byte coalesced, 20-40 independent loads in flight per thread
//                     -- the register file is the staging buffer) -> prologue math -> TF32 -> swizzled smem.
//                     The plain A operand of the temporal case is copied by cp.async straight into the swizzled
//                     layout (the tensor core reads fp32 bit patterns as TF32).
//   full[s] / free[s] mbarriers per stage: 256 builder arrivals / one tcgen05.commit
// While one group waits for its loads the others are computing or their stages are being consumed, so HBM always
// has about two tiles of requests in flight per SM.
//
// Both operands are MN-major (the row index is the contraction dimension, channels contiguous) in the
// SWIZZLE_128B_BASE32B layout -- the only one tcgen05 accepts for MN-major 32-bit operands: blocks of
// [KR rows x 32 channels], 128 B per row, 32-byte chunks XOR-ed with (row % 4).  The M extent of every MMA is 128:
// when A has only 64 channels the upper 64 TMEM lanes hold don't-care values that are never read.
#include "capi_internal.h"
#include "common.cuh"
#include "wgrad.h"
#include <type_traits>

namespace sgcn {

enum { WG_SPATIAL = 0, WG_TEMPORAL = 1, WG_PLAIN = 2 };

constexpr int kWgMaxGroups = 3;
constexpr int kWgGroupThreads = 256;
constexpr int kWgThreads = 32 + kWgMaxGroups * kWgGroupThreads;   // 800
constexpr int kWgSmemBudget = 200 * 1024;

struct WgGeom {
  int G;              // row groups per tile
  int KR;             // tile rows = G * VP  (VP = joints padded to a multiple of 8)
  int blk;            // bytes of one [KR x 32] operand block
  int stage;          // bytes of one operand stage (A blocks then B blocks)
  int nstages;        // operand stages == builder groups
};

__host__ __device__ constexpr int wg_vp(int V) { return ((V + 7) / 8) * 8; }

// largest tile (G in {4,3,2,1}) that leaves room for three stages (two for the widest layers).  mult = 2 in the
// fp32-accurate mode (3xTF32, SgcnWgrad::prec): a stage then holds the TF32 heads of A and B followed by their tails.
__host__ __device__ inline WgGeom wg_geom(int CA, int CB, int V, int mult) {
  WgGeom g;
  const int cands[4] = {4, 3, 2, 1};
  for (int i = 0; i < 4; ++i) {
    g.G = cands[i];
    g.KR = g.G * wg_vp(V);
    g.blk = g.KR * 128;
    g.stage = mult * ((CA + CB) / 32) * g.blk;
    g.nstages = kWgSmemBudget / g.stage;
    if (g.nstages > kWgMaxGroups) g.nstages = kWgMaxGroups;
    if (g.KR <= 128 && g.nstages >= (g.G == 1 ? 1 : 3)) break;
  }
  return g;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts_tf32(uint32_t saddr, float v) { sts32(saddr, tf32_half_ulp(v)); }

// Tile row order: row(g, v) = g * VP + v.  The contraction runs over rows, so any order works as long as A and B
// agree; this one makes the swizzle phase (row & 3) of every row a builder thread writes equal to (warp & 3),
// i.e. all shared-memory offsets are "per-thread constant + compile-time immediate".
__host__ __device__ constexpr int wg_pick_g(int CA, int CB, int V, int mult) {   // same rule as wg_geom, usable as a template argument
  for (int i = 0; i < 4; ++i) {
    const int G = 4 - i;
    const int KR = G * wg_vp(V);
    const int stage = mult * ((CA + CB) / 32) * KR * 128;
    int ns = kWgSmemBudget / stage;
    if (ns > kWgMaxGroups) ns = kWgMaxGroups;
    if (KR <= 128 && ns >= (G == 1 ? 1 : 3)) return G;
  }
  return 1;
}

template <int MODE, int V, int CA, int CB, bool P3>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const SgcnWgrad p, const WgGeom geo, const int rev) {
  constexpr int G = wg_pick_g(CA, CB, V, P3 ? 2 : 1);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // aligned as an OFFSET from the __shared__ array (a pointer rebuilt from an integer becomes generic: LD/ST instead of LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KV = (V + 7) / 8;                           // joint slots per builder warp
  constexpr int VP = KV * 8;
  constexpr int KR = G * VP;
  constexpr int BLK = KR * 128;
  constexpr uint32_t ablocks = CA / 32, bblocks = CB / 32;
  constexpr int mblocks = (CA + 127) / 128;
  const int NG = geo.nstages;
  constexpr uint32_t kTail = (ablocks + bblocks) * BLK;     // P3: byte offset from a head element to its tail
  const uint32_t stage_bytes = (P3 ? 2u : 1u) * kTail;
  // store one operand element: TF32-rounded, or (P3) its TF32 head plus the tail kTail bytes further on
  auto put = [&](uint32_t saddr, float v) {
    if constexpr (P3) {
      float hi, lo;
      split_tf32(v, hi, lo);
      sts32(saddr, hi);
      sts32(saddr + kTail, lo);
    } else {
      sts32(saddr, tf32_half_ulp(v));
    }
  };

  __shared__ uint64_t bar_full[kWgMaxGroups], bar_free[kWgMaxGroups], bar_done;
  __shared__ uint32_t tmem_base_s;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < kWgMaxGroups; ++s) {
      mbar_init(&bar_full[s], kWgGroupThreads);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  constexpr int need_cols = mblocks * CB;
  constexpr uint32_t tmem_cols = need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u));
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  {  // padding rows (joints V..VP-1) must be finite zeros; they are never written afterwards
    const int n16 = NG * (int)stage_bytes / 16;
    for (int i = tid; i < n16; i += kWgThreads) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const long long ntiles = (p.groups + G - 1) / G;
  const long long split = blockIdx.x, nsplit = gridDim.x;
  const long long my_tiles = split < ntiles ? (ntiles - split + nsplit - 1) / nsplit : 0;

  if (warp == 0) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, CB, 1, 1);
      for (long long j = 0; j < my_tiles; ++j) {
        const int s = (int)(j % NG);
        mbar_wait(&bar_full[s], (uint32_t)((j / NG) & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(smem) + (uint32_t)s * stage_bytes;
        const uint32_t sb = sa + ablocks * (uint32_t)BLK;
        for (int mb = 0; mb < mblocks; ++mb)
#pragma unroll
          for (int ks = 0; ks < KR / 8; ++ks) {
            const uint64_t da = umma_desc(sa + (uint32_t)mb * 4u * BLK + ks * 1024, BLK, 512, 1);
            const uint64_t db = umma_desc(sb + ks * 1024, BLK, 512, 1);
            umma_tf32(tmem_base + (uint32_t)(mb * CB), da, db, idesc, (j | ks) ? 1u : 0u);
            if constexpr (P3) {                               // + A_tail^T B_head + A_head^T B_tail
              umma_tf32(tmem_base + (uint32_t)(mb * CB), umma_desc(sa + kTail + (uint32_t)mb * 4u * BLK + ks * 1024, BLK, 512, 1),
                        db, idesc, 1u);
              umma_tf32(tmem_base + (uint32_t)(mb * CB), da, umma_desc(sb + kTail + ks * 1024, BLK, 512, 1), idesc, 1u);
            }
          }
        tc_commit(&bar_free[s]);
      }
      tc_commit(&bar_done);
    }
    __syncwarp();
  } else if ((tid - 32) / kWgGroupThreads < NG) {
    // ================================================================ builders
    const int gt = (tid - 32) % kWgGroupThreads;
    const int grp = (tid - 32) / kWgGroupThreads;          // builder group == operand stage
    const int w = gt >> 5;                                  // warp inside the group, 0..7
    uint8_t* sA = smem + (size_t)grp * stage_bytes;
    const uint32_t sA32 = smem_u32(sA), sB32 = sA32 + ablocks * (uint32_t)BLK;    // 32-bit shared-window addresses
    // offset of (row = g*VP + 8*slot + w, channel = lane) inside a block: thread constant + (g*VP + 8*slot)*128
    const uint32_t toff = (uint32_t)w * 128u + ((((uint32_t)lane >> 3) ^ ((uint32_t)w & 3u)) << 5) + (((uint32_t)lane & 7u) << 2);
    const int T = p.T;

    for (long long j = grp; j < my_tiles; j += NG) {
      const long long tile = rev ? ntiles - 1 - (split + j * nsplit) : split + j * nsplit;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const long long use = j / NG;
      // the MMAs of the previous use of this stage must be done before it is overwritten.  SPATIAL stages everything
      // through registers, so it waits only in front of its first shared-memory store -- the first batch of global
      // loads is already in flight while the tensor core drains the stage.
      // TEMPORAL does the same with the register-staged B operand: its first batch of loads is issued, then the wait
      // and the cp.async copy of A (`stage_ready`).
      if (MODE == WG_PLAIN && use > 0) mbar_wait_relaxed(&bar_free[grp], (uint32_t)((use - 1) & 1));
      bool a_pending = MODE == WG_TEMPORAL;

      auto issue_a = [&]() {
        // ---- A = rows as they are: 16-byte cp.async pieces straight into the swizzled blocks (the tensor core reads
        //      fp32 bit patterns as TF32).  PLAIN: row group g of the tile comes from group (g0+g)*a_gs (strided 1x1 conv)
        {
          constexpr int ppr = CA >> 2;                       // 16-byte pieces per row (16, 32 or 64)
          const int k = gt & (ppr - 1);                      // this thread's piece inside a row
          const uint32_t kk = (uint32_t)k & 7u;
          const uint32_t poff = ((uint32_t)k >> 3) * BLK + ((kk & 1u) << 4);
          const long long ags = MODE == WG_PLAIN ? (p.a_gs > 0 ? p.a_gs : 1) : 1;
          const float* src = p.a_src + (size_t)g0 * ags * V * CA + (size_t)k * 4;
          constexpr int rstep = kWgGroupThreads / ppr;
          for (int q = gt / ppr; q < ng * V; q += rstep) {   // q = g*V + v
            const int g = q / V, v = q - g * V;
            const uint32_t r = (uint32_t)(g * VP + v);
            cp_async16(sA + poff + r * 128u + ((((kk >> 1) ^ (r & 3u))) << 5), src + ((size_t)g * ags * V + v) * CA);
          }
          cp_async_commit();
        }
      };
      // P3: once this thread's cp.async copies have landed it splits ITS OWN 16-byte pieces of a raw-copied operand
      // (blocks at `base`, `C` channels wide, group stride `gs`) into TF32 heads (in place) and tails
      auto split_own = [&](uint8_t* base, int C, long long gs) {
        (void)gs;
        const int ppr = C >> 2;
        const int k = gt & (ppr - 1);
        const uint32_t kk = (uint32_t)k & 7u;
        const uint32_t poff = ((uint32_t)k >> 3) * BLK + ((kk & 1u) << 4);
        const int rstep = kWgGroupThreads / ppr;
        for (int q = gt / ppr; q < ng * V; q += rstep) {
          const int g = q / V, v = q - g * V;
          const uint32_t r = (uint32_t)(g * VP + v);
          float4* ptr = (float4*)(base + poff + r * 128u + ((((kk >> 1) ^ (r & 3u))) << 5));
          const float4 x = *ptr;
          float4 hi, lo;
          split_tf32(x.x, hi.x, lo.x), split_tf32(x.y, hi.y, lo.y), split_tf32(x.z, hi.z, lo.z), split_tf32(x.w, hi.w, lo.w);
          *ptr = hi;
          *(float4*)((uint8_t*)ptr + kTail) = lo;
        }
      };
      auto stage_ready = [&]() {                             // TEMPORAL: called in front of the first store of a tile
        if (a_pending) {
          if (use > 0) mbar_wait_relaxed(&bar_free[grp], (uint32_t)((use - 1) & 1));
          issue_a();
          a_pending = false;
        }
      };
      if (MODE == WG_PLAIN) issue_a();
      if (MODE == WG_PLAIN) {
        // ---- B = rows as they are
        constexpr int ppr = CB >> 2;
        const int k = gt & (ppr - 1);
        const uint32_t kk = (uint32_t)k & 7u;
        const uint32_t poff = ((uint32_t)k >> 3) * BLK + ((kk & 1u) << 4);
        const long long bgs = p.b_gs > 0 ? p.b_gs : 1;          // like a_gs: group g of the tile is group (g0+g)*b_gs
        const float* src = p.b_src + (size_t)g0 * bgs * V * CB + (size_t)k * 4;
        constexpr int rstep = kWgGroupThreads / ppr;
        uint8_t* sBp = sA + (size_t)ablocks * BLK;
        for (int q = gt / ppr; q < ng * V; q += rstep) {
          const int g = q / V, v = q - g * V;
          const uint32_t r = (uint32_t)(g * VP + v);
          cp_async16(sBp + poff + r * 128u + ((((kk >> 1) ^ (r & 3u))) << 5), src + ((size_t)g * bgs * V + v) * CB);
        }
        cp_async_commit();
        cp_async_wait_all();
        if constexpr (P3) {
          split_own(sA, CA, 1);
          split_own(sBp, CB, 1);
        }
      } else if (MODE == WG_TEMPORAL) {
        // ---- B = p[(g,v), c] = (1-f) U(t+y1) + f U(t+y1+1),  U = sa*h + sb inside the sample, 0 outside
        const int t0 = (int)(g0 % T);                        // frame of the tile's first group
        const bool straddle = t0 + ng > T;                   // the tile crosses into the next sample (warp uniform)
        const int rel_lo = (int)(-g0 < -(1 << 20) ? -(1 << 20) : -g0);                     // clamp window of the
        const long long room = p.groups - 1 - g0;                                          // tap frame groups,
        const int rel_hi = (int)(room > (1 << 20) ? (1 << 20) : room);                     // relative to g0
#ifndef SGCN_WG_NBT3
#define SGCN_WG_NBT3 1
#endif
        constexpr int NBT = G >= 3 ? SGCN_WG_NBT3 : 2;       // 32-channel blocks per batch (loads in flight)
        constexpr int kDeep = 8;                             // |floor(ypos)| < kDeep and >= kDeep frames from both sample ends
        const bool deep = ng == G && t0 >= kDeep && t0 + G + kDeep <= T;   // (warp uniform) no clamps, no zero padding
        if (deep) {
          // interior tile: the BN affine commutes with the interpolation,  p = (sa*(1-f)) h[t+y1] + (sa*f) h[t+y1+1] + sb
          for (uint32_t blk0 = 0; blk0 < bblocks; blk0 += NBT) {
            float L[NBT][KV][G + 1], a0[NBT], a1[NBT], sb[NBT];
            bool far[NBT];
#pragma unroll
            for (int b = 0; b < NBT; ++b) {
              const int c = (int)(blk0 + b) * 32 + lane;
              const float ypos = __ldg(p.b_tab2 + c), sa = __ldg(p.b_tab0 + c);
              sb[b] = __ldg(p.b_tab1 + c);
              const float fl = floorf(ypos);
              int y1 = (int)fl;
              far[b] = y1 < -kDeep || y1 + 1 > kDeep;        // taps may leave the sample: handled below
              y1 = min(max(y1, -kDeep), kDeep - 1);
              a1[b] = sa * (ypos - fl);
              a0[b] = sa - a1[b];
              const float* src = p.b_src + ((size_t)(g0 + y1) * V + w) * CB + c;
#pragma unroll
              for (int sl = 0; sl < KV; ++sl) {
                const int dv = (min(w + 8 * sl, V - 1) - w) * CB;
#pragma unroll
                for (int k = 0; k <= G; ++k) L[b][sl][k] = __ldg(src + k * (V * CB) + dv);
              }
            }
            stage_ready();
#pragma unroll
            for (int b = 0; b < NBT; ++b) {
              const uint32_t dst = sB32 + (blk0 + b) * (uint32_t)BLK + toff;
              if (far[b]) {                                  // |ypos| beyond the fast window (rare): exact taps with padding
                const int c = (int)(blk0 + b) * 32 + lane;
                const float ypos = __ldg(p.b_tab2 + c), sa = __ldg(p.b_tab0 + c);
                const float fl = floorf(ypos), f = ypos - fl;
                const int y1 = (int)fl;
                for (int sl = 0; sl < KV; ++sl)
                  if (w + 8 * sl < V)
                    for (int g = 0; g < G; ++g) {
                      const int ta = t0 + g + y1, tb = ta + 1;
                      const float* row = p.b_src + ((size_t)(g0 - t0) * V + (w + 8 * sl)) * CB + c;   // frame 0 of the sample
                      const float u0 = (unsigned)ta < (unsigned)T ? fmaf(sa, __ldg(row + (size_t)ta * V * CB), sb[b]) : 0.f;
                      const float u1 = (unsigned)tb < (unsigned)T ? fmaf(sa, __ldg(row + (size_t)tb * V * CB), sb[b]) : 0.f;
                      put(dst + (g * VP + 8 * sl) * 128, fmaf(f, u1, (1.f - f) * u0));
                    }
                continue;
              }
#pragma unroll
              for (int sl = 0; sl < KV; ++sl)
                if (w + 8 * sl < V) {
#pragma unroll
                  for (int g = 0; g < G; ++g)
                    put(dst + (g * VP + 8 * sl) * 128, fmaf(a0[b], L[b][sl][g], fmaf(a1[b], L[b][sl][g + 1], sb[b])));
                }
            }
          }
        } else
        for (uint32_t blk0 = 0; blk0 < bblocks; blk0 += NBT) {
          float L[NBT][KV][G + 1], sa[NBT], sb[NBT], f[NBT];
          int y1[NBT];
#pragma unroll
          for (int b = 0; b < NBT; ++b) {
            const int c = (int)(blk0 + b) * 32 + lane;
            const float ypos = __ldg(p.b_tab2 + c);
            sa[b] = __ldg(p.b_tab0 + c), sb[b] = __ldg(p.b_tab1 + c);
            const float fl = floorf(ypos);
            y1[b] = (int)fl;
            f[b] = ypos - fl;
            const float* src = p.b_src + ((size_t)g0 * V + w) * CB + c;     // (group g0, joint w, channel c)
            const int gstride = V * CB;
#pragma unroll
            for (int sl = 0; sl < KV; ++sl) {
              const int v = min(w + 8 * sl, V - 1);
#pragma unroll
              for (int k = 0; k <= G; ++k) {
                const int rel = min(max(y1[b] + k, rel_lo), rel_hi);
                L[b][sl][k] = __ldg(src + (long long)rel * gstride + (v - w) * CB);
              }
            }
          }
          stage_ready();
#pragma unroll
          for (int b = 0; b < NBT; ++b) {
            const uint32_t dst = sB32 + (blk0 + b) * (uint32_t)BLK + toff;
            const float f0 = 1.f - f[b];
            if (!straddle) {
              bool ok[G + 1];
#pragma unroll
              for (int k = 0; k <= G; ++k) ok[k] = (unsigned)(t0 + y1[b] + k) < (unsigned)T;
#pragma unroll
              for (int sl = 0; sl < KV; ++sl)
                if (w + 8 * sl < V) {
                  float U[G + 1];
#pragma unroll
                  for (int k = 0; k <= G; ++k) U[k] = ok[k] ? fmaf(sa[b], L[b][sl][k], sb[b]) : 0.f;
#pragma unroll
                  for (int g = 0; g < G; ++g)
                    if (g < ng) put(dst + (g * VP + 8 * sl) * 128, fmaf(f[b], U[g + 1], f0 * U[g]));
                }
            } else {
#pragma unroll
              for (int sl = 0; sl < KV; ++sl)
                if (w + 8 * sl < V) {
#pragma unroll
                  for (int g = 0; g < G; ++g)
                    if (g < ng) {
                      int t = t0 + g;
                      if (t >= T) t -= T;
                      const float u0 = ((unsigned)(t + y1[b]) < (unsigned)T) ? fmaf(sa[b], L[b][sl][g], sb[b]) : 0.f;
                      const float u1 = ((unsigned)(t + y1[b] + 1) < (unsigned)T) ? fmaf(sa[b], L[b][sl][g + 1], sb[b]) : 0.f;
                      put(dst + (g * VP + 8 * sl) * 128, fmaf(f[b], u1, f0 * u0));
                    }
                }
            }
          }
        }
        cp_async_wait_all();
        if constexpr (P3) split_own(sA, CA, 1);
      } else {
        auto build_spatial = [&](auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;   // full tile: compile-time offsets, no predicates
        // Scatter on write: a thread loads SOURCE joint sv = w + 8*slot of channel `lane` (every global load of a warp
        // is one contiguous 128-byte row segment, and the per-(joint, channel) tables are indexed naturally) and
        // writes the operand row the joint shift sends it to.  Consecutive channels land in consecutive rows, which
        // the BASE32B swizzle spreads over all 32 banks.
        // ---- A = xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]          (model/shift_gcn.py:127-129)
        //      loaded as x[g, sv, c] -> row u = (sv - c) mod V;  a_tab0[sv, c] = maskmul[(sv - c) mod V, c]
#ifndef SGCN_WG_NB4
#define SGCN_WG_NB4 1
#endif
#ifndef SGCN_WG_NBB4
#define SGCN_WG_NBB4 1
#endif
        constexpr int NB0 = G >= 4 ? SGCN_WG_NB4 : (G >= 2 ? 2 : 4);  // 32-channel blocks per batch: ~16 independent loads in flight
        constexpr int NB = NB0 < (int)ablocks ? NB0 : (int)ablocks;   // (never more blocks than A has: CA = 64 with G = 1)
        constexpr int NBB = G >= 4 ? SGCN_WG_NBB4 : 2;
        for (uint32_t blk0 = 0; blk0 < ablocks; blk0 += NB) {
          float val[NB][KV][G], mm[NB][KV];
          uint32_t off[NB][KV];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const int c = (int)(blk0 + b) * 32 + lane;
            const int cm = c % V;
            const float* src = p.a_src + (size_t)g0 * V * CA + c;
            const int gstride = V * CA;
#pragma unroll
            for (int sl = 0; sl < KV; ++sl) {
              const int sv = min(w + 8 * sl, V - 1);
              int u = sv - cm;
              if (u < 0) u += V;
              off[b][sl] = (uint32_t)u * 128u + ((((uint32_t)lane >> 3) ^ ((uint32_t)u & 3u)) << 5);
              mm[b][sl] = __ldg(p.a_tab0 + sv * CA + c);
#pragma unroll
              for (int g = 0; g < G; ++g) val[b][sl][g] = __ldg(src + (kFull ? g : min(g, ng - 1)) * gstride + sv * CA);
            }
          }
          if (blk0 == 0 && use > 0) mbar_wait_relaxed(&bar_free[grp], (uint32_t)((use - 1) & 1));
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint32_t dst = sA32 + (blk0 + b) * (uint32_t)BLK + (((uint32_t)lane & 7u) << 2);
#pragma unroll
            for (int sl = 0; sl < KV; ++sl)
              if (w + 8 * sl < V) {
#pragma unroll
                for (int g = 0; g < G; ++g)
                  if (kFull || g < ng) put(dst + off[b][sl] + g * VP * 128, val[b][sl][g] * mm[b][sl]);
              }
          }
        }
        // ---- B = dy[(g,u), d] = dz[g, (u+d) % V, d],  dz = alpha*gh + beta*z + gamma   (BN1d backward folded
        //      into three per-(v,d) tables; inverse of the shift_out gather, model/shift_gcn.py:135-137)
        //      loaded as (gh, z)[g, sv, d] -> row u = (sv - d) mod V
        for (uint32_t blk0 = 0; blk0 < bblocks; blk0 += NBB) {
          constexpr int KH = (KV + 1) / 2;                   // two half passes keep the batch at <= 2*NB*G*KH loads
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float gv[NBB][KH][G], zv[NBB][KH][G], al[NBB][KH], be[NBB][KH], ga[NBB][KH];
            uint32_t off[NBB][KH];
#pragma unroll
            for (int b = 0; b < NBB; ++b) {
              const int d = (int)(blk0 + b) * 32 + lane;
              const int dm = d % V;
              const size_t o = (size_t)g0 * V * CB + d;
              const int gstride = V * CB;
#pragma unroll
              for (int i = 0; i < KH; ++i) {
                const int sl = half * KH + i;
                const int sv = min(w + 8 * sl, V - 1);
                int u = sv - dm;
                if (u < 0) u += V;
                off[b][i] = (uint32_t)u * 128u + ((((uint32_t)lane >> 3) ^ ((uint32_t)u & 3u)) << 5);
                al[b][i] = __ldg(p.b_tab0 + sv * CB + d);
                be[b][i] = __ldg(p.b_tab1 + sv * CB + d);
                ga[b][i] = __ldg(p.b_tab2 + sv * CB + d);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  const size_t og = o + (size_t)((kFull ? g : min(g, ng - 1)) * gstride + sv * CB);
                  gv[b][i][g] = __ldg(p.b_src + og);
                  zv[b][i][g] = __ldg(p.b_src2 + og);
                }
              }
            }
#pragma unroll
            for (int b = 0; b < NBB; ++b) {
              const uint32_t dst = sB32 + (blk0 + b) * (uint32_t)BLK + (((uint32_t)lane & 7u) << 2);
#pragma unroll
              for (int i = 0; i < KH; ++i) {
                const int sl = half * KH + i;
                if (sl < KV && w + 8 * sl < V) {
#pragma unroll
                  for (int g = 0; g < G; ++g)
                    if (kFull || g < ng)
                      put(dst + off[b][i] + g * VP * 128, fmaf(al[b][i], gv[b][i][g], fmaf(be[b][i], zv[b][i][g], ga[b][i])));
                }
              }
            }
          }
        }
              };
        if (ng == G) build_spatial(std::true_type{});
        else build_spatial(std::false_type{});
      }
      if (ng < G) {   // partial last tile: rows of missing groups may hold an earlier tile
        const int nblk = (int)((P3 ? 2 : 1) * (ablocks + bblocks));
        for (int r = ng * VP + w; r < KR; r += 8)
          for (int blk = 0; blk < nblk; ++blk) *(float*)(sA + (size_t)blk * BLK + r * 128 + lane * 4) = 0.f;
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[grp]);
    }

    // ================================================================ flush (warps 1..4): TMEM lane = A channel.
    // Transposed through shared memory so that the atomics of a warp hit 32 consecutive addresses.
    if (warp >= 1 && warp <= 4 && my_tiles > 0) {
      mbar_wait_relaxed(&bar_done, 0);          // every MMA has completed: the operand stages are free to be reused
      tc_fence_after();
      const int q = warp & 3;                               // TMEM lane quarter this warp may read
      float* tr = (float*)smem + q * (32 * 33);
      for (int mb = 0; mb < mblocks; ++mb) {
        const int ch0 = mb * 128 + q * 32;
        if (ch0 >= CA) continue;                            // (warp-uniform)
        for (int c0 = 0; c0 < CB; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * CB + c0), v);
#pragma unroll
          for (int k = 0; k < 32; ++k) tr[lane * 33 + k] = v[k];
          __syncwarp();
#pragma unroll 4
          for (int r = 0; r < 32; ++r) atomicAdd(p.dw + (size_t)(ch0 + r) * CB + c0 + lane, tr[r * 33 + lane]);
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

template <int MODE, int V, int CA, int CB, bool P3>
static int launch_wgrad_p(const SgcnWgrad& p, cudaStream_t s) {
  const WgGeom geo = wg_geom(CA, CB, V, P3 ? 2 : 1);
  constexpr int G = wg_pick_g(CA, CB, V, P3 ? 2 : 1);
  if (geo.nstages < 1 || geo.G != G) return set_error("sgcn_wgrad: operand stage does not fit in shared memory");
  const size_t smem = 1024 + (size_t)geo.nstages * geo.stage + 64;
  auto kern = wgrad_kernel<MODE, V, CA, CB, P3>;
  static std::atomic<unsigned long long> configured{0};           // one bit per device; smem is fixed per instantiation
  if (needs_configure(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("wgrad smem attribute", e);
    mark_configured(configured);
  }
  const long long ntiles = (p.groups + G - 1) / G;
  if (ntiles == 0) return 0;
  long long grid = tile_ctas();
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kWgThreads, smem, s>>>(p, geo, next_direction());
  return check_launch("wgrad_kernel");
}

template <int MODE, int V, int CA, int CB>
static int launch_wgrad_cc(const SgcnWgrad& p, cudaStream_t s) {
  if (p.prec == SGCN_PREC_FP32) return launch_wgrad_p<MODE, V, CA, CB, true>(p, s);
  if (p.prec != SGCN_PREC_TF32) return set_error("sgcn_wgrad: unknown precision");
  return launch_wgrad_p<MODE, V, CA, CB, false>(p, s);
}

template <int MODE, int V>
static int launch_wgrad_v(const SgcnWgrad& p, cudaStream_t s) {
  switch (p.CA * 1000 + p.CB) {
    case 64064: return launch_wgrad_cc<MODE, V, 64, 64>(p, s);
    case 64128: return launch_wgrad_cc<MODE, V, 64, 128>(p, s);
    case 128128: return launch_wgrad_cc<MODE, V, 128, 128>(p, s);
    case 128256: return launch_wgrad_cc<MODE, V, 128, 256>(p, s);
    case 256256: return launch_wgrad_cc<MODE, V, 256, 256>(p, s);
    default: return set_error("sgcn_wgrad: unsupported (CA, CB) channel pair");
  }
}

template <int MODE>
static int launch_wgrad(const SgcnWgrad& p, cudaStream_t s) {
  if (p.V == 25) return launch_wgrad_v<MODE, 25>(p, s);
  if (p.V == 33) return launch_wgrad_v<MODE, 33>(p, s);
  return set_error("sgcn_wgrad: num_point must be 25 (NTU) or 33 (MediaPipe)");
}

}  // namespace sgcn

extern "C" int sgcn_wgrad(const SgcnWgrad* pp, int mode, void* stream) {
  using namespace sgcn;
  if (!pp) return set_error("sgcn_wgrad: null params");
  const SgcnWgrad& p = *pp;
  if ((p.CA != 64 && p.CA != 128 && p.CA != 256) || (p.CB != 64 && p.CB != 128 && p.CB != 256))
    return set_error("sgcn_wgrad: channel counts must be 64, 128 or 256");
  if (!p.a_src || !p.b_src || !p.dw) return set_error("sgcn_wgrad: null tensor");
  if (p.groups < 0) return set_error("sgcn_wgrad: negative group count");
  if (mode == WG_SPATIAL) {
    if (!p.a_tab0 || !p.b_src2 || !p.b_tab0 || !p.b_tab1 || !p.b_tab2) return set_error("sgcn_wgrad(spatial): null table");
    return launch_wgrad<WG_SPATIAL>(p, (cudaStream_t)stream);
  }
  if (mode == WG_TEMPORAL) {
    if (!p.b_tab0 || !p.b_tab1 || !p.b_tab2 || p.T < 1) return set_error("sgcn_wgrad(temporal): null table / bad T");
    return launch_wgrad<WG_TEMPORAL>(p, (cudaStream_t)stream);
  }
  if (mode == WG_PLAIN) return launch_wgrad<WG_PLAIN>(p, (cudaStream_t)stream);
  return set_error("sgcn_wgrad: unknown mode");
}
