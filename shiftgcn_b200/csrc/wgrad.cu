// wgrad.cu -- weight-gradient contractions  dW[a, b] = sum_rows  A[row, a] * B[row, b]
// on tcgen05 tensor cores, operands built on the fly from the activations.
//
//   spatial  (autograd of model/shift_gcn.py:131):  A = xm  (gather + mask of x),  B = dy (BN1d-bwd + inverse rotation)
//            -> Linear_weight.grad [C, D]
//   temporal (autograd of model/shift_gcn.py:69):   A = dpre (grad wrt the conv output, pre-ReLU), B = p (shifted BN(h))
//            -> temporal_linear.weight.grad [Cout, Cin]
//
// Warp-specialised, persistent: one CTA per SM owns an (A-channel block, B-channel block) pair (<= 128 x 128) and a
// strided subset of the row tiles; its partial dW stays in TMEM for the whole kernel and is flushed once.
//   warp 0            issues tcgen05.mma (one lane), owns the TMEM allocation
//   3 builder groups  of 8 warps; group g builds tiles g, g+3, ... of this CTA into operand stage g:
//                     global loads (lane <-> channel, 128-byte coalesced, ~30 independent loads in flight per thread
//                     -- the register file is the staging buffer) -> prologue math -> TF32 -> swizzled smem
//   full[s] / free[s] mbarriers per stage: 256 builder arrivals / one tcgen05.commit
// While one group waits for its loads the other two are computing or their stages are being consumed, so HBM
// always has ~2 tiles of requests in flight per SM.
//
// Both operands are MN-major (the row index is the contraction dimension, channels contiguous) in the
// SWIZZLE_128B_BASE32B layout -- the only one tcgen05 accepts for MN-major 32-bit operands: blocks of
// [KR rows x 32 channels], 128 B per row, 32-byte chunks XOR-ed with (row % 4).  The M extent of the MMA is always
// 128: when the A block has only 64 channels the upper 64 TMEM lanes hold don't-care values that are never read.
#include "capi_internal.h"
#include "common.cuh"
#include "wgrad.h"

namespace sgcn {

enum { WG_SPATIAL = 0, WG_TEMPORAL = 1 };

constexpr int kWgGroups = 3;
constexpr int kWgGroupThreads = 256;
constexpr int kWgThreads = 32 + kWgGroups * kWgGroupThreads;   // 800
constexpr int kWgGMax = 5;                                     // groups per tile (V >= 25)

struct WgGeom {
  int MC, NC;         // channels of the A / B block of one CTA
  int G;              // row groups per tile
  int KR;             // tile rows rounded up to the MMA K granularity (8)
  int blk;            // bytes of one [KR x 32] operand block
  int stage;          // bytes of one operand stage (A blocks then B blocks)
};

__host__ __device__ inline WgGeom wg_geom(int CA, int CB, int V) {
  WgGeom g;
  g.MC = CA < 128 ? CA : 128;
  g.NC = CB < 128 ? CB : 128;
  g.G = (g.MC + g.NC <= 128 ? 128 : 64) / V;
  if (g.G < 1) g.G = 1;
  if (g.G > kWgGMax) g.G = kWgGMax;
  g.KR = (g.G * V + 7) & ~7;
  g.blk = g.KR * 128;
  g.stage = ((g.MC + g.NC) / 32) * g.blk;
  return g;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void sts_tf32(uint8_t* base, uint32_t off, float v) { *(float*)(base + off) = to_tf32(v); }

template <int MODE>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const SgcnWgrad p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V, CA = p.CA, CB = p.CB;
  const WgGeom geo = wg_geom(CA, CB, V);
  const int G = geo.G, MC = geo.MC, NC = geo.NC;
  const int nb = CB / NC, npairs = (CA / MC) * nb;
  const int pair = blockIdx.x % npairs, split = blockIdx.x / npairs, nsplit = gridDim.x / npairs;
  const int a0 = (pair / nb) * MC, b0 = (pair % nb) * NC;
  const uint32_t ablocks = MC / 32;

  __shared__ uint64_t bar_full[kWgGroups], bar_free[kWgGroups], bar_done;
  __shared__ uint32_t tmem_base_s;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < kWgGroups; ++s) {
      mbar_init(&bar_full[s], kWgGroupThreads);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  const uint32_t tmem_cols = NC <= 64 ? 64u : 128u;
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  {  // padding rows (and the never-written upper A blocks) must be finite zeros
    const int n16 = kWgGroups * geo.stage / 16;
    for (int i = tid; i < n16; i += kWgThreads) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const long long ntiles = (p.groups + G - 1) / G;
  const long long my_tiles = split < ntiles ? (ntiles - split + nsplit - 1) / nsplit : 0;

  if (warp == 0) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, NC, 1, 1);
      const int ksteps = geo.KR / 8;
      for (long long j = 0; j < my_tiles; ++j) {
        const int s = (int)(j % kWgGroups);
        mbar_wait(&bar_full[s], (uint32_t)((j / kWgGroups) & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(smem) + (uint32_t)s * (uint32_t)geo.stage;
        const uint32_t sb = sa + ablocks * (uint32_t)geo.blk;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_tf32(tmem_base, umma_desc(sa + ks * 1024, (uint32_t)geo.blk, 512, 1),
                    umma_desc(sb + ks * 1024, (uint32_t)geo.blk, 512, 1), idesc, (j | ks) ? 1u : 0u);
        tc_commit(&bar_free[s]);
      }
      tc_commit(&bar_done);
    }
    __syncwarp();
  } else {
    // ================================================================ builders
    const int gt = tid - 32;
    const int grp = gt / kWgGroupThreads;                 // builder group == operand stage
    const int w = (gt % kWgGroupThreads) >> 5;            // warp inside the group, 0..7
    uint8_t* sA = smem + (size_t)grp * geo.stage;
    uint8_t* sB = sA + (size_t)ablocks * geo.blk;
    // MN-major swizzle: offset(row, ch) = row*128 + (((ch>>3) ^ (row&3)) << 5) + ((ch&7) << 2)
    const uint32_t lane_lo = (uint32_t)((lane & 7) << 2), lane_hi = (uint32_t)(lane >> 3);
    const long long last_group = p.groups - 1;

    for (long long j = grp; j < my_tiles; j += kWgGroups) {
      const long long tile = split + j * nsplit;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const int rows_valid = ng * V;
      const long long row0 = g0 * V;
      const long long use = j / kWgGroups;
      if (use > 0) mbar_wait(&bar_free[grp], (uint32_t)((use - 1) & 1));   // the MMAs of the previous use are done

      if (MODE == WG_TEMPORAL) {
        // ---- A = dpre rows as they are: rows r = w + 8 i share the swizzle phase r & 3 == w & 3
        const uint32_t cw = ((lane_hi ^ (uint32_t)(w & 3)) << 5) + lane_lo;
        for (uint32_t blk = 0; blk < ablocks; ++blk) {
          const float* src = p.a_src + (size_t)row0 * CA + a0 + blk * 32 + lane;
          uint8_t* dst = sA + (size_t)blk * geo.blk + cw;
          float val[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = min(w + 8 * i, rows_valid - 1);
            val[i] = __ldg(src + (size_t)r * CA);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = w + 8 * i;
            if (r < rows_valid) sts_tf32(dst, (uint32_t)r * 128u, val[i]);
          }
        }
        // ---- B = p[(g,v), c] = (1-f) U(t+y1) + f U(t+y1+1),  U = sa*h + sb inside the sample, 0 outside
        for (uint32_t blk = 0; blk < (uint32_t)NC / 32; ++blk) {
          const int c = b0 + blk * 32 + lane;
          const float ypos = __ldg(p.b_tab2 + c), sa = __ldg(p.b_tab0 + c), sb = __ldg(p.b_tab1 + c);
          const float fl = floorf(ypos);
          const int y1 = (int)fl;
          const float f = ypos - fl, f0 = 1.f - f;
          uint8_t* dst = sB + (size_t)blk * geo.blk;
          const int t0 = (int)(g0 % p.T);                    // frame of the tile's first group
          for (int v = w; v < V; v += 8) {
            float L[kWgGMax + 1];
#pragma unroll
            for (int k = 0; k <= kWgGMax; ++k) {
              long long gi = g0 + y1 + k;                    // frame group of tap k (validity applied below)
              gi = gi < 0 ? 0 : (gi > last_group ? last_group : gi);
              L[k] = (k <= ng) ? __ldg(p.b_src + ((size_t)gi * V + v) * CB + c) : 0.f;
            }
#pragma unroll
            for (int g = 0; g < kWgGMax; ++g)
              if (g < ng) {
                int t = t0 + g;                              // tiles may straddle a sample boundary
                if (t >= p.T) t -= p.T;
                const float u0 = ((unsigned)(t + y1) < (unsigned)p.T) ? fmaf(sa, L[g], sb) : 0.f;
                const float u1 = ((unsigned)(t + y1 + 1) < (unsigned)p.T) ? fmaf(sa, L[g + 1], sb) : 0.f;
                const uint32_t r = (uint32_t)(g * V + v);
                sts_tf32(dst, r * 128u + ((lane_hi ^ (r & 3)) << 5) + lane_lo, fmaf(f, u1, f0 * u0));
              }
          }
        }
      } else {
        // ---- A = xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]          (model/shift_gcn.py:127-129)
        for (uint32_t blk = 0; blk < ablocks; ++blk) {
          const int c = a0 + blk * 32 + lane;
          const int cm = c % V;
          uint8_t* dst = sA + (size_t)blk * geo.blk;
          for (int u = w; u < V; u += 8) {
            int sv = u + cm;
            if (sv >= V) sv -= V;
            const float mm = __ldg(p.a_tab0 + u * CA + c);
            const float* src = p.a_src + ((size_t)row0 + sv) * CA + c;
            float val[kWgGMax];
#pragma unroll
            for (int g = 0; g < kWgGMax; ++g) val[g] = __ldg(src + (size_t)min(g, ng - 1) * V * CA);
#pragma unroll
            for (int g = 0; g < kWgGMax; ++g)
              if (g < ng) {
                const uint32_t r = (uint32_t)(g * V + u);
                sts_tf32(dst, r * 128u + ((lane_hi ^ (r & 3)) << 5) + lane_lo, val[g] * mm);
              }
          }
        }
        // ---- B = dy[(g,u), d] = dz[g, (u+d) % V, d],  dz = alpha*gh + beta*z + gamma   (BN1d backward folded
        //      into three per-(v,d) tables; inverse of the shift_out gather, model/shift_gcn.py:135-137)
        for (uint32_t blk = 0; blk < (uint32_t)NC / 32; ++blk) {
          const int d = b0 + blk * 32 + lane;
          const int dm = d % V;
          uint8_t* dst = sB + (size_t)blk * geo.blk;
          for (int u = w; u < V; u += 8) {
            int sv = u + dm;
            if (sv >= V) sv -= V;
            const float al = __ldg(p.b_tab0 + sv * CB + d), be = __ldg(p.b_tab1 + sv * CB + d),
                        ga = __ldg(p.b_tab2 + sv * CB + d);
            const size_t o = ((size_t)row0 + sv) * CB + d;
            float gv[kWgGMax], zv[kWgGMax];
#pragma unroll
            for (int g = 0; g < kWgGMax; ++g) {
              const size_t og = o + (size_t)min(g, ng - 1) * V * CB;
              gv[g] = __ldg(p.b_src + og);
              zv[g] = __ldg(p.b_src2 + og);
            }
#pragma unroll
            for (int g = 0; g < kWgGMax; ++g)
              if (g < ng) {
                const uint32_t r = (uint32_t)(g * V + u);
                sts_tf32(dst, r * 128u + ((lane_hi ^ (r & 3)) << 5) + lane_lo, fmaf(al, gv[g], fmaf(be, zv[g], ga)));
              }
          }
        }
      }
      if (rows_valid < G * V) {   // partial last tile: rows of missing groups may hold an earlier tile
        const int nblk = (MC + NC) / 32;
        for (int r = rows_valid + w; r < G * V; r += 8)
          for (int blk = 0; blk < nblk; ++blk) *(float*)(sA + (size_t)blk * geo.blk + r * 128 + lane * 4) = 0.f;
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[grp]);
    }

    // ================================================================ flush: TMEM lane = A channel
    if (warp >= 1 && warp <= 4 && my_tiles > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
      const int q = warp & 3;                               // TMEM lane quarter this warp may read
      const int ch = q * 32 + lane;
      for (int c0 = 0; c0 < NC; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (ch < MC) {
          float* dst = p.dw + (size_t)(a0 + ch) * CB + b0 + c0;
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(dst + k, v[k]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

template <int MODE>
static int launch_wgrad(const SgcnWgrad& p, cudaStream_t s) {
  const WgGeom geo = wg_geom(p.CA, p.CB, p.V);
  const int npairs = (p.CA / geo.MC) * (p.CB / geo.NC);
  const size_t smem = 1024 + (size_t)kWgGroups * geo.stage + 64;
  auto kern = wgrad_kernel<MODE>;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error("wgrad smem attribute", e);
    configured = smem;
  }
  const long long ntiles = (p.groups + geo.G - 1) / geo.G;
  if (ntiles == 0) return 0;
  long long nsplit = num_sms() / npairs;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ntiles) nsplit = ntiles;
  kern<<<(unsigned)(nsplit * npairs), kWgThreads, smem, s>>>(p);
  return check_launch("wgrad_kernel");
}

}  // namespace sgcn

extern "C" int sgcn_wgrad(const SgcnWgrad* pp, int mode, void* stream) {
  using namespace sgcn;
  if (!pp) return set_error("sgcn_wgrad: null params");
  const SgcnWgrad& p = *pp;
  if (p.V < 25 || p.V > 40) return set_error("sgcn_wgrad: num_point must be in [25, 40]");
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
  return set_error("sgcn_wgrad: unknown mode");
}
