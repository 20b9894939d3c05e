// fused_gemm.cuh -- the warp-specialised "row tile x weight" kernel behind the forward contractions of the
// Shift-GCN hot path (the backward-data pass of the spatial unit has its own copy of this skeleton with four input
// streams in the epilogue, spatial_bwd.cu).
//
// Activations live channels-last: a logical (n, C, T, V) tensor is the row-major matrix [(n,t,v), C].  A CTA (one
// per SM, persistent) processes tiles of G whole (n,t) groups (G*V <= 128 rows), so the joint shifts of the spatial
// unit are permutations inside the tile.  Three warp roles, connected by mbarrier pipelines:
//
//   builder warps    each thread owns ONE source slot (joint, 4 channels) of the 64-channel chunk and walks the
//                    tile's groups with 128-bit loads (the next chunk's loads are issued before the current chunk is
//                    processed, so ~2 chunks of requests are always in flight), applies the prologue and scatters
//                    TF32 values into the K-major SWIZZLE_128B operand chunk (2 buffers, full / free barriers)
//   MMA warp         one lane issues tcgen05.mma kind::tf32 (M = 128, N = out channels, fp32 accumulators in TMEM,
//                    2 accumulator buffers) and streams weight chunks with cp.async.bulk when W exceeds 64 KiB
//   epilogue warps   TMEM -> XOR-swizzled smem staging (8 warps), then each thread owns ONE destination slot
//                    (joint, 4 channels): gathers its four values from the staging tile (this is where the output
//                    rotation happens), applies the fused tail and stores 128 bits; cross-tile reductions
//                    (BatchNorm statistics) stay in registers for the whole kernel
//
// Variants (template PRO x EPI), with the reference code each one replaces:
//   PRO_SPATIAL  x[r,(u+c)%V,c] * (tanh(mask)+1)            model/shift_gcn.py:123-129
//   PRO_LERP     temporal shift of BN(h), zero padded       model/shift_gcn.py:66-68, shift_cuda_kernel.cu:12-76
//   PRO_PLAIN    rows as they are                           (backward data contraction of the 1x1 conv)
//   EPI_ROT_RAW    + bias, rotate z[v,d]=y[(v-d)%V,d], store, per-(v,d) batch statistics   :131-137 (training)
//   EPI_ROT_FUSED  + bias, rotate, folded BN, + residual, ReLU                              :131-141 (eval)
//   EPI_LINEAR     + bias, optional ReLU, store                                             :69-70
#pragma once
#include "capi_internal.h"
#include "common.cuh"
#include "rowgemm.h"

namespace sgcn {
namespace fg {

enum { PRO_SPATIAL = 0, PRO_LERP = 1, PRO_PLAIN = 2 };
enum { EPI_ROT_RAW = 0, EPI_ROT_FUSED = 1, EPI_LINEAR = 2 };

constexpr int kEpiWarps = 13, kBldWarps = 13;
constexpr int kEpiThreads = kEpiWarps * 32, kBldThreads = kBldWarps * 32;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = (kEpiWarps + 1 + kBldWarps) * 32;       // 864
constexpr int kChunkBytes = 128 * 64 * 4;                        // one [128 x 64] fp32 operand chunk / staging tile

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg((const float4*)p); }
__device__ __forceinline__ float f4get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// staging tile [128 rows][64 cols] fp32, 16-byte chunks XOR-ed with (row & 15)
__device__ __forceinline__ uint32_t stage_off(uint32_t row, uint32_t c4, uint32_t j) {
  return row * 256u + ((c4 ^ (row & 15u)) << 4) + (j << 2);
}
// K-major SWIZZLE_128B operand chunk (two 32-channel blocks): byte offset of (row, 16-byte chunk c4 of 16)
__device__ __forceinline__ uint32_t op_off(uint32_t row, uint32_t c4) {
  return (c4 >> 3) * (uint32_t)kBlockBytes + (row >> 3) * 1024u + (row & 7u) * 128u + (((c4 & 7u) ^ (row & 7u)) << 4);
}

template <int PRO, int EPI, int V, int K, int N>
struct Cfg {
  static constexpr int G = 128 / V;
  static constexpr int KC = K / 64, NCH = N / 64;
  static constexpr bool kWRes = (K * N * 4) <= 65536;
  static constexpr int kWBytes = kWRes ? K * N * 4 : N * 256;
  static constexpr int kSlots = V * 16;                            // (joint, 4-channel group) slots of a 64-wide chunk
  static constexpr int kEpiRounds = (kSlots + kEpiThreads - 1) / kEpiThreads;
  static constexpr int kBldRounds = (kSlots + kBldThreads - 1) / kBldThreads;
  static constexpr int kWin = 3;                                   // LERP: largest tap spread handled with 128-bit loads
  static constexpr int kLoads = PRO == PRO_LERP ? G + kWin : G;    // float4 registers of one builder work item
  static constexpr bool kPrefetch = PRO != PRO_LERP;               // double-buffer the builder loads in registers
  static constexpr size_t kSmem = 1024 + kWBytes + 3 * kChunkBytes + 64;
};

// ------------------------------------------------------------------------------------------------ builder work item
template <int PRO, int V, int K, int G, int NL>
struct Item {
  float4 v[NL];
  long long g0;
  int slot, kc, ng;
  // LERP only
  int lo, span;
};

template <int PRO, int EPI, int V, int K, int N>
__device__ __forceinline__ void issue_loads(const SgcnRowGemm& p, long long tile, int kc, int slot,
                                            Item<PRO, V, K, Cfg<PRO, EPI, V, K, N>::G, Cfg<PRO, EPI, V, K, N>::kLoads>& it) {
  using C = Cfg<PRO, EPI, V, K, N>;
  constexpr int G = C::G;
  it.g0 = tile * G;
  it.ng = (int)((p.groups - it.g0) < G ? (p.groups - it.g0) : G);
  it.slot = slot;
  it.kc = kc;
  const int sv = min(slot, C::kSlots - 1) >> 4, c4 = slot & 15;
  const int c = kc * 64 + c4 * 4;
  if (PRO == PRO_LERP) {
    const float4 yp = ldg4(p.pro_c + c);
    const int y0 = (int)floorf(yp.x), y1 = (int)floorf(yp.y), y2 = (int)floorf(yp.z), y3 = (int)floorf(yp.w);
    it.lo = min(min(y0, y1), min(y2, y3));
    it.span = max(max(y0, y1), max(y2, y3)) + 1 - it.lo;           // taps lo .. lo+span
    const long long last = p.groups - 1;
#pragma unroll
    for (int k = 0; k < C::kLoads; ++k) {
      long long gi = it.g0 + it.lo + k;
      gi = gi < 0 ? 0 : (gi > last ? last : gi);
      it.v[k] = ldg4(p.in0 + ((size_t)gi * V + sv) * K + c);
    }
  } else {
    const size_t o = ((size_t)it.g0 * V + sv) * K + c;
#pragma unroll
    for (int g = 0; g < G; ++g) it.v[g] = ldg4(p.in0 + o + (size_t)min(g, it.ng - 1) * V * K);
  }
}

template <int PRO, int EPI, int V, int K, int N>
__device__ __forceinline__ void build_item(const SgcnRowGemm& p, uint8_t* op,
                                           const Item<PRO, V, K, Cfg<PRO, EPI, V, K, N>::G, Cfg<PRO, EPI, V, K, N>::kLoads>& it) {
  using C = Cfg<PRO, EPI, V, K, N>;
  constexpr int G = C::G;
  if (it.slot >= C::kSlots) return;
  const int sv = it.slot >> 4, c4 = it.slot & 15;
  const int c = it.kc * 64 + c4 * 4;
  if (PRO == PRO_PLAIN) {
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (g < it.ng) {
        const float4 x = it.v[g];
        *(float4*)(op + op_off((uint32_t)(g * V + sv), (uint32_t)c4)) = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
      }
  } else if (PRO == PRO_SPATIAL) {
    // xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]: the loaded source joint sv feeds u_j = (sv - c - j) mod V
    int u[4];
    u[0] = sv - c % V;
    if (u[0] < 0) u[0] += V;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      u[j] = u[j - 1] - 1;
      if (u[j] < 0) u[j] += V;
    }
    float mm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mm[j] = __ldg(p.pro_a + u[j] * K + c + j);
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (g < it.ng) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *(float*)(op + op_off((uint32_t)(g * V + u[j]), (uint32_t)c4) + j * 4) = to_tf32(f4get(it.v[g], j) * mm[j]);
      }
  } else {  // PRO_LERP
    // p[(g,v), c] = (1-f) U(t+y1) + f U(t+y1+1),  U = sa*h + sb inside the sample, 0 outside   (K1 with xpos = 0)
    const float4 yp = ldg4(p.pro_c + c), sa = ldg4(p.pro_a + c), sb = ldg4(p.pro_b + c);
    const int T = p.T;
    const int t0 = (int)(it.g0 % T);
    const long long last = p.groups - 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ypos = f4get(yp, j), a = f4get(sa, j), b = f4get(sb, j);
      const float fl = floorf(ypos);
      const int y1 = (int)fl;
      const float f = ypos - fl, f0 = 1.f - f;
      const int idx = y1 - it.lo;                                  // 0 .. span-1
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (g < it.ng) {
          float h0, h1;
          if (it.span <= C::kWin) {                                // both taps are inside the 128-bit window
            const float w0 = f4get(it.v[g], j), w1 = f4get(it.v[g + 1], j), w2 = f4get(it.v[g + 2], j), w3 = f4get(it.v[g + 3], j);
            h0 = idx == 0 ? w0 : (idx == 1 ? w1 : w2);
            h1 = idx == 0 ? w1 : (idx == 1 ? w2 : w3);
          } else {                                                 // widely spread shift positions: scalar taps
            long long ga = it.g0 + g + y1, gb = ga + 1;
            ga = ga < 0 ? 0 : (ga > last ? last : ga);
            gb = gb < 0 ? 0 : (gb > last ? last : gb);
            h0 = __ldg(p.in0 + ((size_t)ga * V + sv) * K + c + j);
            h1 = __ldg(p.in0 + ((size_t)gb * V + sv) * K + c + j);
          }
          int t = t0 + g;                                          // frame of this group (tiles may straddle samples)
          if (t >= T) t %= T;
          const float u0 = ((unsigned)(t + y1) < (unsigned)T) ? fmaf(a, h0, b) : 0.f;
          const float u1 = ((unsigned)(t + y1 + 1) < (unsigned)T) ? fmaf(a, h1, b) : 0.f;
          *(float*)(op + op_off((uint32_t)(g * V + sv), (uint32_t)c4) + j * 4) = to_tf32(fmaf(f, u1, f0 * u0));
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int PRO, int EPI, int V, int K, int N>
__global__ void __launch_bounds__(kThreads, 1) fused_gemm_kernel(const SgcnRowGemm p) {
  using C = Cfg<PRO, EPI, V, K, N>;
  constexpr int G = C::G, KC = C::KC, NCH = C::NCH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sOp = sW + C::kWBytes;                                  // 2 operand chunks
  uint8_t* sSt = sOp + 2 * kChunkBytes;                            // epilogue staging
  __shared__ uint64_t op_full[2], op_free[2], acc_full[2], acc_free[2], w_full, w_free;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&op_full[i], kBldThreads);
      mbar_init(&op_free[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 8 * 32);
    }
    mbar_init(&w_full, 1);
    mbar_init(&w_free, 1);
    fence_mbar_init();
  }
  constexpr uint32_t tmem_cols = 2 * N <= 128 ? 128u : (2 * N <= 256 ? 256u : 512u);
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, tmem_cols);
  if (C::kWRes) {
    for (int i = tid; i < C::kWBytes / 16; i += kThreads) cp_async16(sW + (size_t)i * 16, (const uint8_t*)p.wimg + (size_t)i * 16);
    cp_async_commit();
    cp_async_wait_all();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const long long ntiles = (p.groups + G - 1) / G;
  const long long my_tiles = (long long)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kMmaWarp) {
    // ================================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
      const long long total_chunks = my_tiles * KC;
      if (!C::kWRes && total_chunks > 0) {
        mbar_expect_tx(&w_full, C::kWBytes);
        bulk_load(sW, p.wimg, C::kWBytes, &w_full);
      }
      long long q = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const int buf = (int)(ti & 1);
        if (ti >= 2) mbar_wait(&acc_free[buf], (uint32_t)(((ti >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * N);
        for (int kc = 0; kc < KC; ++kc, ++q) {
          const int s = (int)(q & 1);
          mbar_wait(&op_full[s], (uint32_t)((q >> 1) & 1));
          if (!C::kWRes) mbar_wait(&w_full, (uint32_t)(q & 1));
          tc_fence_after();
          const uint32_t a0 = smem_u32(sOp) + (uint32_t)s * kChunkBytes;
          const uint32_t w0 = smem_u32(sW) + (C::kWRes ? (uint32_t)kc * (uint32_t)N * 256u : 0u);
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const uint32_t blk = k8 >> 2, sub = k8 & 3;
            umma_tf32(acc, umma_desc(a0 + blk * kBlockBytes + sub * 32, 16, 1024),
                      umma_desc(w0 + blk * (uint32_t)N * 128u + sub * 32, 16, 1024), idesc, (kc | k8) ? 1u : 0u);
          }
          tc_commit(&op_free[s]);
          if (!C::kWRes) {
            tc_commit(&w_free);
            if (q + 1 < total_chunks) {                   // single weight buffer: reload once these MMAs have read it
              mbar_wait(&w_free, (uint32_t)(q & 1));
              const int kn = (kc + 1) % KC;
              mbar_expect_tx(&w_full, C::kWBytes);
              bulk_load(sW, (const uint8_t*)p.wimg + (size_t)kn * C::kWBytes, C::kWBytes, &w_full);
            }
          }
        }
        tc_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp > kMmaWarp) {
    // ================================================================================ builders
    // work items = (chunk q, round rd) in order; the loads of item i+1 are issued before item i is processed
    const int bt = tid - (kMmaWarp + 1) * 32;
    constexpr int R = C::kBldRounds;
    const long long total_items = my_tiles * KC * R;
    Item<PRO, V, K, G, C::kLoads> cur, nxt;
    auto issue = [&](long long item, Item<PRO, V, K, G, C::kLoads>& it) {
      const long long q = item / R;
      const int rd = (int)(item - q * R);
      const long long ti = q / KC;
      const int kc = (int)(q - ti * KC);
      issue_loads<PRO, EPI, V, K, N>(p, blockIdx.x + ti * gridDim.x, kc, bt + rd * kBldThreads, it);
    };
    if (total_items > 0) issue(0, cur);
    for (long long item = 0; item < total_items; ++item) {
      if (C::kPrefetch && item + 1 < total_items) issue(item + 1, nxt);
      const long long q = item / R;
      const int rd = (int)(item - q * R);
      const int s = (int)(q & 1);
      if (rd == 0 && q >= 2) mbar_wait(&op_free[s], (uint32_t)(((q >> 1) - 1) & 1));
      build_item<PRO, EPI, V, K, N>(p, sOp + (size_t)s * kChunkBytes, cur);
      if (rd == R - 1) {
        fence_proxy_async();
        mbar_arrive(&op_full[s]);
      }
      if (C::kPrefetch) cur = nxt;
      else if (item + 1 < total_items) issue(item + 1, cur);
    }
  } else {
    // ================================================================================ epilogue
    const int et = tid;                                            // 0 .. kEpiThreads-1
    const float* res = p.res ? p.res : p.in0;                      // absent residual: alias a valid tensor (branch-free loads)
    const float rsel = p.res ? 1.f : 0.f;
    constexpr int NACC = EPI == EPI_ROT_RAW ? NCH * C::kEpiRounds : 1;
    float s1[NACC][4], s2[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) s1[a][c] = 0.f, s2[a][c] = 0.f;

    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = blockIdx.x + ti * gridDim.x;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const size_t row0 = (size_t)g0 * V;
      const int buf = (int)(ti & 1);
      if (warp < 8) {    // only the TMEM readers (who gate acc_free) wait for the accumulator, see spatial_bwd.cu
        mbar_wait(&acc_full[buf], (uint32_t)((ti >> 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int nc = 0; nc < NCH; ++nc) {
        if (warp < 8) {   // TMEM -> staging: lane quarter (warp & 3), column half (warp >> 2)
          const int qd = warp & 3, hf = warp >> 2;
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * N + nc * 64 + hf * 32), v);
          const uint32_t row = (uint32_t)(qd * 32 + lane);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *(float4*)(sSt + stage_off(row, (uint32_t)(hf * 8 + i), 0)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (nc == NCH - 1) {                                     // last read of this accumulator buffer
            tc_fence_before();
            mbar_arrive(&acc_free[buf]);
          }
        }
        epi_sync();
#pragma unroll
        for (int rd = 0; rd < C::kEpiRounds; ++rd) {
          const int slot = et + rd * kEpiThreads;
          if (slot < C::kSlots) {
            const int v = slot >> 4, c4 = slot & 15;
            const int d = nc * 64 + c4 * 4;
            const float4 bias = p.bias ? ldg4(p.bias + d) : make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t o = (row0 + v) * N + d;
            if (EPI == EPI_LINEAR) {
#pragma unroll
              for (int g = 0; g < G; ++g)
                if (g < ng) {
                  float4 y = *(const float4*)(sSt + stage_off((uint32_t)(g * V + v), (uint32_t)c4, 0));
                  y.x += bias.x, y.y += bias.y, y.z += bias.z, y.w += bias.w;
                  if (p.relu) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
                  *(float4*)(p.out + o + (size_t)g * V * N) = y;
                }
            } else {
              // z[(g,v), d+j] = y[(g, (v-d-j) % V), d+j] + bias
              int u[4];
              u[0] = v - d % V;
              if (u[0] < 0) u[0] += V;
#pragma unroll
              for (int j = 1; j < 4; ++j) {
                u[j] = u[j - 1] - 1;
                if (u[j] < 0) u[j] += V;
              }
              float4 sc, sh, rv[G];
              if (EPI == EPI_ROT_FUSED) {
                sc = ldg4(p.epi_a + v * N + d);
                sh = ldg4(p.epi_b + v * N + d);
#pragma unroll
                for (int g = 0; g < G; ++g) rv[g] = ldg4(res + o + (size_t)min(g, ng - 1) * V * N);
              }
#pragma unroll
              for (int g = 0; g < G; ++g)
                if (g < ng) {
                  float4 z;
                  z.x = *(const float*)(sSt + stage_off((uint32_t)(g * V + u[0]), (uint32_t)c4, 0)) + bias.x;
                  z.y = *(const float*)(sSt + stage_off((uint32_t)(g * V + u[1]), (uint32_t)c4, 1)) + bias.y;
                  z.z = *(const float*)(sSt + stage_off((uint32_t)(g * V + u[2]), (uint32_t)c4, 2)) + bias.z;
                  z.w = *(const float*)(sSt + stage_off((uint32_t)(g * V + u[3]), (uint32_t)c4, 3)) + bias.w;
                  if (EPI == EPI_ROT_RAW) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int a = nc * C::kEpiRounds + rd;
                    s1[a][0] += z.x, s1[a][1] += z.y, s1[a][2] += z.z, s1[a][3] += z.w;
                    s2[a][0] = fmaf(z.x, z.x, s2[a][0]), s2[a][1] = fmaf(z.y, z.y, s2[a][1]);
                    s2[a][2] = fmaf(z.z, z.z, s2[a][2]), s2[a][3] = fmaf(z.w, z.w, s2[a][3]);
                  } else {
                    z.x = fmaf(z.x, sc.x, sh.x) + rsel * rv[g].x;
                    z.y = fmaf(z.y, sc.y, sh.y) + rsel * rv[g].y;
                    z.z = fmaf(z.z, sc.z, sh.z) + rsel * rv[g].z;
                    z.w = fmaf(z.w, sc.w, sh.w) + rsel * rv[g].w;
                    if (p.relu) z.x = fmaxf(z.x, 0.f), z.y = fmaxf(z.y, 0.f), z.z = fmaxf(z.z, 0.f), z.w = fmaxf(z.w, 0.f);
                  }
                  *(float4*)(p.out + o + (size_t)g * V * N) = z;
                }
            }
          }
        }
        epi_sync();                                                // staging is reused by the next chunk / tile
      }
    }
    if (EPI == EPI_ROT_RAW) {   // flush the per-(v,d) batch statistics
#pragma unroll
      for (int nc = 0; nc < NCH; ++nc)
#pragma unroll
        for (int rd = 0; rd < C::kEpiRounds; ++rd) {
          const int slot = et + rd * kEpiThreads;
          if (slot < C::kSlots) {
            const int v = slot >> 4, d = nc * 64 + (slot & 15) * 4;
            const int a = nc * C::kEpiRounds + rd;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              atomicAdd(p.stats + 2 * ((size_t)v * N + d + j), (double)s1[a][j]);
              atomicAdd(p.stats + 2 * ((size_t)v * N + d + j) + 1, (double)s2[a][j]);
            }
          }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}

template <int PRO, int EPI, int V, int K, int N>
static int launch(const SgcnRowGemm& p, cudaStream_t s) {
  using C = Cfg<PRO, EPI, V, K, N>;
  auto kern = fused_gemm_kernel<PRO, EPI, V, K, N>;
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem);
    if (e != cudaSuccess) return set_cuda_error("fused_gemm smem attribute", e);
    configured = true;
  }
  const long long ntiles = (p.groups + C::G - 1) / C::G;
  if (ntiles == 0) return 0;
  long long grid = num_sms();
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kThreads, C::kSmem, s>>>(p);
  return check_launch("fused_gemm_kernel");
}

}  // namespace fg
}  // namespace sgcn
