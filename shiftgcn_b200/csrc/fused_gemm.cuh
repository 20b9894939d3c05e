// fused_gemm.cuh -- the warp-specialised "row tile x weight" kernel behind the forward contractions of the
// Shift-GCN hot path (the backward-data pass of the spatial unit has its own copy of this skeleton with four input
// streams in the epilogue, spatial_bwd.cu).
//
// Activations live channels-last: a logical (n, C, T, V) tensor is the row-major matrix [(n,t,v), C].  A CTA (one
// per SM, persistent) processes tiles of G whole (n,t) groups (G*V <= 128 rows), so the joint shifts of the spatial
// unit are permutations inside the tile.  Seven warpgroups, four roles (register budgets set with setmaxnreg):
//
//   row loader 1 lane    TMA (cp.async.bulk.tensor, one tensor map per input): the raw [rows x 64 channel] box of every
//                        chunk goes into a ring of dense shared-memory stages, several chunks ahead, with mbarrier
//                        complete_tx; halo rows in front of the tensor and rows past a partial last tile arrive as
//                        zeros.  PRO_PLAIN: the {32 channels, V, G} boxes land in the SWIZZLE_128B operand stage
//                        directly and no thread touches the data before the tensor core does.
//   builders   12 warps  wait for a raw stage (each warp on its own: no barrier between the builder warps), then run
//                        the prologue as a
//                        shared -> shared pass with lane <-> channel: whatever row a lane reads (joint shift:
//                        row (u+c) % V, temporal shift: frame t + floor(ypos_c)), the bank is the channel, so the
//                        gathers are conflict free; results go TF32-rounded into the K-major SWIZZLE_128B operand
//                        chunk.  PRO_PLAIN has no prologue: cp.async writes the operand layout directly.
//   MMA        1 lane    tcgen05.mma kind::tf32 (M = 128 rows, N = out channels, fp32 accumulators in TMEM,
//                        two accumulator buffers)
//   W loader   1 lane    when the weight image exceeds 64 KiB: streams its 32-channel blocks with cp.async.bulk
//                        (TMA) through a two-stage ring, decoupled from the MMA issuer
//   epilogue   12 warps  TMEM -> XOR-swizzled smem staging (8 warps), then lane <-> channel again: the output
//                        rotation is a conflict-free row gather from the staging tile, global stores are whole
//                        128-byte row segments; BatchNorm statistics stay in registers for the whole kernel
//
// Variants (template PRO x EPI), with the reference code each one replaces:
//   PRO_SPATIAL  x[r,(u+c)%V,c] * (tanh(mask)+1)            model/shift_gcn.py:123-129
//   PRO_LERP     temporal shift of BN(h), zero padded       model/shift_gcn.py:66-68, shift_cuda_kernel.cu:12-76
//   PRO_PLAIN    rows as they are                           (backward data contraction of the 1x1 conv)
//   EPI_ROT_RAW    + bias, rotate z[v,d]=y[(v-d)%V,d], store, per-(v,d) batch statistics   :131-137 (training)
//   EPI_ROT_FUSED  + bias, rotate, folded BN, + residual, ReLU                              :131-141 (eval)
//   EPI_LINEAR     + bias, optional ReLU, store                                             :69-70
//   EPI_TSHIFT     + bias, ReLU, fractional OUTPUT shift, folded BN, + residual, ReLU        :69-73, 161-162 (eval, stride 1)
//                  -- the whole temporal unit in one kernel.  Tile geometry "joints x frames": the template V is a joint
//                  SUBSET (5 of 25) and the G = 25 row groups are consecutive FRAMES of one sample, so the second shift
//                  finds its taps inside the tile; a tile yields G - kWo output frames (q is recomputed for the kWo halo
//                  frames, 12 %) and q never crosses HBM.  Needs all floor(ypos_out) inside one window of kWo values.
#pragma once
#include "capi_internal.h"
#include "common.cuh"
#include "rowgemm.h"
#include "tensormap.h"
#include <string.h>
#include <type_traits>

namespace sgcn {
namespace fg {

enum { PRO_SPATIAL = 0, PRO_LERP = 1, PRO_PLAIN = 2 };
enum { EPI_ROT_RAW = 0, EPI_ROT_FUSED = 1, EPI_LINEAR = 2, EPI_TSHIFT = 4 };   // (3 = the spatial backward, spatial_bwd.cu)

constexpr int kEpiWarps = 12, kBldWarps = 12;
constexpr int kEpiThreads = kEpiWarps * 32, kBldThreads = kBldWarps * 32;
constexpr int kMmaWarp = kEpiWarps, kLoadWarp = kEpiWarps + 1, kRowWarp = kEpiWarps + 2;   // warps 12..15: MMA, W loader, row loader, spare
constexpr int kBld0 = kEpiWarps + 4;                             // first builder warp
constexpr int kThreads = (kEpiWarps + 4 + kBldWarps) * 32;       // 896
constexpr int kChunkBytes = 128 * 64 * 4;                        // one [128 x 64] fp32 operand chunk / staging tile
constexpr int kRegsEpi = 88, kRegsBld = 64, kRegsMma = 40;       // 384*88 + 384*64 + 128*40 <= 896*72

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ void bld_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kBldThreads) : "memory"); }
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg((const float4*)p); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// P3 = fp32-accurate mode (3xTF32, SgcnRowGemm::prec): every operand stage holds the TF32 head of the chunk followed by
// its TF32 tail, the weight image is {head image, tail image} (sgcn_prep_weight_image_split) and the issuer runs three
// MMA groups per 32-channel block: Ah*Wh, Al*Wh, Ah*Wl.
template <int PRO, int EPI, int V, int K, int N, bool P3 = false>
struct Cfg {
  static constexpr int G = 128 / V;
  static constexpr int KC = K / 64, NCH = N / 64;
  static constexpr int kImgBytes = K * N * 4;                      // one canonical weight image
  static constexpr int kOpBytes = (P3 ? 2 : 1) * kChunkBytes;      // one operand stage
  // weights: resident image(s), or 32-channel blocks streamed through a ring (the temporal-shift variant gives shared
  // memory to its wider raw tiles instead)
  static constexpr bool kWRes = ((P3 ? 2 : 1) * kImgBytes) <= (PRO == PRO_LERP ? 32768 : 65536);
  static constexpr int kWStage = N * 128;                          // one streamed 32-channel weight block
  static constexpr int kWB = P3 ? 4 : 2;                           // streamed blocks per chunk: (head, tail) x two halves
  static constexpr int kWRing = (P3 && PRO == PRO_LERP && N >= 256) ? 1 : 2;   // ring depth of the streamed blocks
  static constexpr int kWBytes = kWRes ? (P3 ? 2 : 1) * kImgBytes : kWRing * kWStage;
  // lane <-> channel work split: a warp owns (joint, 32-channel half) pairs k * 12 + warp, k < kJ
  static constexpr int kPairs = 2 * V;
  static constexpr int kJ = (kPairs + kBldWarps - 1) / kBldWarps;  // kBldWarps == kEpiWarps
  // flattened 16-byte pieces of a [rows x 64] chunk: piece = thread + 384 * i  ->  row = (thread >> 4) + 24 * i
  static constexpr int kTileRows = G * V;
#ifndef SGCN_LERP_WIN64
#define SGCN_LERP_WIN64 3
#endif
#ifndef SGCN_LERP_WIN128
#define SGCN_LERP_WIN128 3
#endif
#ifndef SGCN_LERP_WIN256
#define SGCN_LERP_WIN256 3
#endif
  // distinct floor(ypos) values the raw tile covers (channels outside take global taps: slow, see the builders)
  static constexpr int kWin = PRO != PRO_LERP ? 0
                              : (EPI == EPI_TSHIFT ? (K == 64 ? 3 : 2)
                                                   : (K == 64 ? (P3 ? 3 : SGCN_LERP_WIN64) : (K == 128 ? SGCN_LERP_WIN128 : SGCN_LERP_WIN256)));
  static constexpr int kRawRows = (G + kWin) * V;
  static constexpr int kRawBytes = PRO == PRO_PLAIN ? 0 : kRawRows * 256;
  // epilogue staging tile.  ROT_*: [128 rows x 64 channels] with a row pitch of 68 floats -- "thread = row" 128-bit
  // stores are conflict free (8 consecutive rows cover the eight 16-byte bank groups) and the rotated "lane = channel"
  // reads need no swizzle arithmetic: bank = (4*row + channel) % 32, at most 2-way conflicts where the rotation wraps
  static constexpr int kStPitch = 272;
  static constexpr int kStBytes = EPI == EPI_LINEAR ? 128 * 32 * 4 : 128 * kStPitch;
  // EPI_TSHIFT: window of floor(ypos_out) values and output frames per tile
  static constexpr int kWo = 3;
  static constexpr int kFo = G - kWo;
  // HBM latency x bandwidth needs >= 64 KiB of loads in flight per SM: spend what is left on input stages
  static constexpr int kAvail = 232448 - 1600 - kWBytes - kStBytes;
  static constexpr int kRaw2 = PRO == PRO_PLAIN ? 0 : (kAvail - 2 * kOpBytes) / kRawBytes;
  static constexpr int kRaw1 = PRO == PRO_PLAIN ? 0 : (kAvail - kOpBytes) / kRawBytes;
  static constexpr int kOpStages = PRO == PRO_PLAIN ? (kAvail / kOpBytes < 4 ? kAvail / kOpBytes : 4) : (kRaw2 >= 3 ? 2 : 1);
  static constexpr int kRawStages = PRO == PRO_PLAIN ? 0 : ((kOpStages == 2 ? kRaw2 : kRaw1) < 4 ? (kOpStages == 2 ? kRaw2 : kRaw1) : 4);
  static constexpr size_t kSmem = 1024 + kWBytes + kOpStages * kOpBytes + kStBytes + kRawStages * kRawBytes + 64;
  static_assert(PRO == PRO_PLAIN || kRawStages >= 2, "need at least two raw stages");
  static_assert(kOpStages >= (PRO == PRO_PLAIN ? 2 : 1), "need operand stages");
  static_assert(kSmem <= 232448 - 512, "shared memory budget");
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int PRO, int EPI, int V, int K, int N, bool P3>
__global__ void __launch_bounds__(kThreads, 1) fused_gemm_kernel(const SgcnRowGemm p, const int rev,
                                                                 const __grid_constant__ CUtensorMap tm0,
                                                                 const __grid_constant__ CUtensorMap tm1) {
  using C = Cfg<PRO, EPI, V, K, N, P3>;
  constexpr int kOpBytes = C::kOpBytes;
  constexpr int G = C::G, KC = C::KC, NCH = C::NCH, OS = C::kOpStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // opaque to the compiler: otherwise the shared-window base (S2R CgaCtaId + four integer ops) and the lane-derived
  // offsets are re-derived in front of every group of shared-memory accesses of the builder / epilogue loops
  asm volatile("" : "+r"(smem0));
  constexpr int RS = C::kRawStages;
  const uint32_t sW = smem0;
  const uint32_t sOp = sW + C::kWBytes;
  const uint32_t sSt = sOp + OS * kOpBytes;
  const uint32_t sRaw = sSt + C::kStBytes;
  __shared__ uint64_t op_full[4], op_free[4], acc_full[2], acc_free[2], w_full[2], w_free[2], raw_full[4], raw_free[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ int lerp_hist[16], lerp_lo_s, out_lo_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool kTmaOp = PRO == PRO_PLAIN && !P3;                 // operand stages filled by TMA alone
  constexpr bool TV = EPI == EPI_TSHIFT;                           // "joints x frames" tiles (V = joint subset, groups = frames)
  // Built operands are handed to the tensor core per 32-channel BLOCK (two per chunk, barriers 2*stage + block): the
  // builders finish block 0 of a chunk, signal it and carry on with block 1 while its MMAs already run, and they wait
  // for the tensor core per block as well.  With the single operand stage most shapes can afford (shared memory goes
  // to the raw-input ring) the chunk-granular handshake left the builders idle for a quarter of the kernel (ncu: 26 %
  // of their samples on the op_free wait) and starved the epilogue behind them.
  // It costs the builders a second set of per-channel constants per chunk (a warp then works in both blocks), which
  // only pays where the MMAs of a chunk are long: the temporal GEMMs from 128 channels on (lerp/linear 106 -> 99 us at
  // C=128, 123 -> 110 us at C=256; at C=64 it lost 18 %, and the spatial / fused-epilogue kernels are epilogue-bound).
#ifndef SGCN_BLKPIPE
#define SGCN_BLKPIPE 1
#endif
  constexpr bool kBlkPipe = SGCN_BLKPIPE && PRO == PRO_LERP && EPI == EPI_LINEAR && K >= 128;
  static_assert(!kBlkPipe || 2 * OS <= 4, "block-granular handshake: at most four block stages");

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&op_full[i], kTmaOp ? 1 : kBldThreads);
      mbar_init(&op_free[i], 1);
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_free[i], kBldWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 8 * 32);
      mbar_init(&w_full[i], 1);
      mbar_init(&w_free[i], 1);
    }
    fence_mbar_init();
  }
  if constexpr (PRO == PRO_LERP) {
    // window of floor(ypos) values the raw tile covers: the `win` consecutive values holding most channels
    auto pick_window = [&](const float* ypos, int nch, int win, int* result) {
      if (tid < 16) lerp_hist[tid] = 0;
      __syncthreads();
      for (int c = tid; c < nch; c += kThreads) {
        const int fl = (int)floorf(__ldg(ypos + c));
        if (fl >= -8 && fl < 8) atomicAdd(&lerp_hist[fl + 8], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int best = 0, bestn = -1;
        for (int lo = 0; lo + win <= 16; ++lo) {
          int n = 0;
          for (int j = 0; j < win; ++j) n += lerp_hist[lo + j];
          if (n > bestn) bestn = n, best = lo;
        }
        *result = best - 8;
      }
      __syncthreads();
    };
    pick_window(p.pro_c, K, C::kWin, &lerp_lo_s);
    if constexpr (TV) pick_window(p.res2, N, C::kWo, &out_lo_s);   // EPI_TSHIFT: res2 = effective output shift positions
  }
  constexpr uint32_t tmem_cols = 2 * N <= 128 ? 128u : (2 * N <= 256 ? 256u : 512u);
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, tmem_cols);
  if (C::kWRes) {
    for (int i = tid; i < C::kWBytes / 16; i += kThreads) cp16(sW + (uint32_t)i * 16u, (const uint8_t*)p.wimg + (size_t)i * 16);
    cp_async_commit();
    cp_async_wait_all();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // TV tiles: (sample, frame block of kFo output frames, joint subset); p.V is the tensor's joint count, p.groups = n * T
  const int tvNJ = TV ? p.V / V : 1, tvFB = TV ? (p.T + C::kFo - 1) / C::kFo : 1;
  const long long ntiles = TV ? (p.groups / (p.T > 0 ? p.T : 1)) * tvFB * tvNJ : (p.groups + G - 1) / G;
  const int my_tiles = (long long)blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  struct TvTile {
    int n, fb, js, tq0;                                            // sample, frame block, joint subset, first staged q frame
  };
  auto tv_tile = [&](long long tile) -> TvTile {
    TvTile t;
    t.js = (int)(tile % tvNJ);
    const long long r = tile / tvNJ;
    t.fb = (int)(r % tvFB);
    t.n = (int)(r / tvFB);
    t.tq0 = t.fb * C::kFo + out_lo_s;
    return t;
  };
  const int total_chunks = my_tiles * KC;
  // snake traversal (capi_internal.h): rev walks the tiles from the last one down
  auto tile_of = [&](int ti) -> long long {
    const long long t = (long long)blockIdx.x + (long long)ti * gridDim.x;
    return rev ? ntiles - 1 - t : t;
  };

  if (warp >= kMmaWarp && warp < kBld0) {
    reg_dec<kRegsMma>();
    if (warp == kMmaWarp && lane == 0) {
      // ================================================================================ MMA issuer
      const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
      // K-major SWIZZLE_128B descriptors differ only in the 14-bit start-address field (bytes >> 4) of the low word
      const uint64_t d0 = umma_desc(0, 16, 1024);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo = (uint32_t)d0;
      int q = 0;
#pragma unroll 1
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) mbar_wait(&acc_free[buf], (uint32_t)(((ti >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * N);
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc, ++q) {
          const int s = q % OS;
          if constexpr (!kBlkPipe) {
            mbar_wait(&op_full[s], (uint32_t)((q / OS) & 1));
            tc_fence_after();
          }
          const uint32_t a_lo = dlo | ((sOp + (uint32_t)s * kOpBytes) >> 4);      // head half of the stage
          const uint32_t a_tail = a_lo + (uint32_t)(kChunkBytes >> 4);                // tail half (P3)
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            if constexpr (kBlkPipe) {
              mbar_wait(&op_full[2 * s + blk], (uint32_t)((q / OS) & 1));
              tc_fence_after();
            }
            // weight block (head image; in P3 mode followed by the same block of the tail image)
#pragma unroll
            for (int part = 0; part < (P3 ? 2 : 1); ++part) {
              uint32_t w0;
              int slot = 0;
              if (C::kWRes) {
                w0 = sW + (uint32_t)(part * C::kImgBytes) + (uint32_t)(kc * 2 + blk) * (uint32_t)C::kWStage;
              } else {
                const int h = C::kWB * q + (P3 ? 2 * blk + part : blk);   // streamed 32-channel weight block
                slot = h % C::kWRing;
                mbar_wait(&w_full[slot], (uint32_t)((h / C::kWRing) & 1));
                tc_fence_after();
                w0 = sW + (uint32_t)slot * (uint32_t)C::kWStage;
              }
              const uint32_t b_lo = dlo | (w0 >> 4);
#pragma unroll
              for (int sub = 0; sub < 4; ++sub)
                umma_tf32(acc, ((uint64_t)dhi << 32) | (a_lo + (uint32_t)(blk * (kBlockBytes >> 4) + sub * 2)),
                          ((uint64_t)dhi << 32) | (b_lo + (uint32_t)(sub * 2)), idesc, (kc | blk | sub | part) ? 1u : 0u);
              if (P3 && part == 0) {                                 // tail of the activations against the weight heads
#pragma unroll
                for (int sub = 0; sub < 4; ++sub)
                  umma_tf32(acc, ((uint64_t)dhi << 32) | (a_tail + (uint32_t)(blk * (kBlockBytes >> 4) + sub * 2)),
                            ((uint64_t)dhi << 32) | (b_lo + (uint32_t)(sub * 2)), idesc, 1u);
              }
              if (!C::kWRes) tc_commit(&w_free[slot]);
            }
            if constexpr (kBlkPipe) tc_commit(&op_free[2 * s + blk]);
          }
          if constexpr (!kBlkPipe) tc_commit(&op_free[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    } else if (warp == kLoadWarp && lane == 0 && !C::kWRes) {
      // ================================================================================ weight loader (TMA ring)
      const int total_blocks = C::kWB * total_chunks;
      for (int h = 0; h < total_blocks; ++h) {
        const int s = h % C::kWRing;
        if (h >= C::kWRing) mbar_wait(&w_free[s], (uint32_t)(((h / C::kWRing) - 1) & 1));
        const int kc = (h / C::kWB) % KC, j = h % C::kWB;            // P3: j = 2 * half + part, else j = half
        const size_t src = P3 ? (size_t)(j & 1) * C::kImgBytes + (size_t)(kc * 2 + (j >> 1)) * C::kWStage
                              : (size_t)(kc * 2 + j) * C::kWStage;
        mbar_expect_tx(&w_full[s], C::kWStage);
        bulk_load(sW + (uint32_t)s * (uint32_t)C::kWStage, (const uint8_t*)p.wimg + src, C::kWStage, &w_full[s]);
      }
    } else if (warp == kRowWarp && lane == 0 && (PRO != PRO_PLAIN || kTmaOp)) {
      // ================================================================================ row loader (TMA tensor maps)
      tma_prefetch_map(&tm0);
      if constexpr (PRO == PRO_PLAIN) {
        tma_prefetch_map(&tm1);
        // {32 channels, V joints, G groups} boxes = whole 32-channel operand blocks in the SWIZZLE_128B layout.  The K
        // channels may come from two row tensors side by side (p.k0 from in0, the rest from in1: the input-gradient GEMM
        // of a conv + BN side branch contracts [g | x] in one pass), each with its own frame stride (inside its map).
        const int k0 = p.k0 > 0 ? p.k0 : K;
        for (int q = 0; q < total_chunks; ++q) {
          const int s = q % OS, ti = q / KC, kc = q - ti * KC;
          if (q >= OS) mbar_wait(&op_free[s], (uint32_t)(((q / OS) - 1) & 1));
          const int g0 = (int)(tile_of(ti) * G);
          const bool second = kc * 64 >= k0;
          const CUtensorMap* map = second ? &tm1 : &tm0;
          const int ch = second ? kc * 64 - k0 : kc * 64;
          const uint32_t dst = sOp + (uint32_t)s * kOpBytes, bar = smem_u32(&op_full[s]);
          mbar_expect_tx(&op_full[s], 2u * (uint32_t)(G * V * 128));
          tma_load_3d(dst, map, ch, 0, g0, bar);
          tma_load_3d(dst + kBlockBytes, map, ch + 32, 0, g0, bar);
        }
      } else {
        const int lerp_lo = PRO == PRO_LERP ? lerp_lo_s : 0;
        for (int q = 0; q < total_chunks; ++q) {
          const int s = q % RS, ti = q / KC, kc = q - ti * KC;
          if (q >= RS) mbar_wait(&raw_free[s], (uint32_t)(((q / RS) - 1) & 1));
          mbar_expect_tx(&raw_full[s], (uint32_t)C::kRawBytes);
          if constexpr (TV) {                                      // box {64 channels, V joints, G + kWin frames}
            const TvTile t = tv_tile(tile_of(ti));
            tma_load_3d(sRaw + (uint32_t)s * C::kRawBytes, &tm0, kc * 64, t.js * V, t.n * p.T + t.tq0 + lerp_lo,
                        smem_u32(&raw_full[s]));
          } else {
            const long long gfirst = tile_of(ti) * G + lerp_lo;    // first group of the raw tile (may be < 0: zeros)
            tma_load_2d(sRaw + (uint32_t)s * C::kRawBytes, &tm0, kc * 64, (int)(gfirst * V), smem_u32(&raw_full[s]));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= kBld0) {
    // ================================================================================ builders
    reg_dec<kRegsBld>();
    const int bt = tid - kBld0 * 32, bw = bt >> 5;
    const uint32_t lq = (uint32_t)lane >> 2;
    const int prow = bt >> 4, pc4 = bt & 15;                       // this thread's 16-byte piece of rows prow + 24*i
    constexpr int kPieces = (C::kRawRows + 23) / 24;               // copies per thread and chunk (PLAIN: tile rows)

    if constexpr (kTmaOp) {
      // nothing to build: the row loader's TMA boxes ARE the operand blocks
    } else if constexpr (PRO == PRO_PLAIN) {
      // fp32-accurate mode: cp.async straight into the swizzled operand stage, then every thread splits its own pieces
      const uint32_t dst0 = (uint32_t)(pc4 >> 3) * kBlockBytes + (uint32_t)prow * 128u + ((((uint32_t)pc4 ^ (uint32_t)prow) & 7u) << 4);
      // The K channels may come from two row tensors side by side (p.k0 from in0, the rest from in1: the input-gradient
      // GEMM of a conv + BN side branch contracts [g | x] in one pass), each with its own frame stride (1x1 conv, stride 2).
      const int k0 = p.k0 > 0 ? p.k0 : K;
      const bool simple = k0 == K && p.in0_gs <= 1;
      auto issue = [&](int q) {                                    // row & 7 is invariant under +24, so is the swizzle
        const int ti = q / KC, kc = q - ti * KC;
        const long long g0 = tile_of(ti) * G;
        const int nrow = (int)((p.groups - g0) < G ? (p.groups - g0) : G) * V;
        const uint32_t dst = sOp + (uint32_t)(q % OS) * kOpBytes + dst0;
        if (simple) {
          const float* src = p.in0 + ((size_t)g0 * V + prow) * K + kc * 64 + pc4 * 4;
#pragma unroll
          for (int i = 0; i < kPieces; ++i)
            if (prow + 24 * i < nrow) cp16(dst + (uint32_t)i * 3072u, src + (size_t)i * 24 * K);
        } else {
          const bool second = kc * 64 >= k0;
          const float* base = second ? p.in1 : p.in0;
          const int pitch = second ? K - k0 : k0, ch = (second ? kc * 64 - k0 : kc * 64) + pc4 * 4;
          const long long gs = second ? (p.in1_gs > 0 ? p.in1_gs : 1) : (p.in0_gs > 0 ? p.in0_gs : 1);
#pragma unroll
          for (int i = 0; i < kPieces; ++i) {
            const int row = prow + 24 * i;
            if (row < nrow) {
              const int gr = row / V;
              cp16(dst + (uint32_t)i * 3072u, base + ((size_t)((g0 + gr) * gs) * V + (row - gr * V)) * pitch + ch);
            }
          }
        }
      };
      for (int j = 0; j < OS - 1; ++j) {
        if (j < total_chunks) issue(j);
        cp_async_commit();
      }
      for (int q = 0; q < total_chunks; ++q) {
        const int qn = q + OS - 1;                                 // refill the stage chunk q-1 has just left
        if (qn < total_chunks) {
          if (q >= 1) mbar_wait_relaxed(&op_free[(q - 1) % OS], (uint32_t)(((q - 1) / OS) & 1));
          issue(qn);
        }
        cp_async_commit();
        cp_wait<OS - 1>();
        if constexpr (P3) {
          // split this thread's OWN pieces of chunk q (its cp.async copies have completed) into TF32 head (in place)
          // and tail (second half of the stage)
          const int ti = q / KC;
          const long long g0 = tile_of(ti) * G;
          const int nrow = (int)((p.groups - g0) < G ? (p.groups - g0) : G) * V;
          const uint32_t dst = sOp + (uint32_t)(q % OS) * kOpBytes + dst0;
#pragma unroll
          for (int i = 0; i < kPieces; ++i)
            if (prow + 24 * i < nrow) {
              const float4 v = lds128(dst + (uint32_t)i * 3072u);
              float4 hi, lo;
              split_tf32(v.x, hi.x, lo.x), split_tf32(v.y, hi.y, lo.y), split_tf32(v.z, hi.z, lo.z), split_tf32(v.w, hi.w, lo.w);
              sts128(dst + (uint32_t)i * 3072u, hi);
              sts128(dst + (uint32_t)i * 3072u + (uint32_t)kChunkBytes, lo);
            }
        }
        fence_proxy_async();
        mbar_arrive(&op_full[q % OS]);
      }
    } else {
      // ---- raw chunks arrive by TMA (row loader); every builder warp waits and releases on its own
      const int lerp_lo = PRO == PRO_LERP ? lerp_lo_s : 0;
      uint32_t opb = sOp + (uint32_t)(lane & 3) * 4u, rawb = sRaw + (uint32_t)lane * 4u, lqs = lq << 4;
      asm volatile("" : "+r"(opb), "+r"(rawb), "+r"(lqs));
      if constexpr (!kBlkPipe) {
      // ---- chunk-granular hand-over: unit = 2 * joint + block, a warp stays in block (warp & 1)
      for (int q = 0; q < total_chunks; ++q) {
        const int ti = q / KC, kc = q - ti * KC;
        const long long g0 = TV ? 0 : tile_of(ti) * G;
        const int ng = TV ? G : (int)((p.groups - g0) < G ? (p.groups - g0) : G);
        // the per-(joint, channel) / per-channel tables of this chunk are requested BEFORE the wait for its raw rows:
        // loaded right in front of their use, every (joint, half) pair paid a full global-load latency (ncu: 28 % of
        // the builder warps' stall samples sat on the first multiply of each pair)
        const int tc = kc * 64 + (bw & 1) * 32 + lane;             // pairs k*12 + bw keep the parity of bw
        float tab[PRO == PRO_SPATIAL ? C::kJ : 3];
        if constexpr (PRO == PRO_SPATIAL) {
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) tab[k] = __ldg(p.pro_a + min((k * kBldWarps + bw) >> 1, V - 1) * K + tc);
        } else {
          tab[0] = __ldg(p.pro_c + tc), tab[1] = __ldg(p.pro_a + tc), tab[2] = __ldg(p.pro_b + tc);
        }
        mbar_wait_relaxed(&raw_full[q % RS], (uint32_t)((q / RS) & 1));
        const int os = q % OS;
        if (q >= OS) mbar_wait_relaxed(&op_free[os], (uint32_t)(((q / OS) - 1) & 1));
        const uint32_t raw = rawb + (uint32_t)(q % RS) * C::kRawBytes;
        const uint32_t ob = opb + (uint32_t)os * kOpBytes;

        // Operand row r = g*V + v (128-byte pitch, 16-byte chunk XOR-ed with (r & 7)), 32-channel half, channel lane:
        //   address = ob + half*16K + r*128 + (((lane >> 2) ^ r) & 7) * 16 + (lane & 3) * 4
        // with everything but the swizzle term folded into a per-pair base and a compile-time offset.  Rows of groups
        // beyond a partial last tile are written too (their operand rows and outputs are never used).
        auto put = [&](uint32_t base, uint32_t v16, int g, float val) {
          const uint32_t addr = base + ((lqs ^ (v16 + (uint32_t)(g * V * 16))) & 0x70u) + (uint32_t)(g * V * 128);
          if constexpr (P3) {
            float hi, lo;
            split_tf32(val, hi, lo);
            sts32(addr, hi);
            sts32(addr + (uint32_t)kChunkBytes, lo);
          } else {
            sts32(addr, tf32_half_ulp(val));
          }
        };
        if constexpr (PRO == PRO_SPATIAL) {
          // xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]
          const int half = bw & 1;
          const int c = tc, cm = c % V;
          const uint32_t rbh = raw + (uint32_t)(half * 128), obh = ob + (uint32_t)(half * kBlockBytes);
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) {
            const int u = (k * kBldWarps + bw) >> 1;
            if (u < V) {
              int sv = u + cm;
              if (sv >= V) sv -= V;
              const float mm = tab[k];
              const uint32_t rb = rbh + (uint32_t)sv * 256u;
              float s[G];
#pragma unroll
              for (int g = 0; g < G; ++g) s[g] = lds32(rb + (uint32_t)(g * V * 256));
              const uint32_t b = obh + (uint32_t)u * 128u, u16 = (uint32_t)u << 4;
#pragma unroll
              for (int g = 0; g < G; ++g) put(b, u16, g, s[g] * mm);
            }
          }
        } else {
          // p[(g,v), c] = (1-f) U(t+y1) + f U(t+y1+1),  U = sa*h + sb inside the sample, 0 outside   (K1 with xpos = 0)
          const int T = p.T;
          TvTile tvt = {0, 0, 0, 0};
          if constexpr (TV) tvt = tv_tile(tile_of(ti));
          // frame of the tile's first group; TV tiles stay inside one sample and may start in front of it (t0 < 0)
          const int t0 = TV ? tvt.tq0 : (p.groups < (1ll << 31) ? (int)((unsigned)g0 % (unsigned)T) : (int)(g0 % T));
          const bool interior = ng == G && t0 + lerp_lo >= 0 && t0 + G - 1 + lerp_lo + C::kWin < T && (TV || t0 + G <= T);
          const int half = bw & 1;
          const int c = tc;
          const float ypos = tab[0], sa = tab[1], sb = tab[2];
          const float fl = floorf(ypos);
          const int y1 = (int)fl, idx = y1 - lerp_lo;
          const float f = ypos - fl, a1 = sa * f, a0 = sa - a1;
          const bool inwin = idx >= 0 && idx < C::kWin;
          const uint32_t rbh = raw + (uint32_t)(half * 128 + (inwin ? idx : 0) * (V * 256)), obh = ob + (uint32_t)(half * kBlockBytes);
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) {
            const int v = (k * kBldWarps + bw) >> 1;
            if (v < V) {
              const uint32_t b = obh + (uint32_t)v * 128u, v16 = (uint32_t)v << 4;
              float s[G + 1];
              if (inwin) {
                const uint32_t rb = rbh + (uint32_t)v * 256u;
#pragma unroll
                for (int g = 0; g <= G; ++g) s[g] = lds32(rb + (uint32_t)(g * V * 256));
              } else if constexpr (TV) {                           // ... global taps, frames clamped into the sample
                const float* col = p.in0 + (((size_t)tvt.n * T) * p.V + tvt.js * V + v) * K + c;
#pragma unroll
                for (int g = 0; g <= G; ++g) s[g] = __ldg(col + (size_t)min(max(t0 + g + y1, 0), T - 1) * p.V * K);
              } else {                                             // shift position outside the staged window: global taps
                const long long last = p.groups - 1;
#pragma unroll
                for (int g = 0; g <= G; ++g) {
                  long long gi = g0 + g + y1;
                  gi = gi < 0 ? 0 : (gi > last ? last : gi);
                  s[g] = __ldg(p.in0 + ((size_t)gi * V + v) * K + c);
                }
              }
              if (interior && inwin) {                             // every tap inside the sample: the affine commutes with the lerp
#pragma unroll
                for (int g = 0; g < G; ++g) put(b, v16, g, fmaf(a0, s[g], fmaf(a1, s[g + 1], sb)));
              } else {
                const float b1 = sb * f, b0 = sb - b1;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  int t = t0 + g;                                  // frame of this group (tiles may straddle samples)
                  if (!TV && t >= T) t -= T;
                  const float u0 = ((unsigned)(t + y1) < (unsigned)T) ? fmaf(a0, s[g], b0) : 0.f;
                  const float u1 = ((unsigned)(t + y1 + 1) < (unsigned)T) ? fmaf(a1, s[g + 1], b1) : 0.f;
                  put(b, v16, g, u0 + u1);
                }
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&op_full[os]);
        __syncwarp();                                              // every lane of this warp is done with the raw stage
        if (lane == 0) mbar_arrive(&raw_free[q % RS]);
      }
      } else {
      // ---- block-granular hand-over: units are block-major, unit = block * V + joint, dealt round-robin to the 12 warps
      // (unit k*12 + warp in round k), so a warp's units of block 0 come before its units of block 1 and block 0 is
      // complete after ~half of the rounds
      auto unit_of = [&](int k, int& hb, int& u) -> bool {
        const int i = k * kBldWarps + bw;
        hb = i >= V ? 1 : 0;
        u = i - hb * V;
        return i < 2 * V;
      };
      for (int q = 0; q < total_chunks; ++q) {
        const int ti = q / KC, kc = q - ti * KC;
        const long long g0 = TV ? 0 : tile_of(ti) * G;
        const int ng = TV ? G : (int)((p.groups - g0) < G ? (p.groups - g0) : G);
        // the per-(joint, channel) / per-channel tables of this chunk are requested BEFORE the wait for its raw rows:
        // loaded right in front of their use, every unit paid a full global-load latency (ncu: 28 % of the builder
        // warps' stall samples sat on the first multiply of each unit)
        const int tc0 = kc * 64 + lane;                            // this lane's channel in block 0 (block 1: + 32)
        float tab[PRO == PRO_SPATIAL ? C::kJ : 6];
        if constexpr (PRO == PRO_SPATIAL) {
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) {
            int hb, u;
            unit_of(k, hb, u);
            tab[k] = __ldg(p.pro_a + min(u, V - 1) * K + tc0 + hb * 32);
          }
        } else {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb)
            tab[3 * hb] = __ldg(p.pro_c + tc0 + hb * 32), tab[3 * hb + 1] = __ldg(p.pro_a + tc0 + hb * 32),
                     tab[3 * hb + 2] = __ldg(p.pro_b + tc0 + hb * 32);
        }
        mbar_wait_relaxed(&raw_full[q % RS], (uint32_t)((q / RS) & 1));
        const int os = q % OS;
        const uint32_t free_par = (uint32_t)(((q / OS) - 1) & 1);
        const uint32_t raw = rawb + (uint32_t)(q % RS) * C::kRawBytes;
        const uint32_t ob = opb + (uint32_t)os * kOpBytes;

        // Operand row r = g*V + v (128-byte pitch, 16-byte chunk XOR-ed with (r & 7)), 32-channel block, channel lane:
        //   address = ob + block*16K + r*128 + (((lane >> 2) ^ r) & 7) * 16 + (lane & 3) * 4
        // with everything but the swizzle term folded into a per-unit base and a compile-time offset.  Rows of groups
        // beyond a partial last tile are written too (their operand rows and outputs are never used).
        auto put = [&](uint32_t base, uint32_t v16, int g, float val) {
          const uint32_t addr = base + ((lqs ^ (v16 + (uint32_t)(g * V * 16))) & 0x70u) + (uint32_t)(g * V * 128);
          if constexpr (P3) {
            float hi, lo;
            split_tf32(val, hi, lo);
            sts32(addr, hi);
            sts32(addr + (uint32_t)kChunkBytes, lo);
          } else {
            sts32(addr, tf32_half_ulp(val));
          }
        };
        // Block hand-over.  EVERY builder thread arrives once per block and chunk, and only after it has seen the
        // tensor core release the block's previous use (an arrival of use n+1 must not be counted into phase n of the
        // barrier), whether or not this warp has units in the block.
        int cur = -1;                                              // block this warp is building (warp uniform)
        auto enter = [&](int hb) {                                 // hb > cur
          if (cur == 0) fence_proxy_async();
          if (cur < 0 && hb == 1 && q >= OS) mbar_wait_relaxed(&op_free[2 * os], free_par);   // no unit in block 0
          if (hb == 1) mbar_arrive(&op_full[2 * os]);
          if (q >= OS) mbar_wait_relaxed(&op_free[2 * os + hb], free_par);
          cur = hb;
        };
        auto leave = [&]() {
          if (cur >= 0) fence_proxy_async();
          if (cur < 1) {                                           // no unit in block 1 (joint-subset tiles), or no unit at all
            if (cur < 0 && q >= OS) mbar_wait_relaxed(&op_free[2 * os], free_par);
            mbar_arrive(&op_full[2 * os]);
            if (q >= OS) mbar_wait_relaxed(&op_free[2 * os + 1], free_par);
          }
          mbar_arrive(&op_full[2 * os + 1]);
        };
        if constexpr (PRO == PRO_SPATIAL) {
          // xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]
          int cm = 0;
          uint32_t rbh = 0, obh = 0;
          auto set_block = [&](int hb) {
            enter(hb);
            cm = (tc0 + hb * 32) % V;
            rbh = raw + (uint32_t)(hb * 128), obh = ob + (uint32_t)(hb * kBlockBytes);
          };
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) {
            int hb, u;
            if (unit_of(k, hb, u)) {
              if (hb != cur) set_block(hb);
              int sv = u + cm;
              if (sv >= V) sv -= V;
              const float mm = tab[k];
              const uint32_t rb = rbh + (uint32_t)sv * 256u;
              float s[G];
#pragma unroll
              for (int g = 0; g < G; ++g) s[g] = lds32(rb + (uint32_t)(g * V * 256));
              const uint32_t b = obh + (uint32_t)u * 128u, u16 = (uint32_t)u << 4;
#pragma unroll
              for (int g = 0; g < G; ++g) put(b, u16, g, s[g] * mm);
            }
          }
        } else {
          // p[(g,v), c] = (1-f) U(t+y1) + f U(t+y1+1),  U = sa*h + sb inside the sample, 0 outside   (K1 with xpos = 0)
          const int T = p.T;
          TvTile tvt = {0, 0, 0, 0};
          if constexpr (TV) tvt = tv_tile(tile_of(ti));
          // frame of the tile's first group; TV tiles stay inside one sample and may start in front of it (t0 < 0)
          const int t0 = TV ? tvt.tq0 : (p.groups < (1ll << 31) ? (int)((unsigned)g0 % (unsigned)T) : (int)(g0 % T));
          const bool interior = ng == G && t0 + lerp_lo >= 0 && t0 + G - 1 + lerp_lo + C::kWin < T && (TV || t0 + G <= T);
          int c = 0, y1 = 0;
          float sb = 0.f, f = 0.f, a1 = 0.f, a0 = 0.f;
          bool inwin = false;
          uint32_t rbh = 0, obh = 0;
          auto set_block = [&](int hb) {
            enter(hb);
            c = tc0 + hb * 32;
            const float ypos = hb ? tab[3] : tab[0], sa = hb ? tab[4] : tab[1];
            sb = hb ? tab[5] : tab[2];
            const float fl = floorf(ypos);
            y1 = (int)fl;
            const int idx = y1 - lerp_lo;
            f = ypos - fl, a1 = sa * f, a0 = sa - a1;
            inwin = idx >= 0 && idx < C::kWin;
            rbh = raw + (uint32_t)(hb * 128 + (inwin ? idx : 0) * (V * 256)), obh = ob + (uint32_t)(hb * kBlockBytes);
          };
#pragma unroll
          for (int k = 0; k < C::kJ; ++k) {
            int hb, v;
            if (unit_of(k, hb, v)) {
              if (hb != cur) set_block(hb);
              const uint32_t b = obh + (uint32_t)v * 128u, v16 = (uint32_t)v << 4;
              float s[G + 1];
              if (inwin) {
                const uint32_t rb = rbh + (uint32_t)v * 256u;
#pragma unroll
                for (int g = 0; g <= G; ++g) s[g] = lds32(rb + (uint32_t)(g * V * 256));
              } else if constexpr (TV) {                           // ... global taps, frames clamped into the sample
                const float* col = p.in0 + (((size_t)tvt.n * T) * p.V + tvt.js * V + v) * K + c;
#pragma unroll
                for (int g = 0; g <= G; ++g) s[g] = __ldg(col + (size_t)min(max(t0 + g + y1, 0), T - 1) * p.V * K);
              } else {                                             // shift position outside the staged window: global taps
                const long long last = p.groups - 1;
#pragma unroll
                for (int g = 0; g <= G; ++g) {
                  long long gi = g0 + g + y1;
                  gi = gi < 0 ? 0 : (gi > last ? last : gi);
                  s[g] = __ldg(p.in0 + ((size_t)gi * V + v) * K + c);
                }
              }
              if (interior && inwin) {                             // every tap inside the sample: the affine commutes with the lerp
#pragma unroll
                for (int g = 0; g < G; ++g) put(b, v16, g, fmaf(a0, s[g], fmaf(a1, s[g + 1], sb)));
              } else {
                const float b1 = sb * f, b0 = sb - b1;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  int t = t0 + g;                                  // frame of this group (tiles may straddle samples)
                  if (!TV && t >= T) t -= T;
                  const float u0 = ((unsigned)(t + y1) < (unsigned)T) ? fmaf(a0, s[g], b0) : 0.f;
                  const float u1 = ((unsigned)(t + y1 + 1) < (unsigned)T) ? fmaf(a1, s[g + 1], b1) : 0.f;
                  put(b, v16, g, u0 + u1);
                }
              }
            }
          }
        }
        leave();
        __syncwarp();                                              // every lane of this warp is done with the raw stage
        if (lane == 0) mbar_arrive(&raw_free[q % RS]);
      }
      }
    }
  } else {
    // ================================================================================ epilogue
    reg_inc<kRegsEpi>();
    const int et = tid;                                            // 0 .. kEpiThreads-1
    const float* res = p.res ? p.res : p.in0;                      // absent residual: alias a valid tensor (branch-free loads)
    const float rsel = p.res ? 1.f : 0.f;
    constexpr int NACC = EPI == EPI_ROT_RAW ? NCH * C::kJ : 1;
    float s1[NACC], s2[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) s1[a] = 0.f, s2[a] = 0.f;
    uint32_t stb = sSt + (uint32_t)(((warp & 1) * 32 + lane) * 4);   // this lane's channel column of the staging tile
    asm volatile("" : "+r"(stb));

    for (int ti = 0; ti < my_tiles; ++ti) {
      const long long g0 = tile_of(ti) * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const size_t row0 = (size_t)g0 * V;
      const int buf = ti & 1;
      int wv = warp;                                               // opaque per tile: keeps the per-(pair, chunk) address
      asm volatile("" : "+r"(wv));                                 // set from being hoisted out of the tile loop (registers)
      // ROT_FUSED: the BatchNorm tables and the residual rows of a 64-channel chunk do not depend on the accumulator;
      // they are requested before the wait for it (chunk 0) / before the TMEM copy (later chunks) instead of right in
      // front of their use, where every (joint, half) pair paid a full global-load latency.
      constexpr int kPF = EPI == EPI_ROT_FUSED ? C::kJ : 1;
      float scp[kPF], shp[kPF], rvp[kPF][G];
      auto prefetch_t = [&](int nc, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;          // full tile: one pointer per pair + immediate offsets
        const int d = nc * 64 + (wv & 1) * 32 + lane;              // (25 separate 64-bit pointers spilled to local memory)
        const float* rbase = res + row0 * N + d;
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
          const int o = min((k * kEpiWarps + wv) >> 1, V - 1) * N + d;
          scp[k] = __ldg(p.epi_a + o);
          shp[k] = __ldg(p.epi_b + o);
          const float* rptr = rbase + (o - d);
#pragma unroll
          for (int g = 0; g < G; ++g) rvp[k][g] = __ldg(rptr + (kFull ? g : min(g, ng - 1)) * V * N);
        }
      };
      auto prefetch = [&](int nc) {
        if (ng == G) prefetch_t(nc, std::true_type{});
        else prefetch_t(nc, std::false_type{});
      };
      if constexpr (EPI == EPI_ROT_FUSED) prefetch(0);
      if (warp < 8) {    // only the TMEM readers (who gate acc_free) wait for the accumulator, see spatial_bwd.cu
        mbar_wait_relaxed(&acc_full[buf], (uint32_t)((ti >> 1) & 1));
        tc_fence_after();
      }
      if constexpr (EPI == EPI_TSHIFT) {
        // q = relu(acc + bias) of frames tq0 .. tq0+G-1 and V joints sits in the accumulator; per 64-channel chunk it goes
        // through the staging tile (row = frame * V + joint), then lane <-> channel: a warp owns ONE (joint, half) pair and
        // walks the output frames with the two-tap filter of its channel's shift position, folded BatchNorm, residual and
        // ReLU -- the walker kernel tshift_fwd_apply, fed from shared memory (same arithmetic, same rounding).
        const TvTile t = tv_tile(tile_of(ti));
        const int T = p.T, Vr = p.V, out_lo = out_lo_s;
        const int nfo = min(C::kFo, T - t.fb * C::kFo);            // output frames of this tile
        const bool inner = t.tq0 >= 0 && t.tq0 + G <= T;           // every staged frame of q lies inside the sample
        const bool active = wv < 2 * V;
        const int v = min(wv >> 1, V - 1);
        const size_t fstride = (size_t)Vr * N;                     // one frame of the output / residual tensors
        int nch_rt = NCH;                                          // rolled: one chunk's residual values live at a time
        asm volatile("" : "+r"(nch_rt));
#pragma unroll 1
        for (int nc = 0; nc < nch_rt; ++nc) {
          const int d = nc * 64 + (wv & 1) * 32 + lane;
          const float yo = __ldg(p.res2 + d), scb = __ldg(p.epi_a + d), shb = __ldg(p.epi_b + d);
          const float bias = p.bias ? __ldg(p.bias + d) : 0.f;
          const float fl = floorf(yo), f = yo - fl, gw = 1.f - f;
          const int idx = min(max((int)fl - out_lo, 0), C::kWo - 1);
          const size_t o0 = (((size_t)t.n * T + (size_t)t.fb * C::kFo) * Vr + (size_t)(t.js * V + v)) * N + d;
          float rv[C::kFo];
          {
            const float* rp = res + o0;                            // requested in front of the drain and its barrier
#pragma unroll
            for (int fo = 0; fo < C::kFo; ++fo) {
              rv[fo] = __ldg(rp);
              if (fo + 1 < nfo) rp += fstride;
            }
          }
          if (warp < 8) {   // TMEM -> staging: lane quarter (warp & 3), column half (warp >> 2)
            const int qd = warp & 3, hf = warp >> 2;
            const uint32_t row = (uint32_t)(qd * 32 + lane);
            const uint32_t rb = sSt + row * (uint32_t)C::kStPitch + (uint32_t)(hf * 128);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              float a[16];
              tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * N + nc * 64 + hf * 32 + h2 * 16), a);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                sts128(rb + (uint32_t)((h2 * 4 + i) * 16), make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]));
            }
            if (nc == NCH - 1) {                                   // last read of this accumulator buffer
              tc_fence_before();
              mbar_arrive(&acc_free[buf]);
            }
          }
          epi_sync();
          if (active) {
            const uint32_t sb = stb + (uint32_t)((idx * V + v) * C::kStPitch);   // staging row of tap 0 of output frame 0
            const int tq = t.tq0 + idx;                            // its frame inside the sample
            auto qtap = [&](int i) -> float {                      // q of staged frame idx + i, zero padded outside [0, T)
              float qv = fmaxf(lds32(sb + (uint32_t)(i * V * C::kStPitch)) + bias, 0.f);
              if (!inner && (unsigned)(tq + i) >= (unsigned)T) qv = 0.f;
              return qv;
            };
            float* op = p.out + o0;
            float a = qtap(0);
#pragma unroll
            for (int fo = 0; fo < C::kFo; ++fo) {
              const float b = qtap(fo + 1);
              float y = fmaf(fmaf(f, b, gw * a), scb, shb);
              if (p.res) y += rv[fo];
              if (p.relu) y = fmaxf(y, 0.f);
              if (fo < nfo) {
                *op = y;
                op += fstride;
              }
              a = b;
            }
          }
          epi_sync();                                              // staging is reused by the next chunk / tile
        }
      } else if constexpr (EPI == EPI_LINEAR) {
        // 32 output channels per step through a [128 x 32] staging tile (128-byte rows, 16-byte chunk ^ (row & 7)):
        // piece = thread + 384*i  ->  row = (thread >> 3) + 48*i, chunk = thread & 7
        const int lrow = et >> 3, lc = et & 7;
        const int nrow = ng * V;
#pragma unroll
        for (int st = 0; st < 2 * NCH; ++st) {
          const int d = st * 32 + lc * 4;                          // requested before the drain and its barrier
          const float4 bias = p.bias ? ldg4(p.bias + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (warp < 8) {   // TMEM -> staging: lane quarter (warp & 3), 16-column half (warp >> 2)
            const int qd = warp & 3, hf = warp >> 2;
            const uint32_t row = (uint32_t)(qd * 32 + lane);
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * N + st * 32 + hf * 16), v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts128(sSt + row * 128u + ((((uint32_t)(hf * 4 + i)) ^ (row & 7u)) << 4), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            if (st == 2 * NCH - 1) {                               // last read of this accumulator buffer
              tc_fence_before();
              mbar_arrive(&acc_free[buf]);
            }
          }
          epi_sync();
          const uint32_t sb = sSt + (uint32_t)lrow * 128u + ((((uint32_t)lc) ^ ((uint32_t)lrow & 7u)) << 4);
          if (p.out_gs <= 1 && !p.accum) {
            float* optr = p.out + (row0 + lrow) * N + d;
#pragma unroll
            for (int i = 0; i < (C::kTileRows + 47) / 48; ++i)
              if (lrow + 48 * i < nrow) {                          // (row & 7) is invariant under +48
                float4 y = lds128(sb + (uint32_t)(i * 48 * 128));
                y.x += bias.x, y.y += bias.y, y.z += bias.z, y.w += bias.w;
                if (p.relu) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
                *(float4*)(optr + (size_t)i * 48 * N) = y;
              }
          } else {   // output frames strided (transposed stride-2 conv) and / or accumulated into an existing gradient
            const long long ogs = p.out_gs > 0 ? p.out_gs : 1;
#pragma unroll
            for (int i = 0; i < (C::kTileRows + 47) / 48; ++i) {
              const int row = lrow + 48 * i;
              if (row < nrow) {
                const int gr = row / V;
                float4* optr = (float4*)(p.out + ((size_t)((g0 + gr) * ogs) * V + (row - gr * V)) * N + d);
                float4 y = lds128(sb + (uint32_t)(i * 48 * 128));
                y.x += bias.x, y.y += bias.y, y.z += bias.z, y.w += bias.w;
                if (p.relu) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
                if (p.accum) {
                  const float4 o = *optr;
                  y.x += o.x, y.y += o.y, y.z += o.z, y.w += o.w;
                }
                *optr = y;
              }
            }
          }
          epi_sync();                                              // staging is reused by the next step / tile
        }
      } else {
      // ROT_RAW keeps per-chunk statistics in registers (compile-time indices: full unroll); ROT_FUSED has no such
      // arrays and, for four chunks, stays rolled so that only one chunk's prefetch is live (measured: 246 -> 178 us
      // at N = 256; two chunks are faster unrolled, 204 vs 246 us)
      constexpr int kNcUnroll = (EPI == EPI_ROT_FUSED && NCH > 1) ? 1 : NCH;
      int nch_rt = NCH;                                            // ROT_FUSED: opaque trip count, the loop must stay
      if constexpr (EPI == EPI_ROT_FUSED) asm volatile("" : "+r"(nch_rt));   // rolled even for two chunks (spills)
#pragma unroll kNcUnroll
      for (int nc = 0; nc < (EPI == EPI_ROT_FUSED ? nch_rt : NCH); ++nc) {
        if constexpr (EPI == EPI_ROT_FUSED)
          if (nc > 0) prefetch(nc);
        // requested in front of the drain and its barrier, used after them
        const float bias = p.bias ? __ldg(p.bias + nc * 64 + (wv & 1) * 32 + lane) : 0.f;
        if (warp < 8) {   // TMEM -> staging: lane quarter (warp & 3), column half (warp >> 2)
          const int qd = warp & 3, hf = warp >> 2;
          const uint32_t row = (uint32_t)(qd * 32 + lane);
          const uint32_t rb = sSt + row * (uint32_t)C::kStPitch + (uint32_t)(hf * 128);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * N + nc * 64 + hf * 32 + h2 * 16), v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts128(rb + (uint32_t)((h2 * 4 + i) * 16), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          }
          if (nc == NCH - 1) {                                     // last read of this accumulator buffer
            tc_fence_before();
            mbar_arrive(&acc_free[buf]);
          }
        }
        epi_sync();
        {
          // z[(g,v), d] = y[(g, (v-d) % V), d] + bias: lane <-> channel d, the warp owns (joint v, half) pairs.
          // Staging row r = g*V + u: address = stb + r * kStPitch (per-pair base + compile-time group offset)
          const int half = wv & 1;                                 // pairs k*12 + warp keep the parity of the warp
          const int d = nc * 64 + half * 32 + lane, dm = d % V;
          float* obase = p.out + row0 * N + d;
          auto tile_pass = [&](auto full_tag) {
            constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
            for (int k = 0; k < C::kJ; ++k) {
              const int v = (k * kEpiWarps + wv) >> 1;
              if (v < V) {
                int u = v - dm;
                if (u < 0) u += V;
                float* optr = obase + v * N;
                const uint32_t sb = stb + (uint32_t)u * (uint32_t)C::kStPitch;
                float sc = 0.f, sh = 0.f, rv[G];
                if (EPI == EPI_ROT_FUSED) {
                  sc = scp[EPI == EPI_ROT_FUSED ? k : 0];
                  sh = shp[EPI == EPI_ROT_FUSED ? k : 0];
#pragma unroll
                  for (int g = 0; g < G; ++g) rv[g] = rvp[EPI == EPI_ROT_FUSED ? k : 0][g];
                }
#pragma unroll
                for (int g = 0; g < G; ++g)
                  if (kFull || g < ng) {
                    float z = lds32(sb + (uint32_t)(g * V * C::kStPitch)) + bias;
                    if (EPI == EPI_ROT_RAW) {
                      s1[nc * C::kJ + k] += z;
                      s2[nc * C::kJ + k] = fmaf(z, z, s2[nc * C::kJ + k]);
                    } else {
                      z = fmaf(z, sc, sh) + rsel * rv[g];
                      if (p.relu) z = fmaxf(z, 0.f);
                    }
                    optr[g * V * N] = z;
                  }
              }
            }
          };
          if (ng == G) tile_pass(std::true_type{});
          else tile_pass(std::false_type{});
        }
        epi_sync();                                                // staging is reused by the next chunk / tile
      }
      }
    }
    if (EPI == EPI_ROT_RAW) {   // flush the per-(v,d) batch statistics
#pragma unroll
      for (int nc = 0; nc < NCH; ++nc)
#pragma unroll
        for (int k = 0; k < C::kJ; ++k) {
          const int pi = k * kEpiWarps + warp;
          if (pi < C::kPairs) {
            const size_t f = (size_t)(pi >> 1) * N + nc * 64 + (pi & 1) * 32 + lane;
            atomicAdd(p.stats + 2 * f, (double)s1[nc * C::kJ + k]);
            atomicAdd(p.stats + 2 * f + 1, (double)s2[nc * C::kJ + k]);
          }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}

template <int PRO, int EPI, int V, int K, int N, bool P3>
static int launch_p(const SgcnRowGemm& p, cudaStream_t s) {
  using C = Cfg<PRO, EPI, V, K, N, P3>;
  auto kern = fused_gemm_kernel<PRO, EPI, V, K, N, P3>;
  static std::atomic<unsigned long long> configured{0};           // one bit per device (the attribute is per device)
  if (needs_configure(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem);
    if (e != cudaSuccess) return set_cuda_error("fused_gemm smem attribute", e);
    mark_configured(configured);
  }
  constexpr bool TV = EPI == EPI_TSHIFT;
  if (TV && (p.T < 1 || p.groups % p.T != 0 || p.V % V != 0)) return set_error("temporal unit (fused): bad frame / joint geometry");
  const long long ntiles = TV ? (p.groups / p.T) * ((p.T + C::kFo - 1) / C::kFo) * (p.V / V) : (p.groups + C::G - 1) / C::G;
  if (ntiles == 0) return 0;
  long long grid = tile_ctas();
  if (grid > ntiles) grid = ntiles;
  // tensor maps of the activation inputs (tensormap.h): raw [rows x 64] boxes, or whole operand blocks for PRO_PLAIN
  alignas(64) CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  if constexpr (TV) {
    if (int rc = make_frames_map(&tm0, p.in0, p.groups, p.V, K, 64, V, C::G + C::kWin)) return rc;
    tm1 = tm0;
  } else if constexpr (PRO != PRO_PLAIN) {
    if (int rc = make_rows_map(&tm0, p.in0, p.groups * V, K, 64, C::kRawRows)) return rc;
    tm1 = tm0;
  } else if constexpr (!P3) {
    const int k0 = p.k0 > 0 ? p.k0 : K;
    if (int rc = make_groups_map_sw128(&tm0, p.in0, p.groups, p.in0_gs, V, k0, C::G)) return rc;
    tm1 = tm0;
    if (k0 < K)
      if (int rc = make_groups_map_sw128(&tm1, p.in1, p.groups, p.in1_gs, V, K - k0, C::G)) return rc;
  } else {
    tm1 = tm0;
  }
  kern<<<(unsigned)grid, kThreads, C::kSmem, s>>>(p, next_direction(), tm0, tm1);
  return check_launch("fused_gemm_kernel");
}

template <int PRO, int EPI, int V, int K, int N>
static int launch(const SgcnRowGemm& p, cudaStream_t s) {
  if (p.prec == SGCN_PREC_FP32) return launch_p<PRO, EPI, V, K, N, true>(p, s);
  if (p.prec != SGCN_PREC_TF32) return set_error("sgcn_rowgemm: unknown precision");
  return launch_p<PRO, EPI, V, K, N, false>(p, s);
}

}  // namespace fg
}  // namespace sgcn
