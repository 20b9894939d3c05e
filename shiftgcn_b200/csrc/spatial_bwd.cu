// spatial_bwd.cu -- backward-data pass of the spatial unit (autograd of model/shift_gcn.py:123-141), one kernel:
//
//     dz  = alpha*gh + beta*z + gamma                  BatchNorm1d backward, folded into three per-(v,d) tables
//     A[(g,u), d] = dz[g, (u+d) % V, d]                inverse of the shift_out gather
//     dxm = A * W^T                                    tcgen05.mma kind::tf32, fp32 accumulators in TMEM
//     gx[(g,w), c] = dxm[(g,u), c] * maskmul[u, c]     u = (w-c) % V: inverse of the shift_in gather
//                    + gh[(g,w), c]                    gradient of the unit's identity `down` branch   (optional)
//                    + [y > 0] * g_y[(g,w), c]         gradient of the block residual                   (optional)
//     dMask[u, c] += dxm[(g,u), c] * x[(g,w), c]       raw Feature_Mask gradient
//
// Warp-specialised and persistent (one CTA per SM, tiles of G whole (n,t) groups = 125 / 99 rows):
//   epilogue warps 0..12   TMEM -> swizzled smem staging (warps 0..7) -> each thread owns ONE output slot
//                          (joint w, 4 channels) and walks the tile's groups: gh / g_y / x of one group arrive through a
//                          TMA ring (warp 22 issues {64 channels, V rows} boxes, mbarrier complete_tx), four scalar gathers
//                          from the staging tile (the rotation), 128-bit store of gx, dMask partial sums in registers
//                          for the whole kernel.  (fp32-accurate mode: register-staged 128-bit loads instead of the ring)
//   warp 13                issues tcgen05.mma, streams weight chunks (cp.async.bulk) when W does not fit
//   builder warps 14..21   own SOURCE slots (joint sv, 4 channels): 128-bit loads of gh / z, dz, scatter of the
//                          four channels to their rotated rows of the K-major SWIZZLE_128B operand chunk
// Pipelines: operand chunks (2 buffers, full/free mbarriers), TMEM accumulators (2 buffers, full/free), so the
// builders run up to two 64-channel chunks ahead of the tensor core and the epilogue of tile i overlaps tile i+1.
#include "capi_internal.h"
#include "common.cuh"
#include "rowgemm.h"
#include "tensormap.h"
#include <string.h>

namespace sgcn {

namespace sb {

#ifndef SGCN_SB_EPI_WARPS
#define SGCN_SB_EPI_WARPS 13
#endif
#ifndef SGCN_SB_BLD_WARPS
#define SGCN_SB_BLD_WARPS 8
#endif
constexpr int kEpiWarps = SGCN_SB_EPI_WARPS, kBldWarps = SGCN_SB_BLD_WARPS;
static_assert(kEpiWarps >= 8, "warps 0..7 read TMEM");
constexpr int kEpiThreads = kEpiWarps * 32, kBldThreads = kBldWarps * 32;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kLdWarp = kEpiWarps + 1 + kBldWarps;               // epilogue-input loader (TMA), one lane
constexpr int kLdWarp2 = kLdWarp + 1;                            // builder-input loader (TMA), one lane
constexpr int kThreads = (kEpiWarps + 1 + kBldWarps + 2) * 32;   // 768 (13 + 1 + 8 + 2 warps)
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// v[t] <- v[(t + r) & 3].  The rotated row scatters / gathers move ONE word of a 16-byte chunk per thread and step: with
// the same word index in every lane a warp touches only 8 of the 32 banks (5.0 wavefronts per access on average);
// starting each lane at word (lane & 3) spreads the words over all banks (1.9 wavefronts).
#ifndef SGCN_SB_ROT
#define SGCN_SB_ROT 0             // measured: fewer bank conflicts but SLOWER (C=128: 266 -> 312 us) -- the kernel is bound by issued instructions, not by shared-memory wavefronts
#endif
__device__ __forceinline__ uint32_t rot_lane(int lane) { return SGCN_SB_ROT ? (uint32_t)lane & 3u : 0u; }
__device__ __forceinline__ void rot4(float (&v)[4], uint32_t r) {
  const bool r1 = r & 1u, r2 = r & 2u;
  const float a0 = r1 ? v[1] : v[0], a1 = r1 ? v[2] : v[1], a2 = r1 ? v[3] : v[2], a3 = r1 ? v[0] : v[3];
  v[0] = r2 ? a2 : a0, v[1] = r2 ? a3 : a1, v[2] = r2 ? a0 : a2, v[3] = r2 ? a1 : a3;
}

// staging tile [128 rows][64 cols] fp32, 16-byte chunks XOR-ed with (row & 15): conflict-free for "thread = row"
// 128-bit stores and (mostly) for the rotated scalar gathers
__device__ __forceinline__ uint32_t stage_off(uint32_t row, uint32_t c4, uint32_t j) {
  return row * 256u + ((c4 ^ (row & 15u)) << 4) + (j << 2);
}

// P3 = fp32-accurate mode (3xTF32, SgcnRowGemm::prec): operand chunks hold {TF32 head, TF32 tail}, the weight image is
// {head image, tail image} and every chunk runs Ah*Wh, Al*Wh, Ah*Wl.
template <int V, int K, int N, bool P3 = false>
struct Cfg {
  static constexpr int G = 128 / V;
  static constexpr int KC = K / 64, NCH = N / 64;
  static constexpr int kImgBytes = K * N * 4;
  static constexpr int kOpBytes = (P3 ? 2 : 1) * kChunkBytes;
  // weights stay resident only up to 32 KiB: at 128 -> 128 channels streaming the image in [N x 64] chunks (from L2) frees
  // 32 KiB for a fourth input-ring slot, which is worth more (C=128: 274 -> 258 us)
#ifndef SGCN_SB_WRES_MAX
#define SGCN_SB_WRES_MAX 32768
#endif
  static constexpr bool kWRes = ((P3 ? 2 : 1) * kImgBytes) <= (P3 ? 65536 : SGCN_SB_WRES_MAX);
  static constexpr int kWChunk = N * 256;                          // one streamed [N x 64] weight chunk
  static constexpr int kWLoads = P3 ? 2 : 1;                       // streamed loads per operand chunk (head, tail)
  static constexpr int kWBytes = kWRes ? (P3 ? 2 : 1) * kImgBytes : kWChunk;
  static constexpr int kSlots = V * 16;                            // (joint, 4-channel group) slots of a 64-wide chunk
  static constexpr int kEpiRounds = (kSlots + kEpiThreads - 1) / kEpiThreads;
  static constexpr int kBldRounds = (kSlots + kBldThreads - 1) / kBldThreads;
  // operand stages: as many as fit (at most 4), so that the register-staged builder loads run several chunks ahead of
  // the tensor core and the epilogue
#ifndef SGCN_SB_MAX_STAGES
#define SGCN_SB_MAX_STAGES 2      // measured: 3-4 stages are SLOWER (269 vs 250 us at C=64, 382 vs 268 at C=128) -- the builders' loads then delay the epilogue's, which are on the critical path
#endif
  static constexpr int kFit = (232448 - 512 - 1024 - 64 - kWBytes - kChunkBytes) / kOpBytes;
  static constexpr int kOpStages = kFit < 2 ? 2 : (kFit > SGCN_SB_MAX_STAGES ? SGCN_SB_MAX_STAGES : kFit);
  static constexpr int kCore = 1024 + kWBytes + kOpStages * kOpBytes + kChunkBytes + 64;
  // Epilogue inputs (gh, g_y, x of one (n,t) group x 64 channels) arrive through a TMA ring instead of register-staged
  // loads: the epilogue used to issue a batch of 128-bit loads, wait a full memory latency and only then consume it, twice
  // per chunk; the ring keeps kInStages groups in flight while earlier ones are consumed.
#ifndef SGCN_SB_TMA_IN
#define SGCN_SB_TMA_IN 1
#endif
  static constexpr int kInSlot = 3 * V * 256;                      // {res, res2, xin} x [V rows x 64 channels]
  static constexpr int kInFit = (232448 - 512 - kCore) / kInSlot;
  static constexpr bool kTmaIn = SGCN_SB_TMA_IN && !P3 && kInFit >= 2;
#ifndef SGCN_SB_IN_MAX
#define SGCN_SB_IN_MAX 4          // measured at C=64: 2 slots 280 us, 4 slots 232 us, 6 slots 250 us (more requests in flight delay the builders' loads)
#endif
  static constexpr int kInStages = !kTmaIn ? 0 : (kInFit > SGCN_SB_IN_MAX ? SGCN_SB_IN_MAX : kInFit);
  // Builder inputs (gh, z of one group x 64 channels) through a second TMA ring when it fits next to the first: the
  // register-staged builders need two load rounds per chunk (400 slots on 256 threads, 10 x 16 bytes each), i.e. two
  // exposed memory latencies per chunk, and more builder threads do not fit the register file.
#ifndef SGCN_SB_TMA_BLD
#define SGCN_SB_TMA_BLD 0         // measured: SLOWER than the register-staged builders (C=64: 294 vs 232 us, C=128: 289 vs 266): the raw copy costs a second pass through shared memory and two slots of 12.8 KB keep less in flight than 10 x 16 B per thread
#endif
#ifndef SGCN_SB_BLD_MAX
#define SGCN_SB_BLD_MAX 4
#endif
  static constexpr int kBinSlot = 2 * V * 256;                     // {gh, z} x [V rows x 64 channels]
  static constexpr int kBinFit = (232448 - 512 - kCore - kInStages * kInSlot) / kBinSlot;
  static constexpr bool kTmaBld = SGCN_SB_TMA_BLD && kTmaIn && kBinFit >= 2;
  static constexpr int kBinStages = !kTmaBld ? 0 : (kBinFit > SGCN_SB_BLD_MAX ? SGCN_SB_BLD_MAX : kBinFit);
  static constexpr size_t kSmem = kCore + kInStages * kInSlot + kBinStages * kBinSlot;
  static_assert(kSmem <= 232448 - 512, "shared memory budget");
};

template <int V, int K, int N, bool P3>
__global__ void __launch_bounds__(kThreads, 1) spatial_bwd_kernel(const SgcnRowGemm p, const int rev,
                                                                  const __grid_constant__ CUtensorMap tm_res,
                                                                  const __grid_constant__ CUtensorMap tm_res2,
                                                                  const __grid_constant__ CUtensorMap tm_x,
                                                                  const __grid_constant__ CUtensorMap tm_gh,
                                                                  const __grid_constant__ CUtensorMap tm_z) {
  using C = Cfg<V, K, N, P3>;
  constexpr int kOpBytes = C::kOpBytes;
  constexpr int G = C::G, KC = C::KC, NCH = C::NCH, OS = C::kOpStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KiB alignment as an OFFSET from the __shared__ array: a pointer rebuilt from an integer loses its address space
  // and every access through it becomes a generic LD/ST with 64-bit address arithmetic
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sOp = sW + C::kWBytes;                                  // OS operand chunks
  uint8_t* sSt = sOp + OS * kOpBytes;                              // epilogue staging
  uint8_t* sIn = sSt + kChunkBytes;                                // epilogue-input ring (kTmaIn)
  uint8_t* sBin = sIn + C::kInStages * C::kInSlot;                 // builder-input ring (kTmaBld)
  constexpr int IS = C::kInStages > 0 ? C::kInStages : 1;
  constexpr int BS = C::kBinStages > 0 ? C::kBinStages : 1;
  __shared__ uint64_t op_full[4], op_free[4], acc_full[2], acc_free[2], w_full, w_free, in_full[8], in_free[8], bin_full[4],
      bin_free[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&op_full[i], kBldThreads);
      mbar_init(&op_free[i], 1);
      mbar_init(&in_full[i], 1), mbar_init(&in_full[i + 4], 1);
      mbar_init(&in_free[i], kEpiWarps), mbar_init(&in_free[i + 4], kEpiWarps);
      mbar_init(&bin_full[i], 1);
      mbar_init(&bin_free[i], kBldWarps);
    }
    for (int i = 0; i < 2; ++i) {
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
      // streamed weights: ONE [N x 64] buffer, (re)loaded kWLoads times per operand chunk -- load L = kWLoads * q + part
      // is chunk kc of the head (part 0) or tail (part 1) image; it is requested as soon as the MMAs that read load
      // L - 1 have completed
      const long long total_loads = C::kWLoads * total_chunks;
      auto load_w = [&](long long L) {
        const int kc = (int)((L / C::kWLoads) % KC), part = (int)(L % C::kWLoads);
        mbar_expect_tx(&w_full, C::kWBytes);
        bulk_load(sW, (const uint8_t*)p.wimg + (size_t)part * C::kImgBytes + (size_t)kc * C::kWChunk, C::kWBytes, &w_full);
      };
      if (!C::kWRes && total_chunks > 0) load_w(0);
      long long q = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const int buf = (int)(ti & 1);
        if (ti >= 2) mbar_wait(&acc_free[buf], (uint32_t)(((ti >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * N);
        for (int kc = 0; kc < KC; ++kc, ++q) {
          const int s = (int)(q % OS);
          mbar_wait(&op_full[s], (uint32_t)((q / OS) & 1));
          tc_fence_after();
          const uint32_t a0 = smem_u32(sOp) + (uint32_t)s * kOpBytes;
#pragma unroll
          for (int part = 0; part < (P3 ? 2 : 1); ++part) {
            const long long L = C::kWLoads * q + part;
            if (!C::kWRes) {
              mbar_wait(&w_full, (uint32_t)(L & 1));
              tc_fence_after();
            }
            const uint32_t w0 = smem_u32(sW) + (C::kWRes ? (uint32_t)(part * C::kImgBytes) + (uint32_t)kc * (uint32_t)C::kWChunk : 0u);
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const uint32_t blk = k8 >> 2, sub = k8 & 3;
              umma_tf32(acc, umma_desc(a0 + blk * kBlockBytes + sub * 32, 16, 1024),
                        umma_desc(w0 + blk * (uint32_t)N * 128u + sub * 32, 16, 1024), idesc, (kc | k8 | part) ? 1u : 0u);
            }
            if (P3 && part == 0) {                                  // tail of the activations against the weight heads
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8) {
                const uint32_t blk = k8 >> 2, sub = k8 & 3;
                umma_tf32(acc, umma_desc(a0 + kChunkBytes + blk * kBlockBytes + sub * 32, 16, 1024),
                          umma_desc(w0 + blk * (uint32_t)N * 128u + sub * 32, 16, 1024), idesc, 1u);
              }
            }
            if (!C::kWRes) {
              tc_commit(&w_free);
              if (L + 1 < total_loads) {                            // single weight buffer: reload once these MMAs have read it
                mbar_wait(&w_free, (uint32_t)(L & 1));
                load_w(L + 1);
              }
            }
          }
          tc_commit(&op_free[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp == kLdWarp) {
    // ================================================================================ epilogue-input loader
    if (C::kTmaIn && lane == 0) {
      tma_prefetch_map(&tm_res), tma_prefetch_map(&tm_res2), tma_prefetch_map(&tm_x);
      const bool has_res = p.res != nullptr, has_res2 = p.res2 != nullptr;
      const uint32_t bytes = (uint32_t)(V * 256) * (1u + (has_res ? 1u : 0u) + (has_res2 ? 1u : 0u));
      int s = 0;                                                   // ring slot and the parity of its NEXT free phase
      uint32_t ph = 1;                                             // (a fresh mbarrier passes a wait on parity 1)
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const long long tile = rev ? ntiles - 1 - (blockIdx.x + ti * gridDim.x) : blockIdx.x + ti * gridDim.x;
        const long long g0 = tile * G;
        const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
        for (int nc = 0; nc < NCH; ++nc)
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&in_free[s], ph);
            const uint32_t dst = smem_u32(sIn) + (uint32_t)s * C::kInSlot, bar = smem_u32(&in_full[s]);
            const int row = (int)((g0 + g) * V);
            mbar_expect_tx(&in_full[s], bytes);
            tma_load_2d(dst + 2u * V * 256u, &tm_x, nc * 64, row, bar);
            if (has_res) tma_load_2d(dst, &tm_res, nc * 64, row, bar);
            if (has_res2) tma_load_2d(dst + V * 256u, &tm_res2, nc * 64, row, bar);
            if (++s == IS) s = 0, ph ^= 1u;
          }
      }
    }
    __syncwarp();
  } else if (warp == kLdWarp2) {
    // ================================================================================ builder-input loader
    if (C::kTmaBld && lane == 0) {
      tma_prefetch_map(&tm_gh), tma_prefetch_map(&tm_z);
      int s = 0;
      uint32_t ph = 1;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const long long tile = rev ? ntiles - 1 - (blockIdx.x + ti * gridDim.x) : blockIdx.x + ti * gridDim.x;
        const long long g0 = tile * G;
        const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
        for (int kc = 0; kc < KC; ++kc)
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&bin_free[s], ph);
            const uint32_t dst = smem_u32(sBin) + (uint32_t)s * C::kBinSlot, bar = smem_u32(&bin_full[s]);
            const int row = (int)((g0 + g) * V);
            mbar_expect_tx(&bin_full[s], 2u * V * 256u);
            tma_load_2d(dst, &tm_gh, kc * 64, row, bar);
            tma_load_2d(dst + V * 256u, &tm_z, kc * 64, row, bar);
            if (++s == BS) s = 0, ph ^= 1u;
          }
      }
    }
    __syncwarp();
  } else if (warp > kMmaWarp && warp < kLdWarp) {
    // ================================================================================ builders
    const int bt = tid - (kMmaWarp + 1) * 32;
    int bin_s = 0;                                                 // builder-input ring slot / parity (kTmaBld)
    uint32_t bin_ph = 0;
    (void)bin_s, (void)bin_ph;
    long long q = 0;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = rev ? ntiles - 1 - (blockIdx.x + ti * gridDim.x) : blockIdx.x + ti * gridDim.x;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const size_t row0 = (size_t)g0 * V;
      for (int kc = 0; kc < KC; ++kc, ++q) {
        const int s = (int)(q % OS);
        if (q >= OS) mbar_wait_relaxed(&op_free[s], (uint32_t)(((q / OS) - 1) & 1));
        uint8_t* op = sOp + (size_t)s * kOpBytes;
        if constexpr (C::kTmaBld) {
          // ---- inputs from the TMA ring, one (n,t) group per slot {gh, z} x [V x 64]: slot (joint sv, 4 channels c4)
          //      sits at byte (sv*16 + c4)*16 of each
          float4 al[C::kBldRounds], be[C::kBldRounds], ga[C::kBldRounds];
          uint32_t dst0[C::kBldRounds][4];                          // operand address of (row u_j, chunk, j) for group 0
#pragma unroll
          for (int rd = 0; rd < C::kBldRounds; ++rd) {
            const int slot = bt + rd * kBldThreads;
            const int sv = (slot < C::kSlots ? slot : 0) >> 4, c4 = slot & 15;
            const int d = kc * 64 + c4 * 4;
            al[rd] = ldg4(p.pro_a + sv * K + d), be[rd] = ldg4(p.pro_b + sv * K + d), ga[rd] = ldg4(p.pro_c + sv * K + d);
            int u0 = sv - d % V;
            if (u0 < 0) u0 += V;
#pragma unroll
            for (int t = 0; t < 4; ++t) {                           // step t moves word jj = (t + lane) & 3 (rot4)
              const int jj = (t + (int)rot_lane(lane)) & 3;
              int u = u0 - jj;
              if (u < 0) u += V;
              dst0[rd][t] = smem_u32(op) + (uint32_t)(c4 >> 3) * kBlockBytes + (uint32_t)u * 128u + (uint32_t)jj * 4u;
              dst0[rd][t] |= (uint32_t)u << 24;                     // row rides in the top byte (shared addresses are < 2^24)
            }
          }
          for (int g = 0; g < ng; ++g) {
            mbar_wait_relaxed(&bin_full[bin_s], bin_ph);
            const uint32_t in = smem_u32(sBin) + (uint32_t)bin_s * C::kBinSlot;
#pragma unroll
            for (int rd = 0; rd < C::kBldRounds; ++rd) {
              const int slot = bt + rd * kBldThreads;
              if (slot < C::kSlots) {
                const float4 gv = lds128(in + (uint32_t)slot * 16u), zv = lds128(in + V * 256u + (uint32_t)slot * 16u);
                float dz[4] = {fmaf(al[rd].x, gv.x, fmaf(be[rd].x, zv.x, ga[rd].x)), fmaf(al[rd].y, gv.y, fmaf(be[rd].y, zv.y, ga[rd].y)),
                               fmaf(al[rd].z, gv.z, fmaf(be[rd].z, zv.z, ga[rd].z)), fmaf(al[rd].w, gv.w, fmaf(be[rd].w, zv.w, ga[rd].w))};
                rot4(dz, rot_lane(lane));
                const uint32_t cc = (uint32_t)slot & 7u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t row = (dst0[rd][j] >> 24) + (uint32_t)(g * V);
                  const uint32_t dst = (dst0[rd][j] & 0xFFFFFFu) + (uint32_t)(g * V) * 128u + (((cc ^ row) & 7u) << 4);
                  sts32(dst, tf32_half_ulp(dz[j]));
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bin_free[bin_s]);
            if (++bin_s == BS) bin_s = 0, bin_ph ^= 1u;
          }
        } else {
#pragma unroll
        for (int rd = 0; rd < C::kBldRounds; ++rd) {
          const int slot = bt + rd * kBldThreads;
          if (slot < C::kSlots) {
            const int sv = slot >> 4, c4 = slot & 15;
            const int d = kc * 64 + c4 * 4;
            const float4 al = ldg4(p.pro_a + sv * K + d), be = ldg4(p.pro_b + sv * K + d), ga = ldg4(p.pro_c + sv * K + d);
            const size_t o = (row0 + sv) * K + d;
            float4 gv[G], zv[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const size_t og = o + (size_t)min(g, ng - 1) * V * K;
              gv[g] = ldg4(p.in0 + og);
              zv[g] = ldg4(p.in1 + og);
            }
            // destination rows u_j = (sv - d - j) mod V; byte offset of (row, channel) in the K-major swizzled chunk:
            // row * 128 + ((chunk ^ (row & 7)) << 4) -- 32-bit shared-window addresses, four instructions per store
            int uj[4];                                              // step t moves word jj = (t + lane) & 3 (rot4)
            uint32_t wj[4];
            int u0 = sv - d % V;
            if (u0 < 0) u0 += V;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int jj = (t + (int)rot_lane(lane)) & 3;
              uj[t] = u0 - jj;
              if (uj[t] < 0) uj[t] += V;
              wj[t] = (uint32_t)jj * 4u;
            }
            const uint32_t cc = (uint32_t)(c4 & 7);
            const uint32_t op32 = smem_u32(op) + (uint32_t)(c4 >> 3) * kBlockBytes;
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (g < ng) {
                float dz[4] = {fmaf(al.x, gv[g].x, fmaf(be.x, zv[g].x, ga.x)), fmaf(al.y, gv[g].y, fmaf(be.y, zv[g].y, ga.y)),
                               fmaf(al.z, gv[g].z, fmaf(be.z, zv[g].z, ga.z)), fmaf(al.w, gv[g].w, fmaf(be.w, zv[g].w, ga.w))};
                rot4(dz, rot_lane(lane));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t row = (uint32_t)(g * V + uj[j]);
                  const uint32_t dst = op32 + row * 128u + (((cc ^ row) & 7u) << 4) + wj[j];
                  if constexpr (P3) {
                    float hi, lo;
                    split_tf32(dz[j], hi, lo);
                    sts32(dst, hi);
                    sts32(dst + kChunkBytes, lo);
                  } else {
                    sts32(dst, tf32_half_ulp(dz[j]));
                  }
                }
              }
          }
        }
        }
        fence_proxy_async();
        mbar_arrive(&op_full[s]);
      }
    }
  } else {
    // ================================================================================ epilogue
    const int et = tid;                                            // 0 .. kEpiThreads-1
    const float rsel = p.res ? 1.f : 0.f, r2sel = p.res2 ? 1.f : 0.f;
    const float* res = p.res ? p.res : p.xin;                      // absent streams alias x (loads stay branch free)
    const float* res2 = p.res2 ? p.res2 : p.xin;
    const float* res2m = (p.res2 && p.res2m) ? p.res2m : p.xin;
    const bool r2all = p.res2 && !p.res2m;                         // g_y arrives with its ReLU mask already applied
    const bool premask = p.relu != 0;                              // hand gx * [x > 0] to the unit that produced x
    float dm[NCH][C::kEpiRounds][4];
#pragma unroll
    for (int a = 0; a < NCH; ++a)
#pragma unroll
      for (int b = 0; b < C::kEpiRounds; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) dm[a][b][c] = 0.f;

    int in_s = 0;                                                  // input-ring slot and parity of its next full phase (kTmaIn)
    uint32_t in_ph = 0;
    (void)in_s, (void)in_ph;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = rev ? ntiles - 1 - (blockIdx.x + ti * gridDim.x) : blockIdx.x + ti * gridDim.x;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const size_t row0 = (size_t)g0 * V;
      const int buf = (int)(ti & 1);
      // Only the TMEM-reading warps wait for the accumulator: they are the ones whose arrival on acc_free lets the
      // issuer reuse the buffer, so the barrier can never run two phases ahead of a waiter (a warp that waited
      // without gating acc_free could miss a whole phase and spin forever).  Warps 8..12 are ordered by epi_sync.
      if (warp < 8) {   // (every epilogue warp when kEpiWarps == 8)
        mbar_wait_relaxed(&acc_full[buf], (uint32_t)((ti >> 1) & 1));
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
        if constexpr (C::kTmaIn) {
          // ---- inputs from the TMA ring, one (n,t) group per slot: slot layout {res, res2, xin} x [V x 64] floats,
          //      thread (joint w, 4 channels c4) reads 16 bytes at (w*16 + c4)*16 of each -- consecutive lanes,
          //      consecutive 16-byte pieces
          constexpr bool kHoist = C::kEpiRounds <= 2;             // keep the rotation rows / mask multipliers in registers
          constexpr int HR = kHoist ? C::kEpiRounds : 1;
          int u[HR][4];
          float mm[HR][4];
          // uu[t]: rotation row of word jj = (t + lane) & 3 (gather step t, see rot4);  m[j]: mask multiplier of word j
          auto rot_rows = [&](int rd, int (&uu)[4], float (&m)[4]) {
            const int slot = et + rd * kEpiThreads;
            const int w = (slot < C::kSlots ? slot : 0) >> 4, c = nc * 64 + (slot & 15) * 4;
            int u0 = w - c % V;
            if (u0 < 0) u0 += V;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int jj = (t + (int)rot_lane(lane)) & 3;
              uu[t] = u0 - jj;
              if (uu[t] < 0) uu[t] += V;
              int un = u0 - t;
              if (un < 0) un += V;
              m[t] = __ldg(p.epi_a + un * N + c + t);
            }
          };
          if constexpr (kHoist) {
#pragma unroll
            for (int rd = 0; rd < C::kEpiRounds; ++rd) rot_rows(rd, u[rd], mm[rd]);
          }
          for (int g = 0; g < ng; ++g) {
            const int s = in_s;
            float4 yv[C::kEpiRounds];
            if (!r2all && p.res2)                                  // the unmasked-g_y case reads y directly (rare: last unit)
#pragma unroll
              for (int rd = 0; rd < C::kEpiRounds; ++rd) {
                const int slot = et + rd * kEpiThreads;
                if (slot < C::kSlots) yv[rd] = ldg4(res2m + (row0 + (size_t)g * V) * N + (size_t)(slot >> 4) * N + nc * 64 + (slot & 15) * 4);
              }
            mbar_wait_relaxed(&in_full[s], in_ph);
            const uint8_t* in = sIn + (size_t)s * C::kInSlot;
#pragma unroll
            for (int rd = 0; rd < C::kEpiRounds; ++rd) {
              const int slot = et + rd * kEpiThreads;
              if (slot < C::kSlots) {
                const int w = slot >> 4, c4 = slot & 15;
                const float4 xv = *(const float4*)(in + 2 * V * 256 + slot * 16);
                float4 rv = make_float4(0.f, 0.f, 0.f, 0.f), gy = rv;
                if (p.res) rv = *(const float4*)(in + slot * 16);
                if (p.res2) gy = *(const float4*)(in + V * 256 + slot * 16);
                if (!r2all && p.res2) {
                  gy.x = yv[rd].x > 0.f ? gy.x : 0.f;
                  gy.y = yv[rd].y > 0.f ? gy.y : 0.f;
                  gy.z = yv[rd].z > 0.f ? gy.z : 0.f;
                  gy.w = yv[rd].w > 0.f ? gy.w : 0.f;
                }
                if constexpr (!kHoist) rot_rows(rd, u[0], mm[0]);
                const int (&uu)[4] = u[kHoist ? rd : 0];
                const float (&m)[4] = mm[kHoist ? rd : 0];
                float val[4];                                       // step t gathers word (t + lane) & 3 ...
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  val[t] = *(const float*)(sSt + stage_off((uint32_t)(g * V + uu[t]), (uint32_t)c4, (t + rot_lane(lane)) & 3u));
                rot4(val, 4u - rot_lane(lane));              // ... back to val[j] = dxm of word j
                float4 out;
                out.x = fmaf(val[0], m[0], rv.x + gy.x);
                out.y = fmaf(val[1], m[1], rv.y + gy.y);
                out.z = fmaf(val[2], m[2], rv.z + gy.z);
                out.w = fmaf(val[3], m[3], rv.w + gy.w);
                if (premask) {
                  out.x = xv.x > 0.f ? out.x : 0.f;
                  out.y = xv.y > 0.f ? out.y : 0.f;
                  out.z = xv.z > 0.f ? out.z : 0.f;
                  out.w = xv.w > 0.f ? out.w : 0.f;
                }
                *(float4*)(p.out + (row0 + (size_t)g * V + w) * N + nc * 64 + c4 * 4) = out;
                dm[nc][rd][0] = fmaf(val[0], xv.x, dm[nc][rd][0]);
                dm[nc][rd][1] = fmaf(val[1], xv.y, dm[nc][rd][1]);
                dm[nc][rd][2] = fmaf(val[2], xv.z, dm[nc][rd][2]);
                dm[nc][rd][3] = fmaf(val[3], xv.w, dm[nc][rd][3]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&in_free[s]);
            if (++in_s == IS) in_s = 0, in_ph ^= 1u;
          }
        } else {
#pragma unroll
        for (int rd = 0; rd < C::kEpiRounds; ++rd) {
          const int slot = et + rd * kEpiThreads;
          if (slot < C::kSlots) {
            const int w = slot >> 4, c4 = slot & 15;
            const int c = nc * 64 + c4 * 4;
            int u[4];
            u[0] = w - c % V;
            if (u[0] < 0) u[0] += V;
#pragma unroll
            for (int j = 1; j < 4; ++j) {
              u[j] = u[j - 1] - 1;
              if (u[j] < 0) u[j] += V;
            }
            float mm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) mm[j] = __ldg(p.epi_a + u[j] * N + c + j);
            const size_t o = (row0 + w) * N + c;
            // two load batches: groups [0, GA) and [GA, G)
            constexpr int GA = (G + 1) / 2;                        // 3 of 5, 2 of 3
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              constexpr int GB = G - GA;
              const int gbeg = half ? GA : 0, gcnt = half ? GB : GA;
              float4 rv[GA], gy[GA], yv[GA], xv[GA];
#pragma unroll
              for (int i = 0; i < GA; ++i)
                if (i < gcnt) {
                  const size_t og = o + (size_t)min(gbeg + i, ng - 1) * V * N;
                  rv[i] = ldg4(res + og);
                  gy[i] = ldg4(res2 + og);
                  yv[i] = ldg4(res2m + og);
                  xv[i] = ldg4(p.xin + og);
                }
#pragma unroll
              for (int i = 0; i < GA; ++i)
                if (i < gcnt && gbeg + i < ng) {
                  const int g = gbeg + i;
                  float val[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) val[j] = *(const float*)(sSt + stage_off((uint32_t)(g * V + u[j]), (uint32_t)c4, (uint32_t)j));
                  float4 out;
                  out.x = fmaf(val[0], mm[0], fmaf(rsel, rv[i].x, ((r2all || yv[i].x > 0.f) ? r2sel : 0.f) * gy[i].x));
                  out.y = fmaf(val[1], mm[1], fmaf(rsel, rv[i].y, ((r2all || yv[i].y > 0.f) ? r2sel : 0.f) * gy[i].y));
                  out.z = fmaf(val[2], mm[2], fmaf(rsel, rv[i].z, ((r2all || yv[i].z > 0.f) ? r2sel : 0.f) * gy[i].z));
                  out.w = fmaf(val[3], mm[3], fmaf(rsel, rv[i].w, ((r2all || yv[i].w > 0.f) ? r2sel : 0.f) * gy[i].w));
                  if (premask) {
                    out.x = xv[i].x > 0.f ? out.x : 0.f;
                    out.y = xv[i].y > 0.f ? out.y : 0.f;
                    out.z = xv[i].z > 0.f ? out.z : 0.f;
                    out.w = xv[i].w > 0.f ? out.w : 0.f;
                  }
                  *(float4*)(p.out + o + (size_t)g * V * N) = out;
                  dm[nc][rd][0] = fmaf(val[0], xv[i].x, dm[nc][rd][0]);
                  dm[nc][rd][1] = fmaf(val[1], xv[i].y, dm[nc][rd][1]);
                  dm[nc][rd][2] = fmaf(val[2], xv[i].z, dm[nc][rd][2]);
                  dm[nc][rd][3] = fmaf(val[3], xv[i].w, dm[nc][rd][3]);
                }
            }
          }
        }
        }
        epi_sync();                                                // staging is reused by the next chunk / tile
      }
    }
    // ---- flush the raw Feature_Mask gradient
#pragma unroll
    for (int nc = 0; nc < NCH; ++nc)
#pragma unroll
      for (int rd = 0; rd < C::kEpiRounds; ++rd) {
        const int slot = et + rd * kEpiThreads;
        if (slot < C::kSlots) {
          const int w = slot >> 4, c = nc * 64 + (slot & 15) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(p.red0 + (size_t)pmod(w - c - j, V) * N + c + j, (double)dm[nc][rd][j]);
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, tmem_cols);
}

template <int V, int K, int N, bool P3>
static int launch_p(const SgcnRowGemm& p, cudaStream_t s) {
  using C = Cfg<V, K, N, P3>;
  auto kern = spatial_bwd_kernel<V, K, N, P3>;
  static std::atomic<unsigned long long> configured{0};           // one bit per device (the attribute is per device)
  if (needs_configure(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem);
    if (e != cudaSuccess) return set_cuda_error("spatial_bwd smem attribute", e);
    mark_configured(configured);
  }
  const long long ntiles = (p.groups + C::G - 1) / C::G;
  if (ntiles == 0) return 0;
  long long grid = tile_ctas();
  if (grid > ntiles) grid = ntiles;
  alignas(64) CUtensorMap tm_res, tm_res2, tm_x;                   // {64 channels, V rows} boxes of the epilogue inputs
  memset(&tm_x, 0, sizeof(tm_x));
  if (C::kTmaIn) {
    if (!p.xin) return set_error("spatial backward: xin is required");
    if (int rc = make_rows_map(&tm_x, p.xin, p.groups * V, N, 64, V)) return rc;
  }
  tm_res = tm_x, tm_res2 = tm_x;
  if (C::kTmaIn && p.res)
    if (int rc = make_rows_map(&tm_res, p.res, p.groups * V, N, 64, V)) return rc;
  if (C::kTmaIn && p.res2)
    if (int rc = make_rows_map(&tm_res2, p.res2, p.groups * V, N, 64, V)) return rc;
  alignas(64) CUtensorMap tm_gh = tm_x, tm_z = tm_x;               // builder inputs [rows, K]
  if (C::kTmaBld) {
    if (int rc = make_rows_map(&tm_gh, p.in0, p.groups * V, K, 64, V)) return rc;
    if (int rc = make_rows_map(&tm_z, p.in1, p.groups * V, K, 64, V)) return rc;
  }
  kern<<<(unsigned)grid, kThreads, C::kSmem, s>>>(p, next_direction(), tm_res, tm_res2, tm_x, tm_gh, tm_z);
  return check_launch("spatial_bwd_kernel");
}

template <int V, int K, int N>
static int launch(const SgcnRowGemm& p, cudaStream_t s) {
  if (p.prec == SGCN_PREC_FP32) return launch_p<V, K, N, true>(p, s);
  if (p.prec != SGCN_PREC_TF32) return set_error("sgcn_rowgemm: unknown precision");
  return launch_p<V, K, N, false>(p, s);
}

template <int V>
static int launch_v(const SgcnRowGemm& p, cudaStream_t s) {
  const int key = p.K * 1000 + p.N;
  switch (key) {
    case 64064: return launch<V, 64, 64>(p, s);
    case 128064: return launch<V, 128, 64>(p, s);
    case 128128: return launch<V, 128, 128>(p, s);
    case 256128: return launch<V, 256, 128>(p, s);
    case 256256: return launch<V, 256, 256>(p, s);
    default: return set_error("spatial backward: unsupported (out, in) channel pair");
  }
}

}  // namespace sb

int spatial_bwd_launch(const SgcnRowGemm& p, cudaStream_t s) {
  if (p.V == 25) return sb::launch_v<25>(p, s);
  if (p.V == 33) return sb::launch_v<33>(p, s);
  return set_error("spatial backward: num_point must be 25 (NTU) or 33 (MediaPipe)");
}

}  // namespace sgcn
