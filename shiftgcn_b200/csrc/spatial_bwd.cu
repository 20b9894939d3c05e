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
//                          (joint w, 4 channels) and walks the tile's groups: 128-bit loads of gh / g_y / y / x,
//                          four scalar gathers from the staging tile (the rotation), 128-bit store of gx,
//                          dMask partial sums in registers for the whole kernel
//   warp 13                issues tcgen05.mma, streams weight chunks (cp.async.bulk) when W does not fit
//   builder warps 14..21   own SOURCE slots (joint sv, 4 channels): 128-bit loads of gh / z, dz, scatter of the
//                          four channels to their rotated rows of the K-major SWIZZLE_128B operand chunk
// Pipelines: operand chunks (2 buffers, full/free mbarriers), TMEM accumulators (2 buffers, full/free), so the
// builders run up to two 64-channel chunks ahead of the tensor core and the epilogue of tile i overlaps tile i+1.
#include "capi_internal.h"
#include "common.cuh"
#include "rowgemm.h"

namespace sgcn {

namespace sb {

constexpr int kEpiWarps = 13, kBldWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32, kBldThreads = kBldWarps * 32;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = (kEpiWarps + 1 + kBldWarps) * 32;       // 704
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
  static constexpr bool kWRes = ((P3 ? 2 : 1) * kImgBytes) <= 65536;
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
  static constexpr size_t kSmem = 1024 + kWBytes + kOpStages * kOpBytes + kChunkBytes + 64;
  static_assert(kSmem <= 232448 - 512, "shared memory budget");
};

template <int V, int K, int N, bool P3>
__global__ void __launch_bounds__(kThreads, 1) spatial_bwd_kernel(const SgcnRowGemm p, const int rev) {
  using C = Cfg<V, K, N, P3>;
  constexpr int kOpBytes = C::kOpBytes;
  constexpr int G = C::G, KC = C::KC, NCH = C::NCH, OS = C::kOpStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sOp = sW + C::kWBytes;                                  // OS operand chunks
  uint8_t* sSt = sOp + OS * kOpBytes;                              // epilogue staging
  __shared__ uint64_t op_full[4], op_free[4], acc_full[2], acc_free[2], w_full, w_free;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&op_full[i], kBldThreads);
      mbar_init(&op_free[i], 1);
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
  } else if (warp > kMmaWarp) {
    // ================================================================================ builders
    const int bt = tid - (kMmaWarp + 1) * 32;
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
            // destination rows u_j = (sv - d - j) mod V; byte offset of (row, channel) in the K-major swizzled chunk
            int u0 = sv - d % V;
            if (u0 < 0) u0 += V;
            const uint32_t blk_off = (uint32_t)(c4 >> 3) * kBlockBytes;
            const uint32_t cc = (uint32_t)(c4 & 7);
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (g < ng) {
                const float dz[4] = {fmaf(al.x, gv[g].x, fmaf(be.x, zv[g].x, ga.x)), fmaf(al.y, gv[g].y, fmaf(be.y, zv[g].y, ga.y)),
                                     fmaf(al.z, gv[g].z, fmaf(be.z, zv[g].z, ga.z)), fmaf(al.w, gv[g].w, fmaf(be.w, zv[g].w, ga.w))};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  int u = u0 - j;
                  if (u < 0) u += V;
                  const uint32_t row = (uint32_t)(g * V + u);
                  float* dst = (float*)(op + blk_off + (row >> 3) * 1024u + (row & 7u) * 128u + ((cc ^ (row & 7u)) << 4) + j * 4);
                  if constexpr (P3) {
                    float hi, lo;
                    split_tf32(dz[j], hi, lo);
                    *dst = hi;
                    *(float*)((uint8_t*)dst + kChunkBytes) = lo;
                  } else {
                    *dst = tf32_half_ulp(dz[j]);
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

    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = rev ? ntiles - 1 - (blockIdx.x + ti * gridDim.x) : blockIdx.x + ti * gridDim.x;
      const long long g0 = tile * G;
      const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
      const size_t row0 = (size_t)g0 * V;
      const int buf = (int)(ti & 1);
      // Only the TMEM-reading warps wait for the accumulator: they are the ones whose arrival on acc_free lets the
      // issuer reuse the buffer, so the barrier can never run two phases ahead of a waiter (a warp that waited
      // without gating acc_free could miss a whole phase and spin forever).  Warps 8..12 are ordered by epi_sync.
      if (warp < 8) {
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
  kern<<<(unsigned)grid, kThreads, C::kSmem, s>>>(p, next_direction());
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
