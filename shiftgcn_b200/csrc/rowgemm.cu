// rowgemm.cu -- the fused "row tile x weight" tensor-core kernels of the Shift-GCN hot path.
//
// Activations live channels-last: a logical (n, C, T, V) tensor is the row-major matrix [(n,t,v), C].
// A CTA processes tiles of G whole (n,t) groups (G*V <= 128 rows), so the joint shift of the spatial
// unit is a permutation inside the tile.  Per tile:
//     prologue   build the 128 x 64 operand chunk in shared memory (canonical SWIZZLE_128B tile, TF32-rounded)
//     contraction tcgen05.mma (kind::tf32, M = 128, N = out channels, fp32 accumulators in TMEM)
//     epilogue   tcgen05.ld -> padded smem staging -> channel-contiguous pass with the fused tail
//
// Variants (template PRO x EPI), with the reference code each one replaces:
//   PRO_SPATIAL  x[r,(u+c)%V,c] * (tanh(mask)+1)            model/shift_gcn.py:123-129
//   PRO_LERP     temporal shift of BN(h), zero padded       model/shift_gcn.py:66-68, shift_cuda_kernel.cu:12-76
//   PRO_PLAIN    rows as they are                           (1x1 convs, backward data contractions)
//   PRO_DY       BN1d-backward + inverse output rotation    autograd of model/shift_gcn.py:135-137
//   EPI_ROT_RAW    + bias, rotate z[v,d]=y[(v-d)%V,d], store, per-(v,d) batch statistics   :131-137 (training)
//   EPI_ROT_FUSED  + bias, rotate, folded BN, + residual, ReLU                              :131-141 (eval)
//   EPI_LINEAR     + bias, optional ReLU, store                                             :69-70
//   EPI_SPATIAL_BWD  * mask, inverse input rotation, + residual gradients, dMask partials   autograd of :127-129,140
#include "capi_internal.h"
#include "common.cuh"
#include "rowgemm.h"
#include "tile_builders.cuh"

namespace sgcn {

enum { PRO_SPATIAL = 0, PRO_LERP = 1, PRO_PLAIN = 2, PRO_DY = 3 };
enum { EPI_ROT_RAW = 0, EPI_ROT_FUSED = 1, EPI_LINEAR = 2, EPI_SPATIAL_BWD = 3 };

constexpr int kStagePitch = 68;                              // floats per staging row (64 + 4 pad)
constexpr int kStageBytes = kTileRows * kStagePitch * 4;     // 34816
constexpr int kChunkBytes = 2 * kBlockBytes;                 // one 128 x 64 operand chunk
constexpr int kMaxVI = 5;                                    // joints handled per warp: ceil(V / 8), V <= 40
constexpr int kWResidentMax = 64 * 1024;

__host__ __device__ inline size_t rowgemm_w_bytes(int K, int N) {
  size_t full = (size_t)K * N * 4;
  return full <= (size_t)kWResidentMax ? full : (size_t)N * 256;
}
__host__ __device__ inline size_t rowgemm_u_bytes(int pro) {
  return pro == PRO_DY ? (size_t)2 * kTileRows * 256 : (size_t)kStageBytes;
}

template <int PRO, int EPI, int NCH>
__global__ void __launch_bounds__(kThreads, NCH == 1 ? 2 : 1) rowgemm_kernel(const SgcnRowGemm p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V, G = p.G, K = p.K, N = p.N;
  const int kchunks = K >> 6;
  const bool w_resident = (size_t)K * N * 4 <= (size_t)kWResidentMax;

  uint8_t* sA = smem;                                   // operand chunk (2 blocks)
  uint8_t* sW = sA + kChunkBytes;                       // weight image (resident) or one streamed chunk
  uint8_t* sU = sW + rowgemm_w_bytes(K, N);             // input stage(s) / output staging (aliased)
  float* sTab = (float*)(sU + rowgemm_u_bytes(PRO));    // bias[N] (+ LERP: y1, dy, a, b per input channel)
  float* sBias = sTab;
  float* sLerp = sTab + N;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sGrpT[16];                               // LERP: frame index t of each group of the tile

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  const uint32_t tmem_cols = N <= 64 ? 64u : (N <= 128 ? 128u : 256u);
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  for (int i = tid; i < N; i += kThreads) sBias[i] = p.bias ? p.bias[i] : 0.f;
  if (PRO == PRO_LERP) load_lerp_tables(sLerp, p.pro_a, p.pro_b, p.pro_c, K, tid);
  if (w_resident) {
    const int n16 = (K * N * 4) >> 4;
    for (int i = tid; i < n16; i += kThreads) cp_async16(sW + (size_t)i * 16, (const uint8_t*)p.wimg + (size_t)i * 16);
    cp_async_commit();
    cp_async_wait_all();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
  uint32_t mma_phase = 0;

  // per-thread reductions that live across tiles
  float acc0[NCH][2][kMaxVI];   // ROT_RAW: sum ; SPATIAL_BWD: dMask
  float acc1[NCH][2][kMaxVI];   // ROT_RAW: sum of squares
#pragma unroll
  for (int a = 0; a < NCH; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < kMaxVI; ++c) acc0[a][b][c] = 0.f, acc1[a][b][c] = 0.f;

  const long long ntiles = (p.groups + G - 1) / G;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long g0 = tile * G;
    const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
    const int rows_valid = ng * V;
    const long long row0 = g0 * V;                       // first global row of the tile

    if (PRO == PRO_LERP) {
      if (tid < G) sGrpT[tid] = (int)((g0 + tid) % p.T);
      __syncthreads();
    }
    // ================================================================== K loop: prologue + MMA
    for (int kc = 0; kc < kchunks; ++kc) {
      // ---- stage inputs with cp.async (and the streamed weight chunk)
      if (PRO == PRO_SPATIAL || PRO == PRO_DY) stage_rows_async(sU, p.in0, row0, rows_valid, K, kc * 64, tid);
      if (PRO == PRO_DY) stage_rows_async(sU + kTileRows * 256, p.in1, row0, rows_valid, K, kc * 64, tid);
      if (kc > 0) {                                      // previous chunk's MMAs must be done reading sA / sW
        mbar_wait(&bar_mma, mma_phase);
        mma_phase ^= 1;
      }
      if (!w_resident) {
        const int n16 = (N * 256) >> 4;
        const uint8_t* src = (const uint8_t*)p.wimg + (size_t)kc * N * 256;
        for (int i = tid; i < n16; i += kThreads) cp_async16(sW + (size_t)i * 16, src + (size_t)i * 16);
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();

      // ---- build the operand chunk
      if (PRO == PRO_SPATIAL) {
        build_spatial_chunk<false>(sA, (const float*)sU, p.pro_a, K, kc * 64, V, ng, warp, lane);
      } else if (PRO == PRO_DY) {
        build_dy_chunk<false>(sA, (const float*)sU, (const float*)(sU + kTileRows * 256), p.pro_a, p.pro_b, p.pro_c, K,
                       kc * 64, V, ng, warp, lane);
      } else if (PRO == PRO_LERP) {
        build_lerp_chunk<false>(sA, p.in0, sLerp, sGrpT, K, kc * 64, V, p.T, g0, rows_valid, warp, lane);
      } else {
        build_plain_chunk<false>(sA, p.in0, K, kc * 64, row0, rows_valid, warp, lane);
      }
      fence_proxy_async();
      __syncthreads();

      // ---- contraction: D[128 x N] (+)= A[128 x 64] * W[N x 64]^T
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA);
        const uint32_t w0 = smem_u32(sW) + (w_resident ? (uint32_t)kc * (uint32_t)N * 256u : 0u);
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          const uint32_t blk = k8 >> 2, sub = k8 & 3;
          umma_tf32(tmem_base, umma_desc(a0 + blk * kBlockBytes + sub * 32, 16, 1024),
                    umma_desc(w0 + blk * (uint32_t)N * 128u + sub * 32, 16, 1024), idesc, (kc | k8) ? 1u : 0u);
        }
        tc_commit(&bar_mma);
      }
    }
    mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();

    // ================================================================== epilogue, 64 output channels at a time
    float* sS = (float*)sU;
#pragma unroll
    for (int nc = 0; nc < NCH; ++nc) {
      {  // TMEM -> staging: warp handles lane quarter (warp & 3), column half (warp >> 2)
        const int q = warp & 3, hf = warp >> 2;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nc * 64 + hf * 32), v);
        const int row = q * 32 + lane;
        float4* dst = (float4*)(sS + row * kStagePitch + hf * 32);
        const float* bb = sBias + nc * 64 + hf * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(v[4 * i] + bb[4 * i], v[4 * i + 1] + bb[4 * i + 1], v[4 * i + 2] + bb[4 * i + 2],
                               v[4 * i + 3] + bb[4 * i + 3]);
      }
      tc_fence_before();
      __syncthreads();

      if (EPI == EPI_ROT_RAW || EPI == EPI_ROT_FUSED) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int dl = j * 32 + lane, d = nc * 64 + dl;
#pragma unroll
          for (int vi = 0; vi < kMaxVI; ++vi) {
            const int v = warp + vi * kWarps;
            if (v < V) {
              const int sv = pmod(v - d, V);
              float sc = 1.f, sh = 0.f;
              if (EPI == EPI_ROT_FUSED) sc = __ldg(p.epi_a + v * N + d), sh = __ldg(p.epi_b + v * N + d);
              float s1 = 0.f, s2 = 0.f;
              for (int g = 0; g < ng; ++g) {
                const float z = sS[(g * V + sv) * kStagePitch + dl];
                const size_t o = (size_t)(row0 + g * V + v) * N + d;
                if (EPI == EPI_ROT_RAW) {
                  p.out[o] = z;
                  s1 += z;
                  s2 = fmaf(z, z, s2);
                } else {
                  float r = fmaf(z, sc, sh);
                  if (p.res) r += __ldg(p.res + o);
                  p.out[o] = p.relu ? fmaxf(r, 0.f) : r;
                }
              }
              acc0[nc][j][vi] += s1;
              acc1[nc][j][vi] += s2;
            }
          }
        }
      } else if (EPI == EPI_LINEAR) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int dl = j * 32 + lane, d = nc * 64 + dl;
          for (int r = warp; r < rows_valid; r += kWarps) {
            const float z = sS[r * kStagePitch + dl];
            p.out[(size_t)(row0 + r) * N + d] = p.relu ? fmaxf(z, 0.f) : z;
          }
        }
      } else {  // EPI_SPATIAL_BWD: N == input channels of the unit
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cl = j * 32 + lane, c = nc * 64 + cl;
#pragma unroll
          for (int vi = 0; vi < kMaxVI; ++vi) {
            const int w = warp + vi * kWarps;
            if (w < V) {
              const int u = pmod(w - c, V);
              const float mm = __ldg(p.epi_a + u * N + c);
              float dm = 0.f;
              for (int g = 0; g < ng; ++g) {
                const float dxm = sS[(g * V + u) * kStagePitch + cl];
                const size_t o = (size_t)(row0 + g * V + w) * N + c;
                float r = dxm * mm;
                if (p.res) r += __ldg(p.res + o);
                if (p.res2) r += (__ldg(p.res2m + o) > 0.f) ? __ldg(p.res2 + o) : 0.f;
                dm = fmaf(dxm, __ldg(p.xin + o), dm);
                p.out[o] = r;
              }
              acc0[nc][j][vi] += dm;
            }
          }
        }
      }
      __syncthreads();  // staging is reused by the next chunk / next tile's input stage
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---------------------------------------------------------------- flush cross-tile reductions
  if (EPI == EPI_ROT_RAW) {
#pragma unroll
    for (int nc = 0; nc < NCH; ++nc)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int vi = 0; vi < kMaxVI; ++vi) {
          const int v = warp + vi * kWarps, d = nc * 64 + j * 32 + lane;
          if (v < V) {
            atomicAdd(p.stats + 2 * ((size_t)v * N + d), (double)acc0[nc][j][vi]);
            atomicAdd(p.stats + 2 * ((size_t)v * N + d) + 1, (double)acc1[nc][j][vi]);
          }
        }
  }
  if (EPI == EPI_SPATIAL_BWD) {
#pragma unroll
    for (int nc = 0; nc < NCH; ++nc)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int vi = 0; vi < kMaxVI; ++vi) {
          const int w = warp + vi * kWarps, c = nc * 64 + j * 32 + lane;
          if (w < V) atomicAdd(p.red0 + (size_t)pmod(w - c, V) * N + c, (double)acc0[nc][j][vi]);
        }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
template <int PRO, int EPI, int NCH>
static int launch_variant(const SgcnRowGemm& p, cudaStream_t s) {
  const size_t smem = 1024 + kChunkBytes + rowgemm_w_bytes(p.K, p.N) + rowgemm_u_bytes(PRO) +
                      (size_t)(p.N + (PRO == PRO_LERP ? 4 * p.K : 0)) * 4 + 64;
  auto kern = rowgemm_kernel<PRO, EPI, NCH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error("rowgemm smem attribute", e);
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
  if (e != cudaSuccess) return set_cuda_error("rowgemm occupancy", e);
  if (occ < 1) return set_error("rowgemm: kernel does not fit on an SM");
  const uint32_t tmem_cols = p.N <= 64 ? 64u : (p.N <= 128 ? 128u : 256u);
  while (occ > 1 && (uint32_t)occ * tmem_cols > 512u) --occ;     // TMEM columns are per SM
  const long long ntiles = (p.groups + p.G - 1) / p.G;
  if (ntiles == 0) return 0;
  long long grid = (long long)num_sms() * occ;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kThreads, smem, s>>>(p);
  return check_launch("rowgemm_kernel");
}

template <int PRO, int EPI>
static int launch_nch(const SgcnRowGemm& p, cudaStream_t s) {
  switch (p.N >> 6) {
    case 1: return launch_variant<PRO, EPI, 1>(p, s);
    case 2: return launch_variant<PRO, EPI, 2>(p, s);
    case 4: return launch_variant<PRO, EPI, 4>(p, s);
    default: return set_error("rowgemm: N must be 64, 128 or 256");
  }
}

}  // namespace sgcn

extern "C" int sgcn_rowgemm(const SgcnRowGemm* pp, int pro, int epi, void* stream) {
  using namespace sgcn;
  if (!pp) return set_error("sgcn_rowgemm: null params");
  const SgcnRowGemm& p = *pp;
  if (p.V < 1 || p.V > 8 * kMaxVI) return set_error("sgcn_rowgemm: num_point must be in [1, 40]");
  if (p.G < 1 || p.G > 16 || p.G * p.V > kTileRows) return set_error("sgcn_rowgemm: need G <= 16 and G*V <= 128");
  if (p.K != 64 && p.K != 128 && p.K != 256) return set_error("sgcn_rowgemm: K must be 64, 128 or 256");
  if (p.N != 64 && p.N != 128 && p.N != 256) return set_error("sgcn_rowgemm: N must be 64, 128 or 256");
  if (p.groups < 0) return set_error("sgcn_rowgemm: negative group count");
  if (!p.in0 || !p.out || !p.wimg) return set_error("sgcn_rowgemm: null tensor");
  cudaStream_t s = (cudaStream_t)stream;
  if (pro == PRO_SPATIAL && epi == EPI_ROT_RAW) {
    if (!p.pro_a || !p.stats) return set_error("spatial fwd: null mask / stats");
    return launch_nch<PRO_SPATIAL, EPI_ROT_RAW>(p, s);
  }
  if (pro == PRO_SPATIAL && epi == EPI_ROT_FUSED) {
    if (!p.pro_a || !p.epi_a || !p.epi_b) return set_error("spatial fwd (fused): null table");
    return launch_nch<PRO_SPATIAL, EPI_ROT_FUSED>(p, s);
  }
  if (pro == PRO_LERP && epi == EPI_LINEAR) {
    if (!p.pro_a || !p.pro_b || !p.pro_c || p.T < 1) return set_error("temporal fwd: null table / bad T");
    return launch_nch<PRO_LERP, EPI_LINEAR>(p, s);
  }
  if (pro == PRO_PLAIN && epi == EPI_LINEAR) return launch_nch<PRO_PLAIN, EPI_LINEAR>(p, s);
  if (pro == PRO_DY && epi == EPI_SPATIAL_BWD) {
    if (!p.in1 || !p.pro_a || !p.pro_b || !p.pro_c || !p.epi_a || !p.xin || !p.red0)
      return set_error("spatial bwd: null tensor / table");
    return launch_nch<PRO_DY, EPI_SPATIAL_BWD>(p, s);
  }
  return set_error("sgcn_rowgemm: unsupported prologue/epilogue combination");
}
