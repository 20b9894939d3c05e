// wgrad.cu -- weight-gradient contractions  dW[a, b] = sum_rows  A[row, a] * B[row, b]
// on tcgen05 tensor cores, with the operands built on the fly by the same prologues as rowgemm.cu.
//
//   spatial  (autograd of model/shift_gcn.py:131):  A = xm  (gather + mask of x),  B = dy (BN1d-bwd + inverse rotation)
//            -> Linear_weight.grad [C, D]
//   temporal (autograd of model/shift_gcn.py:69):   A = dpre (grad wrt the conv output, pre-ReLU), B = p (shifted BN(h))
//            -> temporal_linear.weight.grad [Cout, Cin]
//
// Both operands are 128-row tiles in the SWIZZLE_128B_BASE32B layout (the only one tcgen05 accepts for MN-major
// 32-bit operands) read through MN-major UMMA descriptors: the row index is the contraction dimension, 8 rows
// (two 4-row swizzle atoms) per instruction.  A CTA owns one (A-channel block, B-channel block) pair and
// a strided subset of the row tiles, keeps its partial dW in TMEM across all of them and flushes it once with
// atomics.  The M extent is always 128: when the A block has only 64 channels the upper 64 TMEM lanes hold
// don't-care values that are never read back.
#include "capi_internal.h"
#include "common.cuh"
#include "tile_builders.cuh"
#include "wgrad.h"

namespace sgcn {

enum { WG_SPATIAL = 0, WG_TEMPORAL = 1 };

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const SgcnWgrad p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V, G = p.G;
  const int MC = p.CA < 128 ? p.CA : 128, NC = p.CB < 128 ? p.CB : 128;   // channels per block pair
  const int na = p.CA / MC, nb = p.CB / NC;
  const int npairs = na * nb;
  const int pair = blockIdx.x % npairs, split = blockIdx.x / npairs, nsplit = gridDim.x / npairs;
  const int a0 = (pair / nb) * MC, b0 = (pair % nb) * NC;

  uint8_t* sA = smem;                       // 4 blocks (M = 128 channels addressed, MC valid)
  uint8_t* sB = sA + 4 * kBlockBytes;       // NC / 32 blocks
  uint8_t* sS = sB + (NC / 32) * kBlockBytes;   // two [128 x 64] input stages (SPATIAL: x, then gh + z)
  float* sLerp = (float*)(sS + (MODE == WG_SPATIAL ? 2 * kTileRows * 256 : 0));
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sGrpT[16];

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  const uint32_t tmem_cols = NC <= 64 ? 64u : 128u;
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (MODE == WG_TEMPORAL) load_lerp_tables(sLerp, p.b_tab0, p.b_tab1, p.b_tab2, p.CB, tid);
  for (int i = tid; i < (4 * kBlockBytes + (NC / 32) * kBlockBytes) / 16; i += kThreads)
    ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(128, NC, 1, 1);
  uint32_t mma_phase = 0;
  bool first = true;

  const long long ntiles = (p.groups + G - 1) / G;
  for (long long tile = split; tile < ntiles; tile += nsplit) {
    const long long g0 = tile * G;
    const int ng = (int)((p.groups - g0) < G ? (p.groups - g0) : G);
    const int rows_valid = ng * V;
    const long long row0 = g0 * V;

    if (!first) {                           // previous tile's MMAs must be done reading sA / sB
      mbar_wait(&bar_mma, mma_phase);
      mma_phase ^= 1;
    }
    if (MODE == WG_TEMPORAL) {
      if (tid < G) sGrpT[tid] = (int)((g0 + tid) % p.T);
      __syncthreads();
    }
    // ---- A operand, 64 channels at a time
    for (int ac = 0; ac < MC; ac += 64) {
      uint8_t* chunk = sA + (ac / 32) * kBlockBytes;
      if (MODE == WG_SPATIAL) {
        stage_rows_async(sS, p.a_src, row0, rows_valid, p.CA, a0 + ac, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        build_spatial_chunk<true>(chunk, (const float*)sS, p.a_tab0, p.CA, a0 + ac, V, ng, warp, lane);
        __syncthreads();
      } else {
        build_plain_chunk<true>(chunk, p.a_src, p.CA, a0 + ac, row0, rows_valid, warp, lane);
      }
      if (rows_valid < G * V) zero_tail_rows<true>(chunk, rows_valid, warp, lane);
    }
    // ---- B operand
    for (int bc = 0; bc < NC; bc += 64) {
      uint8_t* chunk = sB + (bc / 32) * kBlockBytes;
      if (MODE == WG_SPATIAL) {
        stage_rows_async(sS, p.b_src, row0, rows_valid, p.CB, b0 + bc, tid);
        stage_rows_async(sS + kTileRows * 256, p.b_src2, row0, rows_valid, p.CB, b0 + bc, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        build_dy_chunk<true>(chunk, (const float*)sS, (const float*)(sS + kTileRows * 256), p.b_tab0,
                       p.b_tab1, p.b_tab2, p.CB, b0 + bc, V, ng, warp, lane);
        __syncthreads();
      } else {
        build_lerp_chunk<true>(chunk, p.b_src, sLerp, sGrpT, p.CB, b0 + bc, V, p.T, g0, rows_valid, warp, lane);
      }
      if (rows_valid < G * V) zero_tail_rows<true>(chunk, rows_valid, warp, lane);
    }
    fence_proxy_async();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
#pragma unroll
      for (int r8 = 0; r8 < 16; ++r8)
        umma_tf32(tmem_base, umma_desc(sa + r8 * 1024, kBlockBytes, 512, 1), umma_desc(sb + r8 * 1024, kBlockBytes, 512, 1),
                  idesc, (first && r8 == 0) ? 0u : 1u);
      tc_commit(&bar_mma);
    }
    first = false;
  }

  if (!first) {
    mbar_wait(&bar_mma, mma_phase);
    tc_fence_after();
    if (warp < 4) {  // warp w reads TMEM lanes [32w, 32w+32) = A channels a0 + 32w + lane
      const int ch = warp * 32 + lane;
      for (int c0 = 0; c0 < NC; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (ch < MC) {
          float* dst = p.dw + (size_t)(a0 + ch) * p.CB + b0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
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
  const int MC = p.CA < 128 ? p.CA : 128, NC = p.CB < 128 ? p.CB : 128;
  const int npairs = (p.CA / MC) * (p.CB / NC);
  const size_t smem = 1024 + (size_t)4 * kBlockBytes + (size_t)(NC / 32) * kBlockBytes +
                      (MODE == WG_SPATIAL ? (size_t)2 * kTileRows * 256 : (size_t)4 * p.CB * 4) + 64;
  auto kern = wgrad_kernel<MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error("wgrad smem attribute", e);
  const long long ntiles = (p.groups + p.G - 1) / p.G;
  if (ntiles == 0) return 0;
  long long nsplit = num_sms() / npairs;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ntiles) nsplit = ntiles;
  kern<<<(unsigned)(nsplit * npairs), kThreads, smem, s>>>(p);
  return check_launch("wgrad_kernel");
}

}  // namespace sgcn

extern "C" int sgcn_wgrad(const SgcnWgrad* pp, int mode, void* stream) {
  using namespace sgcn;
  if (!pp) return set_error("sgcn_wgrad: null params");
  const SgcnWgrad& p = *pp;
  if (p.V < 25 || p.V > 40) return set_error("sgcn_wgrad: num_point must be in [25, 40]");
  if (p.G < 1 || p.G > 16 || p.G * p.V > kTileRows) return set_error("sgcn_wgrad: need G <= 16 and G*V <= 128");
  if ((p.CA != 64 && p.CA != 128 && p.CA != 256) || (p.CB != 64 && p.CB != 128 && p.CB != 256))
    return set_error("sgcn_wgrad: channel counts must be 64, 128 or 256");
  if (!p.a_src || !p.b_src || !p.dw) return set_error("sgcn_wgrad: null tensor");
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
