// tile_builders.cuh -- prologues that build one 128-row x 64-channel operand chunk (two canonical
// blocks, TF32-rounded) in shared memory.  Shared by rowgemm.cu (forward / backward-data contractions)
// and wgrad.cu (weight-gradient contractions), so both see bit-identical operands.
//
// Thread mapping everywhere: lane <-> channel (32 consecutive channels = one 128-byte swizzled row,
// conflict-free for both the smem gather and the smem store), warp <-> joints / rows.
#pragma once
#include "common.cuh"

namespace sgcn {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// cp.async a [rows_valid x 64] fp32 slab (row pitch `pitch` floats, channel offset ch0) into a dense stage
__device__ __forceinline__ void stage_rows_async(uint8_t* stage, const float* src, long long row0, int rows_valid,
                                                 int pitch, int ch0, int tid) {
  for (int i = tid; i < rows_valid * 16; i += kThreads) {
    const int r = i >> 4, seg = i & 15;
    cp_async16(stage + r * 256 + seg * 16,
               (const uint8_t*)src + ((size_t)(row0 + r) * pitch + (size_t)ch0) * 4 + (size_t)seg * 16);
  }
}

// zero rows [from, 128) of a 2-block chunk (needed when the rows are the contraction dimension)
template <bool MN>
__device__ __forceinline__ void zero_tail_rows(uint8_t* chunk, int from, int warp, int lane) {
  for (int r = from + warp; r < kTileRows; r += kWarps) {
    *(float*)(chunk + tile_off<MN>(r, lane)) = 0.f;
    *(float*)(chunk + kBlockBytes + tile_off<MN>(r, lane)) = 0.f;
  }
}

// xm[(g,u), c] = x[g, (u+c) % V, c] * maskmul[u, c]          (model/shift_gcn.py:127-129)
// sX: staged x slab [rows][64]; maskmul: [V, C] (= tanh(Feature_Mask) + 1)
template <bool MN>
__device__ __forceinline__ void build_spatial_chunk(uint8_t* chunk, const float* sX, const float* __restrict__ maskmul,
                                                    int C, int ch0, int V, int ng, int warp, int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = ch0 + j * 32 + lane;
    uint8_t* blk = chunk + j * kBlockBytes;
    for (int u = warp; u < V; u += kWarps) {
      const int sv = (u + c) % V;
      const float mm = __ldg(maskmul + u * C + c);
      for (int g = 0; g < ng; ++g) {
        const float x = sX[(g * V + sv) * 64 + j * 32 + lane];
        *(float*)(blk + tile_off<MN>(g * V + u, lane)) = to_tf32(x * mm);
      }
    }
  }
}

// dy[(g,u), d] = dz[g, (u+d) % V, d],  dz = alpha*gh + beta*z + gamma   (BN1d backward folded into three
// per-(v,d) tables by sgcn_bn_bwd_finalize; inverse of the shift_out gather, model/shift_gcn.py:135-137)
template <bool MN>
__device__ __forceinline__ void build_dy_chunk(uint8_t* chunk, const float* sG, const float* sZ,
                                               const float* __restrict__ alpha, const float* __restrict__ beta,
                                               const float* __restrict__ gamma, int D, int ch0, int V, int ng, int warp,
                                               int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int d = ch0 + j * 32 + lane;
    uint8_t* blk = chunk + j * kBlockBytes;
    for (int u = warp; u < V; u += kWarps) {
      const int sv = (u + d) % V;
      const float al = __ldg(alpha + sv * D + d), be = __ldg(beta + sv * D + d), ga = __ldg(gamma + sv * D + d);
      for (int g = 0; g < ng; ++g) {
        const int si = (g * V + sv) * 64 + j * 32 + lane;
        *(float*)(blk + tile_off<MN>(g * V + u, lane)) = to_tf32(fmaf(al, sG[si], fmaf(be, sZ[si], ga)));
      }
    }
  }
}

// p[(g,v), c] = (1-fy) * U(t+y1) + fy * U(t+y1+1),  U(t') = a_c * h[n,t',v,c] + b_c inside [0,T) else 0
// (BN followed by the stride-1 temporal shift, model/shift_gcn.py:66-68; K1 of shift_cuda_kernel.cu with xpos = 0)
// sLerp: per-channel tables [4][C] = {floor(y), frac(y), a, b};  sGrpT[g] = frame index of group g
template <bool MN>
__device__ __forceinline__ void build_lerp_chunk(uint8_t* chunk, const float* __restrict__ h, const float* sLerp,
                                                 const int* sGrpT, int C, int ch0, int V, int T, long long g0,
                                                 int rows_valid, int warp, int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = ch0 + j * 32 + lane;
    uint8_t* blk = chunk + j * kBlockBytes;
    const int y1 = (int)sLerp[c];
    const float fy = sLerp[C + c], sa = sLerp[2 * C + c], sb = sLerp[3 * C + c];
    for (int r = warp; r < rows_valid; r += kWarps) {
      const int g = r / V, v = r - g * V;
      const int t = sGrpT[g];
      const int ta = t + y1;
      const float* base = h + ((size_t)(g0 + g - t) * V + v) * C + c;  // element (n, t = 0, v, c)
      float u0 = 0.f, u1 = 0.f;
      if (ta >= 0 && ta < T) u0 = fmaf(sa, __ldg(base + (size_t)ta * V * C), sb);
      if (ta + 1 >= 0 && ta + 1 < T) u1 = fmaf(sa, __ldg(base + (size_t)(ta + 1) * V * C), sb);
      *(float*)(blk + tile_off<MN>(r, lane)) = to_tf32(u0 * (1.f - fy) + u1 * fy);
    }
  }
}

// rows as they are
template <bool MN>
__device__ __forceinline__ void build_plain_chunk(uint8_t* chunk, const float* __restrict__ src, int C, int ch0,
                                                  long long row0, int rows_valid, int warp, int lane) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = ch0 + j * 32 + lane;
    uint8_t* blk = chunk + j * kBlockBytes;
    for (int r = warp; r < rows_valid; r += kWarps)
      *(float*)(blk + tile_off<MN>(r, lane)) = to_tf32(__ldg(src + (size_t)(row0 + r) * C + c));
  }
}

__device__ __forceinline__ void load_lerp_tables(float* sLerp, const float* __restrict__ scale,
                                                 const float* __restrict__ shift, const float* __restrict__ ypos_eff,
                                                 int C, int tid) {
  for (int c = tid; c < C; c += kThreads) {
    const float y = ypos_eff[c];
    const float y1 = floorf(y);
    sLerp[c] = y1;
    sLerp[C + c] = y - y1;
    sLerp[2 * C + c] = scale[c];
    sLerp[3 * C + c] = shift[c];
  }
}

}  // namespace sgcn
