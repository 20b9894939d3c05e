// selftest.cu -- minimal tcgen05 contractions that pin the descriptor / canonical-tile conventions
// of common.cuh on real hardware (tests/test_gpu_tcgen05.py).  Two shapes:
//   mode 0 ("row tile x weight"):  D[128 x N] = A[128 x K] * B[N x K]^T          (both K-major)
//   mode 1 ("weight gradient"):    D[M  x N] = A[128 x M]^T * B[128 x N], M = 128 (both MN-major)
#include "common.cuh"
#include "capi_internal.h"

namespace sgcn {

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int mode, int K, int N,
                int M2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // operand staging -------------------------------------------------------------------------
  uint8_t* sa = smem;
  uint8_t* sb;
  if (mode == 0) {
    const int kblocks = K / 32;
    sb = sa + kblocks * kBlockBytes;
    for (int i = tid; i < 128 * K; i += blockDim.x) {
      int r = i / K, k = i % K;
      *(float*)(sa + (k >> 5) * kBlockBytes + canon_off(r, k & 31)) = to_tf32(a[i]);
    }
    const uint32_t bblock = (uint32_t)N * 128u;
    for (int i = tid; i < N * K; i += blockDim.x) {
      int n = i / K, k = i % K;
      *(float*)(sb + (k >> 5) * bblock + canon_off(n, k & 31)) = to_tf32(b[i]);
    }
  } else {
    const int mblocks = (M2 + 31) / 32;
    sb = sa + 4 * kBlockBytes;  // A is always addressed as 4 blocks (M = 128); unused ones stay zero
    for (int i = tid; i < 4 * kBlockBytes / 4; i += blockDim.x) ((float*)sa)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 128 * M2; i += blockDim.x) {
      int r = i / M2, m = i % M2;
      *(float*)(sa + (m >> 5) * kBlockBytes + canon_off_mn(r, m & 31)) = to_tf32(a[i]);
    }
    (void)mblocks;
    for (int i = tid; i < 128 * N; i += blockDim.x) {
      int r = i / N, n = i % N;
      *(float*)(sb + (n >> 5) * kBlockBytes + canon_off_mn(r, n & 31)) = to_tf32(b[i]);
    }
  }
  fence_proxy_async();

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  const uint32_t ncols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
      const uint32_t bblock = (uint32_t)N * 128u;
      uint32_t acc = 0;
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const int blk = k8 >> 2, sub = k8 & 3;
        uint64_t ad = umma_desc(smem_u32(sa) + blk * kBlockBytes + sub * 32, 16, 1024);
        uint64_t bd = umma_desc(smem_u32(sb) + blk * bblock + sub * 32, 16, 1024);
        umma_tf32(tmem_base, ad, bd, idesc, acc);
        acc = 1;
      }
    } else {
      const uint32_t idesc = umma_idesc_tf32(128, N, 1, 1);
      uint32_t acc = 0;
      for (int r8 = 0; r8 < 16; ++r8) {  // K = 128 rows, 8 per instruction
        uint64_t ad = umma_desc(smem_u32(sa) + r8 * 1024, kBlockBytes, 512, 1);
        uint64_t bd = umma_desc(smem_u32(sb) + r8 * 1024, kBlockBytes, 512, 1);
        umma_tf32(tmem_base, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();

  // epilogue: warp w owns TMEM lanes [32w, 32w+32)
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) d[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace sgcn

extern "C" int sgcn_selftest_umma(const float* a, const float* b, float* d, int mode, int K, int N, int M2, void* stream) {
  using namespace sgcn;
  if (mode == 0) {
    if (K % 32 != 0 || K < 32 || K > 256 || N % 32 != 0 || N < 32 || N > 256) return set_error("selftest: bad K/N");
    if ((size_t)K * (128 + N) * 4 > 200 * 1024) return set_error("selftest: operands do not fit in shared memory");
  } else {
    if (N % 32 != 0 || N < 32 || N > 256 || M2 < 1 || M2 > 128) return set_error("selftest: bad M/N");
  }
  size_t smem = 1024;
  if (mode == 0) smem += (size_t)(K / 32) * kBlockBytes + (size_t)(K / 32) * N * 128;
  else smem += (size_t)4 * kBlockBytes + (size_t)(N / 32) * kBlockBytes;
  cudaError_t e = cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error("selftest attr", e);
  selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, mode, K, N, M2);
  return check_launch("selftest_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// Descriptor probe (tests/tools only): D[128 x N] = A * B with each operand either K-major (canonical SW128) or
// MN-major with a caller-chosen swizzle function / layout type / LBO / SBO.  Used once to establish the MN-major
// conventions on hardware; kept because it documents them executable-y.
//   a: a_mn ? [K][128] : [128][K]      b: b_mn ? [K][N] : [N][K]        K in {32, 64, 128}
namespace sgcn {
__device__ __forceinline__ uint32_t probe_off(int swz, int row, int ch) {
  if (swz == 0) return canon_off_mn(row, ch);
  if (swz == 1) return canon_off(row, ch);
  return (uint32_t)(row * 128 + ch * 4);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int K, int N, int a_mn,
             int b_mn, int swz, int layout, int lbo, int sbo, int kstep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 4 * kBlockBytes;
  for (int i = tid; i < 12 * kBlockBytes / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  if (a_mn) {
    for (int i = tid; i < K * 128; i += blockDim.x) {
      int k = i / 128, m = i % 128;
      *(float*)(sa + (m >> 5) * kBlockBytes + probe_off(swz, k, m & 31)) = to_tf32(a[i]);
    }
  } else {
    for (int i = tid; i < 128 * K; i += blockDim.x) {
      int r = i / K, k = i % K;
      *(float*)(sa + (k >> 5) * kBlockBytes + canon_off(r, k & 31)) = to_tf32(a[i]);
    }
  }
  const uint32_t bblock = (uint32_t)N * 128u;
  if (b_mn) {
    for (int i = tid; i < K * N; i += blockDim.x) {
      int k = i / N, n = i % N;
      *(float*)(sb + (n >> 5) * kBlockBytes + probe_off(swz, k, n & 31)) = to_tf32(b[i]);
    }
  } else {
    for (int i = tid; i < N * K; i += blockDim.x) {
      int n = i / K, k = i % K;
      *(float*)(sb + (k >> 5) * bblock + canon_off(n, k & 31)) = to_tf32(b[i]);
    }
  }
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  const uint32_t ncols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N, a_mn, b_mn);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const int blk = k8 >> 2, sub = k8 & 3;
      uint64_t ad = a_mn ? umma_desc(smem_u32(sa) + k8 * kstep, lbo, sbo, layout)
                         : umma_desc(smem_u32(sa) + blk * kBlockBytes + sub * 32, 16, 1024);
      uint64_t bd = b_mn ? umma_desc(smem_u32(sb) + k8 * kstep, lbo, sbo, layout)
                         : umma_desc(smem_u32(sb) + blk * bblock + sub * 32, 16, 1024);
      umma_tf32(tmem_base, ad, bd, idesc, k8 ? 1u : 0u);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) d[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}
}  // namespace sgcn

extern "C" int sgcn_selftest_probe(const float* a, const float* b, float* d, int K, int N, int a_mn, int b_mn, int swz,
                                   int layout, int lbo, int sbo, int kstep, void* stream) {
  using namespace sgcn;
  if ((K != 32 && K != 64 && K != 128) || N % 32 != 0 || N < 32 || N > 128) return set_error("probe: bad K/N");
  const size_t smem = 1024 + (size_t)12 * kBlockBytes;
  cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error("probe attr", e);
  probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, K, N, a_mn, b_mn, swz, layout, lbo, sbo, kstep);
  return check_launch("probe_kernel");
}
