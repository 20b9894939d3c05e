// common.cuh -- sm_100a device toolkit shared by every kernel of the Shift-GCN hot path.
//
// Hand-written PTX wrappers (no CUTLASS/CuTe dependency): mbarrier, cp.async, tcgen05 (TMEM
// alloc / mma / commit / ld), UMMA shared-memory + instruction descriptors, and the canonical
// 128B-swizzled operand tile used by all tensor-core contractions in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgcn {

// ------------------------------------------------------------------------------------------------
// Canonical operand tile ("block"): 128 rows x 32 fp32 (=128 B per row), 8-row groups of 1024 B,
// the 16-byte chunk index XOR-ed with (row % 8)  == UMMA/TMA SWIZZLE_128B.
//   * read as a K-major operand (MN = rows, K = the 32 channels): SBO = 1024 B.
// A [128 x C] tile is C/32 consecutive blocks of kBlockBytes.  (The MN-major variant used by the weight-gradient
// contraction has the same block shape but a different swizzle, see canon_off_mn.)
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kBlockCh = 32;
constexpr int kBlockBytes = kTileRows * kBlockCh * 4;  // 16 KiB

__device__ __forceinline__ uint32_t canon_off(int row, int ch) {  // byte offset inside one block
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((ch >> 2) ^ (row & 7)) & 7) << 4) + ((ch & 3) << 2));
}

// MN-major 32-bit operands (channels contiguous, rows = contraction dimension) must use SWIZZLE_128B_BASE32B:
// same 128-byte row pitch, but 32-byte chunks XOR-ed with (row % 4), atoms of 4 rows (512 B).
__device__ __forceinline__ uint32_t canon_off_mn(int row, int ch) {
  return (uint32_t)(row * 128 + ((((ch >> 3) ^ (row & 3)) & 3) << 5) + ((ch & 7) << 2));
}

template <bool MN>
__device__ __forceinline__ uint32_t tile_off(int row, int ch) {
  return MN ? canon_off_mn(row, ch) : canon_off(row, ch);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 32-bit shared-window accesses (generic pointers cost 64-bit address arithmetic in the inner loops)
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float to_tf32(float x) {  // round-to-nearest (ties away) to 10-bit mantissa
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Same rounding for values the TENSOR CORE will read: kind::tf32 ignores the low 13 mantissa bits, so adding half a
// TF32 ulp to the magnitude is round-to-nearest (ties away) in ONE integer instruction (cvt.rna.tf32 expands to four).
__device__ __forceinline__ float tf32_half_ulp(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// 3xTF32 ("fp32-accurate") operand split: v = hi + lo with hi = TF32-rounded v (exactly representable, low 13 mantissa
// bits zero) and lo = v - hi (exact in fp32), itself pre-rounded for the tensor core.  A contraction then runs as
// Ah*Bh + Al*Bh + Ah*Bl: the products of two 11-bit mantissas are exact in the fp32 accumulator and the dropped Al*Bl
// term is ~2^-22 relative.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
  lo = tf32_half_ulp(v - hi);
}

// ------------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// Wait of a MANY-thread role (builder / epilogue warps) on a hand-over that is usually not there yet: spinning on
// try_wait costs three issue slots per probe on the sub-partition the working warps share, so back off with nanosleep
// between probes.  The single MMA-issuing lane keeps the tight loop above (its latency is on the critical path).
#ifndef SGCN_WAIT_NS
#define SGCN_WAIT_NS 0
#endif
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#if SGCN_WAIT_NS > 0
  const uint32_t addr = smem_u32(bar);
  while (!mbar_try(addr, parity)) __nanosleep(SGCN_WAIT_NS);
#else
  mbar_wait(bar, parity);
#endif
}

// ------------------------------------------------------------------------------------------------ cp.async
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// UMMA shared-memory matrix descriptor (version 1).  lbo/sbo in bytes.
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major 32-bit tiles)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

// UMMA instruction descriptor: kind::tf32, fp32 accumulate, M x N, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // D format = F32
         | (2u << 7)                        // A format = TF32
         | (2u << 10)                       // B format = TF32
         | ((uint32_t)a_mn_major << 15)     //
         | ((uint32_t)b_mn_major << 16)     //
         | ((uint32_t)(N >> 3) << 17)       //
         | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]   (issued by ONE thread)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM -> registers: 16 consecutive fp32 columns (half the register footprint of tmem_ld32)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ misc
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int pmod(int a, int m) {  // non-negative remainder
  int r = a % m;
  return r < 0 ? r + m : r;
}

}  // namespace sgcn
