// tensormap.h -- TMA tensor maps over the channels-last row tensors [(group, joint), channel] of the hot path.
//
// Host: cuTensorMapEncodeTiled is taken from the driver through cudaGetDriverEntryPoint (the library links cudart only).
// Maps are encoded per launch (a pure host-side bit-packing call, ~1 us) and travel to the kernel by value as
// __grid_constant__ parameters, so a captured CUDA graph replays them with the same addresses.
// Device: cp.async.bulk.tensor loads with mbarrier complete_tx; out-of-range coordinates (halo rows in front of the
// tensor, rows past a partial last tile) are ZERO filled by the unit and still count their full box in the tx bytes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgcn {

// 2-D map of a row tensor [rows, pitch] fp32, box = {box_ch channels, box_rows rows}, dense rows in shared memory
// (no swizzle: the builders read it lane <-> channel).  Returns 0 / error code.
int make_rows_map(CUtensorMap* m, const float* base, long long rows, int pitch, int box_ch, int box_rows);
// 3-D map [groups, V, pitch] of the same tensor where group g starts at frame-group g * gs (1x1 convolution with a frame
// stride): box = {32 channels, V joints, box_groups groups} written as [box_groups * V rows x 128 B] in the UMMA
// K-major SWIZZLE_128B layout, i.e. directly as one 32-channel operand block.
int make_groups_map_sw128(CUtensorMap* m, const float* base, long long groups, long long gs, int V, int pitch, int box_groups);

// 3-D map [frames, joints, pitch] of a row tensor, dense box {box_ch channels, box_v joints, box_frames frames}: the
// "joints x frames" tiles of the fused temporal unit (a joint subset over consecutive frames of one sample).
int make_frames_map(CUtensorMap* m, const float* base, long long frames, int V, int pitch, int box_ch, int box_v,
                    int box_frames);

#ifdef __CUDACC__
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace sgcn
