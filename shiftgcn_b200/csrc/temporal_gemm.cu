// temporal_gemm.cu -- instantiations of the warp-specialised row-GEMM (fused_gemm.cuh) for the 1x1 convolution of the
// temporal unit: forward with the BN + fractional temporal shift prologue (model/shift_gcn.py:66-70) and the plain
// backward-data contraction.
#include "fused_gemm.cuh"

namespace sgcn {

template <int PRO, int V>
static int temporal_gemm_v(const SgcnRowGemm& p, cudaStream_t s) {
  using namespace fg;
  switch (p.K * 1000 + p.N) {
    case 64064: return launch<PRO, EPI_LINEAR, V, 64, 64>(p, s);
    case 128128: return launch<PRO, EPI_LINEAR, V, 128, 128>(p, s);
    case 256256: return launch<PRO, EPI_LINEAR, V, 256, 256>(p, s);
    default: break;
  }
  if constexpr (PRO == PRO_PLAIN) {   // 1x1 conv side branches: forward (C -> D) and input gradient ([g | x]: D + C -> C)
    switch (p.K * 1000 + p.N) {
      case 64128: return launch<PRO, EPI_LINEAR, V, 64, 128>(p, s);
      case 128256: return launch<PRO, EPI_LINEAR, V, 128, 256>(p, s);
      case 192064: return launch<PRO, EPI_LINEAR, V, 192, 64>(p, s);
      case 384128: return launch<PRO, EPI_LINEAR, V, 384, 128>(p, s);
      default: break;
    }
  }
  return set_error("row GEMM: unsupported (in, out) channel pair");
}

// the whole eval-mode temporal unit in one kernel (EPI_TSHIFT): joint subsets of 5 over 25 consecutive frames
int temporal_unit_launch(const SgcnRowGemm& p, cudaStream_t s) {
  using namespace fg;
  if (p.V != 25) return set_error("fused temporal unit: num_point must be 25");
  switch (p.K * 1000 + p.N) {
    case 64064: return launch<PRO_LERP, EPI_TSHIFT, 5, 64, 64>(p, s);
    case 128128: return launch<PRO_LERP, EPI_TSHIFT, 5, 128, 128>(p, s);
    case 256256: return launch<PRO_LERP, EPI_TSHIFT, 5, 256, 256>(p, s);
    default: return set_error("fused temporal unit: unsupported channel count");
  }
}

int temporal_gemm_launch(const SgcnRowGemm& p, int lerp, cudaStream_t s) {
  if (p.V == 25) return lerp ? temporal_gemm_v<fg::PRO_LERP, 25>(p, s) : temporal_gemm_v<fg::PRO_PLAIN, 25>(p, s);
  if (p.V == 33) return lerp ? temporal_gemm_v<fg::PRO_LERP, 33>(p, s) : temporal_gemm_v<fg::PRO_PLAIN, 33>(p, s);
  return set_error("temporal 1x1 convolution: num_point must be 25 (NTU) or 33 (MediaPipe)");
}

}  // namespace sgcn
