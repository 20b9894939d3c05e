// spatial_fwd.cu -- instantiations of the warp-specialised row-GEMM (fused_gemm.cuh) for the forward pass of the
// spatial unit: joint-shift gather + mask prologue, rotate + {raw store and BN statistics | folded BN + residual + ReLU}
// epilogue (model/shift_gcn.py:123-141).
#include "fused_gemm.cuh"

namespace sgcn {

template <int EPI, int V>
static int spatial_fwd_v(const SgcnRowGemm& p, cudaStream_t s) {
  using namespace fg;
  switch (p.K * 1000 + p.N) {
    case 64064: return launch<PRO_SPATIAL, EPI, V, 64, 64>(p, s);
    case 64128: return launch<PRO_SPATIAL, EPI, V, 64, 128>(p, s);
    case 128128: return launch<PRO_SPATIAL, EPI, V, 128, 128>(p, s);
    case 128256: return launch<PRO_SPATIAL, EPI, V, 128, 256>(p, s);
    case 256256: return launch<PRO_SPATIAL, EPI, V, 256, 256>(p, s);
    default: return set_error("spatial forward: unsupported (in, out) channel pair");
  }
}

int spatial_fwd_launch(const SgcnRowGemm& p, int fused, cudaStream_t s) {
  if (p.V == 25) return fused ? spatial_fwd_v<fg::EPI_ROT_FUSED, 25>(p, s) : spatial_fwd_v<fg::EPI_ROT_RAW, 25>(p, s);
  if (p.V == 33) return fused ? spatial_fwd_v<fg::EPI_ROT_FUSED, 33>(p, s) : spatial_fwd_v<fg::EPI_ROT_RAW, 33>(p, s);
  return set_error("spatial forward: num_point must be 25 (NTU) or 33 (MediaPipe)");
}

}  // namespace sgcn
