// rowgemm.cu -- entry point of the fused "row tile x weight" tensor-core kernels of the Shift-GCN hot path.
//
// Activations live channels-last: a logical (n, C, T, V) tensor is the row-major matrix [(n,t,v), C].  sgcn_rowgemm
// validates the parameter block and hands it to the warp-specialised kernels:
//   fused_gemm.cuh (instantiated in spatial_fwd.cu / temporal_gemm.cu)
//     PRO_SPATIAL  x[r,(u+c)%V,c] * (tanh(mask)+1)            model/shift_gcn.py:123-129
//     PRO_LERP     temporal shift of BN(h), zero padded       model/shift_gcn.py:66-68, shift_cuda_kernel.cu:12-76
//     PRO_PLAIN    rows as they are                           (1x1 conv side branches, backward data of the 1x1 conv)
//     EPI_ROT_RAW    + bias, rotate z[v,d]=y[(v-d)%V,d], store, per-(v,d) batch statistics   :131-137 (training)
//     EPI_ROT_FUSED  + bias, rotate, folded BN, + residual, ReLU                              :131-141 (eval)
//     EPI_LINEAR     + bias, optional ReLU, store                                             :69-70
//     EPI_TSHIFT     + bias, ReLU, output shift, folded BN, + residual, ReLU (eval, stride 1)  :69-73, 161-162
//   spatial_bwd.cu
//     PRO_DY x EPI_SPATIAL_BWD   BN1d-backward + inverse output rotation, * mask, inverse input rotation, + residual
//                                gradients, dMask partial sums                      autograd of model/shift_gcn.py:123-141
// The kernels exist for the skeletons of the reference's configs, num_point 25 (NTU) and 33 (MediaPipe); any other
// joint count is an error (there is no generic fallback kernel).
#include "capi_internal.h"
#include "rowgemm.h"

namespace sgcn {

enum { PRO_SPATIAL = 0, PRO_LERP = 1, PRO_PLAIN = 2, PRO_DY = 3 };
enum { EPI_ROT_RAW = 0, EPI_ROT_FUSED = 1, EPI_LINEAR = 2, EPI_SPATIAL_BWD = 3, EPI_TSHIFT = 4 };

int spatial_bwd_launch(const SgcnRowGemm& p, cudaStream_t s);            // spatial_bwd.cu
int spatial_fwd_launch(const SgcnRowGemm& p, int fused, cudaStream_t s);  // spatial_fwd.cu
int temporal_gemm_launch(const SgcnRowGemm& p, int lerp, cudaStream_t s); // temporal_gemm.cu
int temporal_unit_launch(const SgcnRowGemm& p, cudaStream_t s);           // temporal_gemm.cu

}  // namespace sgcn

extern "C" int sgcn_rowgemm(const SgcnRowGemm* pp, int pro, int epi, void* stream) {
  using namespace sgcn;
  if (!pp) return set_error("sgcn_rowgemm: null params");
  const SgcnRowGemm& p = *pp;
  if (p.V != 25 && p.V != 33) return set_error("sgcn_rowgemm: num_point must be 25 (NTU) or 33 (MediaPipe)");
  if (epi != EPI_TSHIFT && p.G != 128 / p.V) return set_error("sgcn_rowgemm: G must be 128 / num_point whole (n,t) groups per tile");
  const bool concat = pro == PRO_PLAIN && p.k0 > 0 && (p.K == 192 || p.K == 384);   // [g | x] input-gradient GEMM
  if (p.K != 64 && p.K != 128 && p.K != 256 && !concat) return set_error("sgcn_rowgemm: K must be 64, 128 or 256");
  if (p.N != 64 && p.N != 128 && p.N != 256) return set_error("sgcn_rowgemm: N must be 64, 128 or 256");
  if (p.groups < 0) return set_error("sgcn_rowgemm: negative group count");
  if (!p.in0 || !p.out || !p.wimg) return set_error("sgcn_rowgemm: null tensor");
  cudaStream_t s = (cudaStream_t)stream;
  if (pro == PRO_SPATIAL && epi == EPI_ROT_RAW) {
    if (!p.pro_a || !p.stats) return set_error("spatial fwd: null mask / stats");
    return spatial_fwd_launch(p, 0, s);
  }
  if (pro == PRO_SPATIAL && epi == EPI_ROT_FUSED) {
    if (!p.pro_a || !p.epi_a || !p.epi_b) return set_error("spatial fwd (fused): null table");
    return spatial_fwd_launch(p, 1, s);
  }
  if (pro == PRO_LERP && epi == EPI_LINEAR) {
    if (!p.pro_a || !p.pro_b || !p.pro_c || p.T < 1) return set_error("temporal fwd: null table / bad T");
    if (p.K != p.N) return set_error("temporal fwd: the 1x1 convolution of Shift_tcn has in == out channels");
    return temporal_gemm_launch(p, 1, s);
  }
  if (pro == PRO_LERP && epi == EPI_TSHIFT) {
    if (!p.pro_a || !p.pro_b || !p.pro_c || !p.res2 || !p.epi_a || !p.epi_b || p.T < 1)
      return set_error("fused temporal unit: null table / bad T");
    if (p.K != p.N) return set_error("fused temporal unit: in == out channels");
    return temporal_unit_launch(p, s);
  }
  if (pro == PRO_PLAIN && epi == EPI_LINEAR) {
    if (p.k0 > 0 && (!p.in1 || p.k0 % 64 != 0 || p.k0 >= p.K)) return set_error("plain GEMM: bad two-source split");
    return temporal_gemm_launch(p, 0, s);
  }
  if (pro == PRO_DY && epi == EPI_SPATIAL_BWD) {
    if (!p.in1 || !p.pro_a || !p.pro_b || !p.pro_c || !p.epi_a || !p.xin || !p.red0)
      return set_error("spatial bwd: null tensor / table");
    return spatial_bwd_launch(p, s);
  }
  return set_error("sgcn_rowgemm: unsupported prologue/epilogue combination");
}
