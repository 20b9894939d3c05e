// modality.cu -- input side of the path: the bone / motion / bone-motion streams of the 4-stream ensemble derived on the
// device from the joint batch (inference_pipeline.py:284-309, data_gen/gen_bone_data.py:44-58,
// data_gen/gen_motion_data.py:18-34), optionally fused with the model's input BatchNorm in inference form and the
// change to the channels-last row layout the units consume (model/shift_gcn.py:193-198).
//
//   bone[n,c,t,v,m]   = joint[n,c,t,v,m] - joint[n,c,t,parent[v],m]            (parent[v] == v: the root stays 0)
//   motion[n,c,t,v,m] = s[n,c,t+1,v,m] - s[n,c,t,v,m]  for t < T-1,  0 at t = T-1        (s = joint or bone)
//
// The subtraction order follows the reference exactly (bone first, then the frame difference of the bone), so the fp32
// results are bit-identical to the numpy code.  The tensor is tiny (5.76 MB at NTU batch 64) and stays in L2; the kernel
// is one thread per OUTPUT element so that the writes of both layouts are coalesced.
#include "capi_internal.h"
#include "shiftgcn_b200.h"

namespace sgcn {

template <bool ROWS>
__global__ void __launch_bounds__(256) input_stream_kernel(const float* __restrict__ joint, float* __restrict__ out,
                                                           const int* __restrict__ parent,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift, long long total, int C, int T,
                                                           int V, int M, int motion) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c, t, v, m;
  long long n;
  long long r = i;
  if (ROWS) {   // out index = (((n*M + m)*T + t)*V + v)*C + c
    c = (int)(r % C), r /= C;
    v = (int)(r % V), r /= V;
    t = (int)(r % T), r /= T;
    m = (int)(r % M), n = r / M;
  } else {      // out index = ((((n*C + c)*T + t)*V + v)*M + m
    m = (int)(r % M), r /= M;
    v = (int)(r % V), r /= V;
    t = (int)(r % T), r /= T;
    c = (int)(r % C), n = r / C;
  }
  const long long frame = (long long)V * M;
  const float* base = joint + ((n * C + c) * T + t) * frame;
  const int pv = parent ? __ldg(parent + v) : v;
  const bool bone = parent != nullptr;
  float val = __ldg(base + (long long)v * M + m);
  if (bone) val = val - __ldg(base + (long long)pv * M + m);
  if (motion) {
    if (t + 1 < T) {
      float nxt = __ldg(base + frame + (long long)v * M + m);
      if (bone) nxt = nxt - __ldg(base + frame + (long long)pv * M + m);
      val = nxt - val;
    } else {
      val = 0.f;
    }
  }
  if (scale) {  // data_bn in inference form: feature index (m, v, c)  (model/shift_gcn.py:196-197)
    const int f = (m * V + v) * C + c;
    val = fmaf(val, __ldg(scale + f), __ldg(shift + f));
  }
  out[i] = val;
}

}  // namespace sgcn

extern "C" int sgcn_input_stream(const float* joint, float* out, const int* parent, const float* scale,
                                 const float* shift, long long N, int C, int T, int V, int M, int motion, int rows,
                                 void* stream) {
  using namespace sgcn;
  if (N < 0 || C < 1 || T < 1 || V < 1 || M < 1) return set_error("sgcn_input_stream: bad shape");
  if (N == 0) return 0;                                         // empty batch: nothing to do (pointers may be NULL)
  if (!joint || !out) return set_error("sgcn_input_stream: null pointer");
  if ((scale == nullptr) != (shift == nullptr)) return set_error("sgcn_input_stream: scale and shift come together");
  if (joint == out && (parent || motion)) return set_error("sgcn_input_stream: in-place derivation is not possible");
  const long long total = N * C * T * V * M;
  if (total == 0) return 0;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error("sgcn_input_stream: tensor too large");
  if (rows)
    input_stream_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(joint, out, parent, scale, shift, total, C, T, V, M, motion);
  else
    input_stream_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(joint, out, parent, scale, shift, total, C, T, V, M, motion);
  return check_launch("input_stream_kernel");
}
