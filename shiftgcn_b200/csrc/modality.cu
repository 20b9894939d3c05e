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

// Sliding windows over ONE sequence (inference_pipeline.py:252-281 + :284-309 per window): window w covers the frames
// start[w] .. start[w] + win - 1 of seq [C, Ttot, V, M], zero padded past the end of the sequence; the streams are
// derived from the PADDED window exactly like the reference does (the frame difference at the last real frame is
// 0 - value, at the last window frame 0).  The windows are never materialised on the host or in HBM.
template <bool ROWS>
__global__ void __launch_bounds__(256) window_stream_kernel(const float* __restrict__ seq, float* __restrict__ out,
                                                            const int* __restrict__ start, const int* __restrict__ parent,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, long long total, int C,
                                                            int Ttot, int win, int V, int M, int motion) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c, t, v, m;
  long long w, r = i;
  if (ROWS) {   // out index = (((w*M + m)*win + t)*V + v)*C + c
    c = (int)(r % C), r /= C;
    v = (int)(r % V), r /= V;
    t = (int)(r % win), r /= win;
    m = (int)(r % M), w = r / M;
  } else {      // out index = ((((w*C + c)*win + t)*V + v)*M + m
    m = (int)(r % M), r /= M;
    v = (int)(r % V), r /= V;
    t = (int)(r % win), r /= win;
    c = (int)(r % C), w = r / C;
  }
  const long long frame = (long long)V * M;
  const int f0 = __ldg(start + w) + t;                             // frame of the sequence
  const float* base = seq + ((long long)c * Ttot + f0) * frame;
  const int pv = parent ? __ldg(parent + v) : v;
  const bool bone = parent != nullptr;
  auto value = [&](const float* b, bool inside) -> float {
    if (!inside) return 0.f;
    float x = __ldg(b + (long long)v * M + m);
    if (bone) x = x - __ldg(b + (long long)pv * M + m);
    return x;
  };
  float val = value(base, f0 < Ttot);
  if (motion) val = (t + 1 < win) ? value(base + frame, f0 + 1 < Ttot) - val : 0.f;
  if (scale) {
    const int f = (m * V + v) * C + c;
    val = fmaf(val, __ldg(scale + f), __ldg(shift + f));
  }
  out[i] = val;
}

// fall score of every window = softmax(ensemble logits)[cls] in fp64 (inference_pipeline.py:358-360 works on a float64
// accumulator), then the per-frame mean over the windows that cover a frame with REAL data (:377-386)
__global__ void window_scores_kernel(const float* __restrict__ logits, double* __restrict__ score, int W, int K, int cls) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const float* l = logits + (long long)w * K;
  double mx = (double)l[0];
  for (int k = 1; k < K; ++k) mx = fmax(mx, (double)l[k]);
  double sum = 0.0;
  for (int k = 0; k < K; ++k) sum += exp((double)l[k] - mx);
  score[w] = exp((double)l[cls] - mx) / sum;
}

__global__ void frame_aggregate_kernel(const double* __restrict__ score, const int* __restrict__ start,
                                       const int* __restrict__ real, double* __restrict__ per_frame, int W, int total) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= total) return;
  double s = 0.0, n = 0.0;
  for (int w = 0; w < W; ++w) {
    const int a = __ldg(start + w);
    if (f >= a && f < a + __ldg(real + w)) s += score[w], n += 1.0;
  }
  per_frame[f] = s / fmax(n, 1.0);
}

// feeders/tools.py:58-101 random_move on the device, for a whole batch [N, C, T, V, M] in place: the frames of sample n
// are rotated by a(t), scaled by s(t) and translated by (tx(t), ty(t)) in the x/y channels (0, 1); a, s, tx, ty are the
// piecewise np.linspace interpolations between the K+1 node values the HOST drew (np.random.choice stays on the host so
// that a seeded run reproduces the reference's augmentation).  fp64 arithmetic like the numpy code, one rounding to fp32.
__global__ void __launch_bounds__(256) random_move_kernel(float* __restrict__ data, const double* __restrict__ vals,
                                                          const int* __restrict__ node, long long N, int C, int T, int VM,
                                                          int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (n, t, j): one x/y pair
  if (i >= N * T * VM) return;
  const int j = (int)(i % VM);
  const int t = (int)((i / VM) % T);
  const long long n = i / ((long long)VM * T);
  int seg = 0;
  while (seg + 1 < K && t >= __ldg(node + seg + 1)) ++seg;         // node[seg] <= t < node[seg + 1]
  const int t0 = __ldg(node + seg), num = __ldg(node + seg + 1) - t0, k = t - t0;
  const double* v = vals + (n * 4) * (K + 1) + seg;                // vals[n][param][node]
  auto lin = [&](int prm) -> double {                              // np.linspace(a, b, num)[k]
    const double a = v[prm * (K + 1)], b = v[prm * (K + 1) + 1];
    if (num > 1 && k == num - 1) return b;
    const double step = num > 1 ? (b - a) / (double)(num - 1) : 0.0;
    return (double)k * step + a;
  };
  const double ang = lin(0) * 3.141592653589793 / 180.0, sc = lin(1), tx = lin(2), ty = lin(3);
  const double cs = cos(ang) * sc, sn = sin(ang) * sc;
  float* px = data + ((n * C + 0) * T + t) * VM + j;
  float* py = data + ((n * C + 1) * T + t) * VM + j;
  const double x = (double)*px, y = (double)*py;
  *px = (float)(cs * x + (-sn) * y + tx);
  *py = (float)(sn * x + cs * y + ty);
}

// Input BatchNorm of the model in TRAINING mode (model/shift_gcn.py:196-198: BatchNorm1d over the M*V*C features of
// the (N, M*V*C, T) view).  Statistics of feature (m, v, c) over (n, t), read in the input layout [N, C, T, V, M]: a block
// owns channel c and a chunk of (n, t) pairs, thread j the (v, m) column j of the contiguous V*M floats of a frame.
// stats[f][2] += {sum x, sum x^2} with f = (m*V + v)*C + c (fp64); sgcn_bn_fwd_finalize turns them into scale / shift
// (and updates the running statistics), which sgcn_input_stream applies together with the layout change.
__global__ void __launch_bounds__(128) data_bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, long long N,
                                                            int C, int T, int V, int M, int per) {
  const int c = blockIdx.y, j = threadIdx.x, VM = V * M;
  if (j >= VM) return;
  const long long first = (long long)blockIdx.x * per, last = min(first + per, N * T);
  double s1 = 0.0, s2 = 0.0;
  for (long long i = first; i < last; ++i) {                       // i = n*T + t
    const long long n = i / T, t = i - n * T;
    const double v = (double)__ldg(x + ((n * C + c) * T + t) * VM + j);
    s1 += v, s2 += v * v;
  }
  const int v = j / M, m = j - v * M;
  const size_t f = ((size_t)m * V + v) * C + c;
  atomicAdd(stats + 2 * f, s1);
  atomicAdd(stats + 2 * f + 1, s2);
}

// backward sums of the same BatchNorm: g = gradient wrt its output in the ROW layout [(N*M), T, V, C];
// sums[f][2] += {sum g, sum g * xhat},  xhat = (x - mean_f) * invstd_f.  A block owns person m and a chunk of (n, t),
// thread j the (v, c) column j of the contiguous V*C floats of a row group.
__global__ void __launch_bounds__(128) data_bn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ invstd,
                                                          double* __restrict__ sums, long long N, int C, int T, int V, int M,
                                                          int per) {
  const int m = blockIdx.y, j = threadIdx.x, VC = V * C;
  if (j >= VC) return;
  const int v = j / C, c = j - v * C;
  const size_t f = ((size_t)m * V + v) * C + c;
  const double mu = (double)__ldg(mean + f), is = (double)__ldg(invstd + f);
  const long long first = (long long)blockIdx.x * per, last = min(first + per, N * T);
  double s1 = 0.0, s2 = 0.0;
  for (long long i = first; i < last; ++i) {
    const long long n = i / T, t = i - n * T;
    const double gv = (double)__ldg(g + ((n * M + m) * T + t) * VC + j);
    const double xv = (double)__ldg(x + (((n * C + c) * T + t) * V + v) * M + m);
    s1 += gv, s2 += gv * (xv - mu) * is;
  }
  atomicAdd(sums + 2 * f, s1);
  atomicAdd(sums + 2 * f + 1, s2);
}

}  // namespace sgcn

extern "C" int sgcn_data_bn_stats(const float* x, double* stats, long long N, int C, int T, int V, int M, void* stream) {
  using namespace sgcn;
  if (N < 0 || C < 1 || T < 1 || V < 1 || M < 1 || V * M > 128) return set_error("sgcn_data_bn_stats: bad shape (V*M <= 128)");
  if (N == 0) return 0;
  if (!x || !stats) return set_error("sgcn_data_bn_stats: null pointer");
  const long long pairs = N * T;
  const int chunks = (int)(pairs < 256 ? pairs : 256), per = (int)((pairs + chunks - 1) / chunks);
  data_bn_stats_kernel<<<dim3((unsigned)((pairs + per - 1) / per), C), 128, 0, (cudaStream_t)stream>>>(x, stats, N, C, T, V, M, per);
  return check_launch("data_bn_stats_kernel");
}

extern "C" int sgcn_data_bn_bwd(const float* g, const float* x, const float* mean, const float* invstd, double* sums,
                                long long N, int C, int T, int V, int M, void* stream) {
  using namespace sgcn;
  if (N < 0 || C < 1 || T < 1 || V < 1 || M < 1 || V * C > 128) return set_error("sgcn_data_bn_bwd: bad shape (V*C <= 128)");
  if (N == 0) return 0;
  if (!g || !x || !mean || !invstd || !sums) return set_error("sgcn_data_bn_bwd: null pointer");
  const long long pairs = N * T;
  const int chunks = (int)(pairs < 256 ? pairs : 256), per = (int)((pairs + chunks - 1) / chunks);
  data_bn_bwd_kernel<<<dim3((unsigned)((pairs + per - 1) / per), M), 128, 0, (cudaStream_t)stream>>>(g, x, mean, invstd, sums, N, C, T, V, M, per);
  return check_launch("data_bn_bwd_kernel");
}

extern "C" int sgcn_random_move(float* data, const double* vals, const int* node, long long N, int C, int T, int V, int M,
                                int K, void* stream) {
  using namespace sgcn;
  if (N < 0 || C < 2 || T < 1 || V < 1 || M < 1 || K < 1) return set_error("sgcn_random_move: bad shape (needs >= 2 channels)");
  if (N == 0) return 0;
  if (!data || !vals || !node) return set_error("sgcn_random_move: null pointer");
  const long long total = N * T * V * M;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error("sgcn_random_move: tensor too large");
  random_move_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(data, vals, node, N, C, T, V * M, K);
  return check_launch("random_move_kernel");
}

extern "C" int sgcn_window_stream(const float* seq, float* out, const int* start, const int* parent, const float* scale,
                                  const float* shift, long long W, int C, int Ttot, int win, int V, int M, int motion,
                                  int rows, void* stream) {
  using namespace sgcn;
  if (W < 0 || C < 1 || Ttot < 1 || win < 1 || V < 1 || M < 1) return set_error("sgcn_window_stream: bad shape");
  if (W == 0) return 0;
  if (!seq || !out || !start) return set_error("sgcn_window_stream: null pointer");
  if ((scale == nullptr) != (shift == nullptr)) return set_error("sgcn_window_stream: scale and shift come together");
  const long long total = W * C * win * V * M;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_error("sgcn_window_stream: tensor too large");
  if (rows)
    window_stream_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seq, out, start, parent, scale, shift, total, C, Ttot, win, V, M, motion);
  else
    window_stream_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seq, out, start, parent, scale, shift, total, C, Ttot, win, V, M, motion);
  return check_launch("window_stream_kernel");
}

extern "C" int sgcn_window_scores(const float* logits, const int* start, const int* real, double* score,
                                  double* per_frame, int W, int num_class, int cls, int total_frames, void* stream) {
  using namespace sgcn;
  if (W < 0 || num_class < 1 || cls < 0 || cls >= num_class || total_frames < 0) return set_error("sgcn_window_scores: bad shape");
  if (W == 0 && total_frames == 0) return 0;
  if ((W > 0 && !score) || (total_frames > 0 && (!per_frame || (W > 0 && (!start || !real)))))
    return set_error("sgcn_window_scores: null pointer");
  if (W > 0 && logits) window_scores_kernel<<<(W + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, score, W, num_class, cls);
  if (total_frames > 0)
    frame_aggregate_kernel<<<(total_frames + 255) / 256, 256, 0, (cudaStream_t)stream>>>(score, start, real, per_frame, W, total_frames);
  return check_launch("window_scores_kernel");
}

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
