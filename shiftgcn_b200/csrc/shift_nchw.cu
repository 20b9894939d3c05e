// shift_nchw.cu -- the stand-alone learnable fractional shift on contiguous (N, C, H, W) tensors.
//
// Replaces the reference extension model/Temporal_shift/cuda/shift_cuda_kernel.cu:
//   forward            K1  (:12-76,  host :405-431)
//   grad wrt input     K2 / K3 (:79-152, :156-256)
//   grad wrt positions K4 + at::mean/sum (:278-363, :480-509) and K5 (:371-395)
// Differences by design: outputs are fully written (no at::zeros pre-fill), the position gradient is
// reduced in-kernel (warp shuffles + one double atomic per warp; no (N,C,H,W) temporaries), indices
// are 64-bit, and everything runs on the caller's stream.
#include "capi_internal.h"
#include "common.cuh"

namespace sgcn {

template <typename T>
__device__ __forceinline__ int floor_pos(T p) {
  return (int)floorf((float)p);  // the reference floors the fp32 value in both instantiations
}

template <typename T>
__device__ __forceinline__ T tap(const T* __restrict__ plane, int hh, int ww, int r, int q) {
  return (r >= 0 && q >= 0 && r < hh && q < ww) ? plane[(size_t)r * ww + q] : (T)0;
}

template <typename T>
__device__ __forceinline__ T tap_top(const T* __restrict__ plane, int ho, int ww, int r, int q, int stride) {
  if (r % stride != 0) return (T)0;
  return tap(plane, ho, ww, r / stride, q);
}

// one thread per OUTPUT element; blockIdx.x enumerates (n, c) planes, blockIdx.y chunks of the plane
template <typename T>
__global__ void __launch_bounds__(256) shift_fwd_nchw_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                             const T* __restrict__ xpos, const T* __restrict__ ypos,
                                                             int C, int H, int W, int stride) {
  const int Ho = H / stride;
  const size_t plane_id = blockIdx.x;
  const int c = (int)(plane_id % C);
  const T x = xpos[c], y = ypos[c];
  const int x1 = floor_pos(x), y1 = floor_pos(y);
  const T dx = x - (T)x1, dy = y - (T)y1;
  const T* plane = in + plane_id * H * W;
  T* oplane = out + plane_id * Ho * W;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < Ho * W; i += gridDim.y * blockDim.x) {
    const int h = i / W, w = i - h * W;
    const int r = h * stride + y1, q = w + x1;
    const T q11 = tap(plane, H, W, r, q), q21 = tap(plane, H, W, r, q + 1);
    const T q12 = tap(plane, H, W, r + 1, q), q22 = tap(plane, H, W, r + 1, q + 1);
    oplane[i] = q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy + q22 * dx * dy;
  }
}

// one thread per INPUT element (gather form of the adjoint)
template <typename T>
__global__ void __launch_bounds__(256) shift_bwd_input_nchw_kernel(const T* __restrict__ gout, T* __restrict__ gin,
                                                                   const T* __restrict__ xpos,
                                                                   const T* __restrict__ ypos, int C, int H, int W,
                                                                   int stride) {
  const int Ho = H / stride;
  const size_t plane_id = blockIdx.x;
  const int c = (int)(plane_id % C);
  const T x = -xpos[c], y = -ypos[c];
  const int x1 = floor_pos(x), y1 = floor_pos(y);
  const T dx = x - (T)x1, dy = y - (T)y1;
  const T* gplane = gout + plane_id * Ho * W;
  T* iplane = gin + plane_id * H * W;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < H * W; i += gridDim.y * blockDim.x) {
    const int h = i / W, w = i - h * W;
    const int r = h + y1, q = w + x1;
    T q11, q21, q12, q22;
    if (stride == 1) {
      q11 = tap(gplane, H, W, r, q), q21 = tap(gplane, H, W, r, q + 1);
      q12 = tap(gplane, H, W, r + 1, q), q22 = tap(gplane, H, W, r + 1, q + 1);
    } else {
      q11 = tap_top(gplane, Ho, W, r, q, stride), q21 = tap_top(gplane, Ho, W, r, q + 1, stride);
      q12 = tap_top(gplane, Ho, W, r + 1, q, stride), q22 = tap_top(gplane, Ho, W, r + 1, q + 1, stride);
    }
    iplane[i] = q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy + q22 * dx * dy;
  }
}

// raw position gradients: acc[c] += sum over the block's share of (n, h, w); acc is double [2][C] (x then y)
template <typename T>
__global__ void __launch_bounds__(256) shift_bwd_pos_nchw_kernel(const T* __restrict__ in, const T* __restrict__ gout,
                                                                 const T* __restrict__ xpos,
                                                                 const T* __restrict__ ypos, double* __restrict__ acc,
                                                                 int C, int H, int W, int stride) {
  const int Ho = H / stride;
  const size_t plane_id = blockIdx.x;
  const int c = (int)(plane_id % C);
  const T x = xpos[c], y = ypos[c];
  const int x1 = floor_pos(x), y1 = floor_pos(y);
  const T dx = x - (T)x1, dy = y - (T)y1;
  const T* plane = in + plane_id * H * W;
  const T* gplane = gout + plane_id * Ho * W;
  double sx = 0.0, sy = 0.0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < Ho * W; i += gridDim.y * blockDim.x) {
    const int h = i / W, w = i - h * W;
    const int r = h * stride + y1, q = w + x1;
    const T q11 = tap(plane, H, W, r, q), q21 = tap(plane, H, W, r, q + 1);
    const T q12 = tap(plane, H, W, r + 1, q), q22 = tap(plane, H, W, r + 1, q + 1);
    const T g = gplane[i];
    sx += (double)(((1 - dy) * (q21 - q11) + dy * (q22 - q12)) * g);
    sy += (double)(((1 - dx) * (q12 - q11) + dx * (q22 - q21)) * g);
  }
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc + c, sx);
    atomicAdd(acc + C + c, sy);
  }
}

// K5 on the reduced sums (acc holds SUMS over the batch; the reference takes the mean over N first).
// Optionally exports the raw means, then clears acc so the scratch buffer can be reused.
template <typename T>
__global__ void shift_constraint_kernel(double* __restrict__ acc, T* __restrict__ gx, T* __restrict__ gy,
                                        T* __restrict__ raw_out, int C, double inv_n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const T vx = (T)(acc[c] * inv_n), vy = (T)(acc[C + c] * inv_n);
  if (raw_out) {
    raw_out[c] = vx;
    raw_out[C + c] = vy;
  }
  const T dr = (T)sqrt((double)(vy * vy));
  if (dr != (T)0) {
    gx[c] = (T)(vx / dr * 0.0);
    gy[c] = (T)(vy / dr * 0.01);
  } else {
    gx[c] = (T)0.0;
    gy[c] = (T)0.0001;
  }
  acc[c] = 0.0;
  acc[C + c] = 0.0;
}

static int check_dims(long long n, int c, int h, int w, int stride) {
  if (n < 0 || c <= 0 || h <= 0 || w <= 0) return set_error("shift: bad tensor dims");
  if (stride < 1) return set_error("shift: stride must be >= 1");
  if (n * c > 0x7fffffffLL) return set_error("shift: too many (n, c) planes");
  if ((long long)h * w > 0x7fffffffLL) return set_error("shift: plane too large");
  return 0;
}

static inline unsigned chunks_for(long long elems) {
  long long b = (elems + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 32 ? 32 : b));
}

template <typename T>
int shift_fwd_nchw(const T* in, T* out, const T* xpos, const T* ypos, long long n, int c, int h, int w, int stride,
                   cudaStream_t s) {
  if (int rc = check_dims(n, c, h, w, stride)) return rc;
  const long long planes = n * c;
  const int ho = h / stride;
  if (planes == 0 || ho == 0) return 0;
  shift_fwd_nchw_kernel<T><<<dim3((unsigned)planes, chunks_for((long long)ho * w)), 256, 0, s>>>(in, out, xpos, ypos, c,
                                                                                                h, w, stride);
  return check_launch("shift_fwd_nchw_kernel");
}

template <typename T>
int shift_bwd_nchw(const T* gout, const T* in, const T* xpos, const T* ypos, T* gin, T* gx, T* gy, T* raw_pos,
                   double* scratch, long long n, int c, int h, int w, int stride, cudaStream_t s) {
  if (int rc = check_dims(n, c, h, w, stride)) return rc;
  const long long planes = n * c;
  const int ho = h / stride;
  if (planes > 0) {
    shift_bwd_input_nchw_kernel<T><<<dim3((unsigned)planes, chunks_for((long long)h * w)), 256, 0, s>>>(
        gout, gin, xpos, ypos, c, h, w, stride);
    if (int rc = check_launch("shift_bwd_input_nchw_kernel")) return rc;
    if (ho > 0) {
      shift_bwd_pos_nchw_kernel<T><<<dim3((unsigned)planes, chunks_for((long long)ho * w)), 256, 0, s>>>(
          in, gout, xpos, ypos, scratch, c, h, w, stride);
      if (int rc = check_launch("shift_bwd_pos_nchw_kernel")) return rc;
    }
  }
  shift_constraint_kernel<T><<<(c + 127) / 128, 128, 0, s>>>(scratch, gx, gy, raw_pos, c, n > 0 ? 1.0 / (double)n : 0.0);
  return check_launch("shift_constraint_kernel");
}

}  // namespace sgcn

using sgcn::set_error;

#define SGCN_SHIFT_ENTRY(SUF, T)                                                                                       \
  extern "C" int sgcn_shift_fwd_nchw_##SUF(const T* in, T* out, const T* xpos, const T* ypos, long long n, int c,     \
                                           int h, int w, int stride, void* stream) {                                  \
    if (!in || !out || !xpos || !ypos) return set_error("sgcn_shift_fwd_nchw: null pointer");                         \
    return sgcn::shift_fwd_nchw<T>(in, out, xpos, ypos, n, c, h, w, stride, (cudaStream_t)stream);                    \
  }                                                                                                                    \
  extern "C" int sgcn_shift_bwd_nchw_##SUF(const T* grad_out, const T* in, const T* xpos, const T* ypos, T* grad_in,  \
                                           T* grad_xpos, T* grad_ypos, T* raw_pos, double* scratch, long long n,      \
                                           int c, int h, int w, int stride, void* stream) {                           \
    if (!grad_out || !in || !xpos || !ypos || !grad_in || !grad_xpos || !grad_ypos || !scratch)                       \
      return set_error("sgcn_shift_bwd_nchw: null pointer");                                                          \
    return sgcn::shift_bwd_nchw<T>(grad_out, in, xpos, ypos, grad_in, grad_xpos, grad_ypos, raw_pos, scratch, n, c,   \
                                   h, w, stride, (cudaStream_t)stream);                                               \
  }

SGCN_SHIFT_ENTRY(f32, float)
SGCN_SHIFT_ENTRY(f64, double)
