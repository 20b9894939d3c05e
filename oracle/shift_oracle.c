/*
 * oracle/shift_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar CPU restatement of the reference's temporal-shift CUDA op
 * (model/Temporal_shift/cuda/shift_cuda_kernel.cu), one output element at a time,
 * in both fp32 and fp64.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call into this file.
 *
 * Parity status: the reference ships no golden vectors for this op ("parity unpinned" by the
 * reference's own tests).  This restatement is pinned (a) here against an independent
 * autograd formulation (tests/test_oracle_shift.py) and (b) on the GPU box against the
 * reference's own kernels compiled into oracle/_ref/ (tests/test_gpu_shift_op.py).
 *
 * Tensor layout everywhere: contiguous (N, C, H, W); output height Ho = H / stride.
 * "ypos" is the already-offset position (the +0.5 for stride != 1 is applied by the caller,
 * reference cuda/shift.py:14-19).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* The reference takes floorf() of the position even in the double instantiation
 * (shift_cuda_kernel.cu:49-52, 113-116, 195-198, 325-328). */
static int floor_pos_f(float p) { return (int)floorf(p); }
static int floor_pos_d(double p) { return (int)floorf((float)p); }

#define DEFINE_ORACLE(SUF, T, FLOORPOS)                                                          \
                                                                                                 \
  /* zero-padded fetch from one (n,c) plane of height hh, width ww */                            \
  static T fetch_##SUF(const T *plane, int hh, int ww, int r, int q) {                           \
    if (r < 0 || q < 0 || r >= hh || q >= ww) return (T)0;                                       \
    return plane[(size_t)r * ww + q];                                                            \
  }                                                                                              \
                                                                                                 \
  /* K1: shift_cuda_kernel.cu:12-76 (host side :405-431).                                        \
   * out[n,c,h,w] = bilinear sample of in[n,c] at (h*stride + ypos[c], w + xpos[c]). */          \
  ORACLE_API void oracle_shift_fwd_##SUF(const T *in, T *out, const T *xpos, const T *ypos,      \
                                         int N, int C, int H, int W, int stride) {               \
    const int Ho = H / stride;                                                                   \
    for (int n = 0; n < N; ++n)                                                                  \
      for (int c = 0; c < C; ++c) {                                                              \
        const T *plane = in + ((size_t)n * C + c) * H * W;                                       \
        T *oplane = out + ((size_t)n * C + c) * Ho * W;                                          \
        const T x = xpos[c], y = ypos[c];                                                        \
        const int x1 = FLOORPOS(x), y1 = FLOORPOS(y);                                            \
        const T dx = x - x1, dy = y - y1;                                                        \
        for (int h = 0; h < Ho; ++h)                                                             \
          for (int w = 0; w < W; ++w) {                                                          \
            const int r = h * stride + y1, q = w + x1;                                           \
            const T q11 = fetch_##SUF(plane, H, W, r, q);                                        \
            const T q21 = fetch_##SUF(plane, H, W, r, q + 1);                                    \
            const T q12 = fetch_##SUF(plane, H, W, r + 1, q);                                    \
            const T q22 = fetch_##SUF(plane, H, W, r + 1, q + 1);                                \
            oplane[(size_t)h * W + w] =                                                          \
                q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy +          \
                q22 * dx * dy;                                                                   \
          }                                                                                      \
      }                                                                                          \
  }                                                                                              \
                                                                                                 \
  /* one tap of the strided adjoint (shift_cuda_kernel.cu:203-248): the bottom-row coordinate    \
   * r contributes top row r/stride only when r % stride == 0 (C semantics for negatives). */    \
  static T fetch_top_##SUF(const T *plane, int ho, int ww, int r, int q, int stride) {           \
    if (r % stride != 0) return (T)0;                                                            \
    return fetch_##SUF(plane, ho, ww, r / stride, q);                                            \
  }                                                                                              \
                                                                                                 \
  /* K2 (stride 1, :79-152) and K3 (stride 2, :156-256): gradient w.r.t. the input; the same     \
   * interpolation evaluated on grad_out at the negated positions (:108-109, :191-192). */       \
  ORACLE_API void oracle_shift_bwd_input_##SUF(const T *gout, T *gin, const T *xpos,             \
                                               const T *ypos, int N, int C, int H, int W,        \
                                               int stride) {                                     \
    const int Ho = H / stride;                                                                   \
    for (int n = 0; n < N; ++n)                                                                  \
      for (int c = 0; c < C; ++c) {                                                              \
        const T *gplane = gout + ((size_t)n * C + c) * Ho * W;                                   \
        T *iplane = gin + ((size_t)n * C + c) * H * W;                                           \
        const T x = -xpos[c], y = -ypos[c];                                                      \
        const int x1 = FLOORPOS(x), y1 = FLOORPOS(y);                                            \
        const T dx = x - x1, dy = y - y1;                                                        \
        for (int h = 0; h < H; ++h)                                                              \
          for (int w = 0; w < W; ++w) {                                                          \
            const int r = h + y1, q = w + x1;                                                    \
            T q11, q21, q12, q22;                                                                \
            if (stride == 1) {                                                                   \
              q11 = fetch_##SUF(gplane, H, W, r, q);                                             \
              q21 = fetch_##SUF(gplane, H, W, r, q + 1);                                         \
              q12 = fetch_##SUF(gplane, H, W, r + 1, q);                                         \
              q22 = fetch_##SUF(gplane, H, W, r + 1, q + 1);                                     \
            } else {                                                                             \
              q11 = fetch_top_##SUF(gplane, Ho, W, r, q, stride);                                \
              q21 = fetch_top_##SUF(gplane, Ho, W, r, q + 1, stride);                            \
              q12 = fetch_top_##SUF(gplane, Ho, W, r + 1, q, stride);                            \
              q22 = fetch_top_##SUF(gplane, Ho, W, r + 1, q + 1, stride);                        \
            }                                                                                    \
            iplane[(size_t)h * W + w] =                                                          \
                q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy +          \
                q22 * dx * dy;                                                                   \
          }                                                                                      \
      }                                                                                          \
  }                                                                                              \
                                                                                                 \
  /* K4 (:278-363) followed by the ATen reductions at :501-509: mean over the batch, then sum    \
   * over W, then sum over H.  Outputs the RAW per-channel sums (before K5). */                  \
  ORACLE_API void oracle_shift_bwd_pos_raw_##SUF(const T *in, const T *gout, const T *xpos,      \
                                                 const T *ypos, int N, int C, int H, int W,      \
                                                 int stride, T *raw_gx, T *raw_gy) {             \
    const int Ho = H / stride;                                                                   \
    for (int c = 0; c < C; ++c) {                                                                \
      const T x = xpos[c], y = ypos[c];                                                          \
      const int x1 = FLOORPOS(x), y1 = FLOORPOS(y);                                              \
      const T dx = x - x1, dy = y - y1;                                                          \
      T acc_x = 0, acc_y = 0;                                                                    \
      for (int h = 0; h < Ho; ++h) {                                                             \
        T row_x = 0, row_y = 0;                                                                  \
        for (int w = 0; w < W; ++w) {                                                            \
          T mean_x = 0, mean_y = 0;                                                              \
          for (int n = 0; n < N; ++n) {                                                          \
            const T *plane = in + ((size_t)n * C + c) * H * W;                                   \
            const T g = gout[(((size_t)n * C + c) * Ho + h) * W + w];                            \
            const int r = h * stride + y1, q = w + x1;                                           \
            const T q11 = fetch_##SUF(plane, H, W, r, q);                                        \
            const T q21 = fetch_##SUF(plane, H, W, r, q + 1);                                    \
            const T q12 = fetch_##SUF(plane, H, W, r + 1, q);                                    \
            const T q22 = fetch_##SUF(plane, H, W, r + 1, q + 1);                                \
            const T val_x = (1 - dy) * (q21 - q11) + dy * (q22 - q12);                           \
            const T val_y = (1 - dx) * (q12 - q11) + dx * (q22 - q21);                           \
            mean_x += val_x * g;                                                                 \
            mean_y += val_y * g;                                                                 \
          }                                                                                      \
          row_x += mean_x / (T)N;                                                                \
          row_y += mean_y / (T)N;                                                                \
        }                                                                                        \
        acc_x += row_x;                                                                          \
        acc_y += row_y;                                                                          \
      }                                                                                          \
      raw_gx[c] = acc_x;                                                                         \
      raw_gy[c] = acc_y;                                                                         \
    }                                                                                            \
  }                                                                                              \
                                                                                                 \
  /* K5 (:371-395): sign-only ypos gradient of magnitude 0.01 (1e-4 when the raw sum is exactly  \
   * zero); xpos gradient is multiplied by 0.0. */                                               \
  ORACLE_API void oracle_shift_constraint_##SUF(T *gx, T *gy, int C) {                           \
    for (int c = 0; c < C; ++c) {                                                                \
      const T vx = gx[c], vy = gy[c];                                                            \
      const T dr = (T)sqrt((double)(vy * vy));                                                   \
      if (dr != 0) {                                                                             \
        gx[c] = (T)(vx / dr * 0.0);                                                              \
        gy[c] = (T)(vy / dr * 0.01);                                                             \
      } else {                                                                                   \
        gx[c] = (T)0.0;                                                                          \
        gy[c] = (T)0.0001;                                                                       \
      }                                                                                          \
    }                                                                                            \
  }

DEFINE_ORACLE(f32, float, floor_pos_f)
DEFINE_ORACLE(f64, double, floor_pos_d)
