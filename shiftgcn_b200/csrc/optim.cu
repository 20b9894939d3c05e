// optim.cu -- the step after backward: ONE kernel over the flat gradient buffer that applies, per element,
//   (i)   the 1/world scale of a summed all-reduce (1 when the collective already averaged),
//   (ii)  kernel K5 of the reference (shift_cuda_kernel.cu:371-395) to the temporal-shift positions -- on the REDUCED
//         raw sums that travel behind the gradients in the same buffer: grad_ypos = sign(raw) * 0.01, or 1e-4 if raw == 0,
//   (iii) the reference's per-parameter weight decay (main.py:307-317; a per-element table here) and
//   (iv)  SGD with momentum / Nesterov exactly as torch.optim.SGD does it (main.py:319-322, 412-414):
//             d = g + wd * p;   buf = momentum * buf + d;   p -= lr * (nesterov ? d + momentum * buf : buf)
//         (torch initialises buf with the first d, which is what a zero-initialised buffer gives).
// lr, momentum and the gradient scale are read from DEVICE memory, so a captured CUDA graph follows the learning-rate
// schedule (main.py:342-351) without being re-captured.
#include "capi_internal.h"
#include "shiftgcn_b200.h"

namespace sgcn {

__global__ void __launch_bounds__(256) sgd_epilogue_kernel(float* __restrict__ param, float* __restrict__ grad,
                                                           float* __restrict__ mbuf, const float* __restrict__ wd,
                                                           const int* __restrict__ ypos_src,
                                                           const float* __restrict__ hyper, long long n, int nesterov) {
  const float lr = hyper[0], mom = hyper[1], gscale = hyper[2];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = grad[i] * gscale;
    if (ypos_src) {
      const int src = ypos_src[i];
      if (src >= 0) {
        const float raw = grad[n + src];                       // sign only: the scale cannot change it
        g = raw != 0.f ? (raw > 0.f ? 0.01f : -0.01f) : 0.0001f;
        if (raw != raw) g = raw;                               // NaN stays visible, as in the reference (raw / |raw|)
      }
    }
    grad[i] = g;                                               // p.grad (a view of this buffer) shows the final gradient
    const float p = param[i];
    const float d = fmaf(wd[i], p, g);
    float upd = d;
    if (mom != 0.f) {
      const float b = fmaf(mom, mbuf[i], d);
      mbuf[i] = b;
      upd = nesterov ? fmaf(mom, b, d) : b;
    }
    param[i] = fmaf(-lr, upd, p);
  }
}

}  // namespace sgcn

extern "C" int sgcn_sgd_epilogue(float* param, float* grad, float* momentum_buf, const float* weight_decay,
                                 const int* ypos_src, const float* hyper, long long n_param, int nesterov, void* stream) {
  using namespace sgcn;
  if (!param || !grad || !momentum_buf || !weight_decay || !hyper) return set_error("sgcn_sgd_epilogue: null pointer");
  if (n_param < 0) return set_error("sgcn_sgd_epilogue: negative size");
  if (n_param == 0) return 0;
  long long blocks = (n_param + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  sgd_epilogue_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, weight_decay, ypos_src,
                                                                         hyper, n_param, nesterov);
  return check_launch("sgd_epilogue_kernel");
}
