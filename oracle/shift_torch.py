"""Vectorised torch restatement of the reference's temporal-shift op (TEST INFRASTRUCTURE ONLY).

Follows model/Temporal_shift/cuda/shift_cuda_kernel.cu (K1 :12-76, K2 :79-152, K3 :156-256,
K4 :278-363, reductions :501-509, K5 :371-395) and the autograd glue in
model/Temporal_shift/cuda/shift.py:9-46.  Works on any device torch supports, in fp32 or fp64;
the CPU fp64 run is the arbiter used by the parity tests.
"""
import torch
from torch import nn


def _floor_pos(p):
    # the reference applies floorf() to the position, i.e. floors the fp32 value (kernel.cu:49-52)
    return torch.floor(p.to(torch.float32)).to(torch.int64)


def _sample(src, rows, cols, row_ok=None):
    """src (N,C,Hs,W); rows (C,Ho) int64 row per output row; cols (C,W) int64 -> (N,C,Ho,W), zero padded."""
    n, c, hs, w = src.shape
    ho = rows.shape[1]
    ok_r = (rows >= 0) & (rows < hs)
    if row_ok is not None:
        ok_r = ok_r & row_ok
    ok_c = (cols >= 0) & (cols < w)
    rr = rows.clamp(0, hs - 1)[None, :, :, None].expand(n, c, ho, w)
    cc = cols.clamp(0, w - 1)[None, :, None, :].expand(n, c, ho, w)
    t = src.gather(2, rr).gather(3, cc)
    ok = ok_r[None, :, :, None] & ok_c[None, :, None, :]
    return t * ok.to(src.dtype)


def _taps(src, x, y, out_rows, stride_in, top_stride=1):
    """Four bilinear taps of ``src`` at rows out_row*stride_in + floor(y) (+1), cols w + floor(x) (+1).

    ``top_stride`` > 1 selects the strided-adjoint addressing of K3: a bottom-row coordinate r maps
    to row r/top_stride of ``src`` only when r % top_stride == 0.
    """
    c = x.shape[0]
    w = src.shape[3]
    x1 = _floor_pos(x)
    y1 = _floor_pos(y)
    dx = (x - x1.to(x.dtype))[None, :, None, None]
    dy = (y - y1.to(y.dtype))[None, :, None, None]
    hh = torch.arange(out_rows, device=src.device)[None, :] * stride_in + y1[:, None]  # (C,Ho)
    ww = torch.arange(w, device=src.device)[None, :] + x1[:, None]                      # (C,W)

    def grab(r, q):
        if top_stride == 1:
            return _sample(src, r, q)
        ok = torch.remainder(r, top_stride) == 0
        return _sample(src, torch.div(r, top_stride, rounding_mode="floor"), q, row_ok=ok)

    q11 = grab(hh, ww)
    q21 = grab(hh, ww + 1)
    q12 = grab(hh + 1, ww)
    q22 = grab(hh + 1, ww + 1)
    return q11, q21, q12, q22, dx, dy


def shift_forward(inp, xpos, ypos, stride):
    """K1: ``ypos`` already carries the +0.5 offset when stride != 1."""
    h = inp.shape[2]
    q11, q21, q12, q22, dx, dy = _taps(inp, xpos, ypos, h // stride, stride)
    return q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy + q22 * dx * dy


def shift_backward_input(grad_out, xpos, ypos, in_h, stride):
    """K2 (stride 1) / K3 (strided): the interpolation of grad_out at the negated positions."""
    q11, q21, q12, q22, dx, dy = _taps(grad_out, -xpos, -ypos, in_h, 1, top_stride=stride)
    return q11 * (1 - dx) * (1 - dy) + q21 * dx * (1 - dy) + q12 * (1 - dx) * dy + q22 * dx * dy


def shift_backward_pos_raw(inp, grad_out, xpos, ypos, stride):
    """K4 + at::mean(0) / at::sum(2) / at::sum(1): raw (pre-K5) per-channel position gradients."""
    h = inp.shape[2]
    q11, q21, q12, q22, dx, dy = _taps(inp, xpos, ypos, h // stride, stride)
    val_x = (1 - dy) * (q21 - q11) + dy * (q22 - q12)
    val_y = (1 - dx) * (q12 - q11) + dx * (q22 - q21)
    gx = (val_x * grad_out).mean(0).sum(2).sum(1)
    gy = (val_y * grad_out).mean(0).sum(2).sum(1)
    return gx, gy


def shift_constraint(gx, gy):
    """K5: sign(gy) * 0.01, or 1e-4 where gy == 0; gx * 0."""
    dr = torch.sqrt(gy * gy)
    nz = dr != 0
    safe = torch.where(nz, dr, torch.ones_like(dr))
    out_y = torch.where(nz, gy / safe * 0.01, torch.full_like(gy, 0.0001))
    out_x = torch.where(nz, gx / safe * 0.0, torch.zeros_like(gx))
    return out_x, out_y


# When set to a dict, every backward records its raw (pre-K5) position sums under id(xpos); the golden
# generator uses it to store the magnitudes that decide where the K5 sign is numerically meaningful.
RAW_POS_LOG = None
# oracle/build_ref_ext.load(): when set, CUDA inputs go through the REFERENCE's compiled shift_cuda kernels (the
# "reference PyTorch + custom-shift-CUDA on one B200" baseline of bench.py --impl reference-gpu)
REF_EXT = None


class OracleShiftFunction(torch.autograd.Function):
    """Restatement of cuda/shift.py:9-30 (ShiftFunction) on top of the functions above."""

    @staticmethod
    def forward(ctx, inp, xpos, ypos, stride=1):
        inp = inp.contiguous()
        if stride != 1:
            ypos = ypos + 0.5
        ctx.stride = stride
        if REF_EXT is not None and inp.is_cuda:          # the reference's own kernels, called as cuda/shift.py:12-23 does
            out = REF_EXT.forward(inp, xpos, ypos, stride)
            ctx.save_for_backward(inp, out, xpos, ypos)
            ctx.ref_ext = True
            return out
        out = shift_forward(inp, xpos, ypos, stride)
        ctx.save_for_backward(inp, xpos, ypos)
        ctx.ref_ext = False
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_output = grad_output.contiguous()
        if ctx.ref_ext:                                  # cuda/shift.py:26-30
            inp, out, xpos, ypos = ctx.saved_tensors
            gin, gx, gy = REF_EXT.backward(grad_output, inp, out, xpos, ypos, ctx.stride)
            return gin, gx, gy, None
        inp, xpos, ypos = ctx.saved_tensors
        gin = shift_backward_input(grad_output, xpos, ypos, inp.shape[2], ctx.stride)
        gx, gy = shift_backward_pos_raw(inp, grad_output, xpos, ypos, ctx.stride)
        if RAW_POS_LOG is not None:
            RAW_POS_LOG[id(xpos)] = (gx.detach().clone(), gy.detach().clone())
        gx, gy = shift_constraint(gx, gy)
        return gin, gx, gy, None


class OracleShift(nn.Module):
    """Restatement of cuda/shift.py:32-46 (Shift) without the hard-coded device='cuda'."""

    def __init__(self, channel, stride, init_scale=3):
        super().__init__()
        self.stride = stride
        self.xpos = nn.Parameter(torch.zeros(channel))
        self.ypos = nn.Parameter(torch.zeros(channel))
        self.xpos.data.uniform_(-1e-8, 1e-8)
        self.ypos.data.uniform_(-init_scale, init_scale)

    def forward(self, inp):
        return OracleShiftFunction.apply(inp, self.xpos, self.ypos, self.stride)
