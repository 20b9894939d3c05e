"""Learnable fractional temporal shift: the ``Shift`` module and ``ShiftFunction`` autograd op.

Drop-in for model/Temporal_shift/cuda/shift.py:9-46 of the reference, on top of the sm_100a C ABI
(``sgcn_shift_fwd_nchw_*`` / ``sgcn_shift_bwd_nchw_*``).  Also exposes ``shift_cuda``, a namespace with the two
functions of the reference's pybind module (shift_cuda.cpp:44-47) so code written against it keeps working.

Semantics kept from the reference:
  * input is made contiguous; for stride != 1 the sampling position is ypos + 0.5 (shift.py:13-19);
  * backward returns (grad_input, grad_xpos, grad_ypos, None) where grad_xpos is an explicit zero tensor and
    grad_ypos is the sign-constrained +-0.01 / 1e-4 value of kernel K5 (shift_cuda_kernel.cu:371-395);
  * parameters are created on the GPU: xpos ~ U(-1e-8, 1e-8), ypos ~ U(-init_scale, init_scale) (shift.py:39-43).
"""
import types

import torch
from torch import nn

from . import ops

HALF_FRAME = 0.5


def _native_forward(input, xpos, ypos, stride):
    """shift_cuda.forward(input, xpos, ypos, stride) -> output"""
    if not input.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")          # CHECK_CUDA, shift_cuda.cpp:15
    if not input.is_contiguous():
        raise RuntimeError("input must be contiguous")             # CHECK_CONTIGUOUS, shift_cuda.cpp:16
    return ops.shift_forward(input, xpos, ypos, stride)


def _native_backward(grad_output, input, output, xpos, ypos, stride):
    """shift_cuda.backward(grad_output, input, output, xpos, ypos, stride) -> [grad_input, grad_xpos, grad_ypos]"""
    for name, t in (("grad_output", grad_output), ("output", output)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")       # shift_cuda.cpp:33-34
    return ops.shift_backward(grad_output, input.contiguous(), output, xpos, ypos, stride)


shift_cuda = types.SimpleNamespace(forward=_native_forward, backward=_native_backward)


class ShiftFunction(torch.autograd.Function):
    """``ShiftFunction.apply(input, xpos, ypos, stride)``"""

    @staticmethod
    def forward(ctx, input, xpos, ypos, stride=1):
        src = input.contiguous()
        pos_y = ypos if stride == 1 else ypos + HALF_FRAME
        out = _native_forward(src, xpos, pos_y, stride)
        ctx.stride = stride
        ctx.save_for_backward(src, out, xpos, pos_y)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        src, out, xpos, pos_y = ctx.saved_tensors
        d_input, d_xpos, d_ypos = _native_backward(grad_output.contiguous(), src, out, xpos, pos_y, ctx.stride)
        return d_input, d_xpos, d_ypos, None


class Shift(nn.Module):
    """``Shift(channel, stride, init_scale=3)`` with parameters ``xpos`` and ``ypos`` of shape (channel,)."""

    def __init__(self, channel, stride, init_scale=3):
        super().__init__()
        self.stride = stride
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.xpos = nn.Parameter(torch.empty(channel, device=device).uniform_(-1e-8, 1e-8))
        self.ypos = nn.Parameter(torch.empty(channel, device=device).uniform_(-init_scale, init_scale))

        self._load_generation = 0

    def _load_from_state_dict(self, *args, **kwargs):
        self._load_generation += 1          # Shift_tcn re-validates its cached "xpos is 0" decision (modules.py)
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, input):
        return ShiftFunction.apply(input, self.xpos, self.ypos, self.stride)
