"""ctypes front-end of oracle/shift_oracle.c (TEST INFRASTRUCTURE ONLY).

``build()`` compiles the C restatement with gcc into ``oracle/_build/libshift_oracle.so``;
the numpy wrappers mirror the reference's ``shift_cuda.forward`` / ``shift_cuda.backward``
(model/Temporal_shift/cuda/shift_cuda.cpp:19-42) on contiguous (N, C, H, W) arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "shift_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libshift_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -shared oracle/shift_oracle.c -> oracle/_build/libshift_oracle.so"""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
           "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _suffix(dtype):
    if dtype == np.float32:
        return "f32", ctypes.c_float
    if dtype == np.float64:
        return "f64", ctypes.c_double
    raise TypeError(f"oracle shift supports float32/float64, got {dtype}")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def shift_forward(inp, xpos, ypos, stride):
    """K1.  ``ypos`` must already carry the +0.5 offset for stride != 1."""
    inp = np.ascontiguousarray(inp)
    suf, _ = _suffix(inp.dtype)
    n, c, h, w = inp.shape
    xpos = np.ascontiguousarray(xpos, dtype=inp.dtype)
    ypos = np.ascontiguousarray(ypos, dtype=inp.dtype)
    out = np.zeros((n, c, h // stride, w), dtype=inp.dtype)
    getattr(lib(), f"oracle_shift_fwd_{suf}")(_p(inp), _p(out), _p(xpos), _p(ypos),
                                               n, c, h, w, int(stride))
    return out


def shift_backward_input(grad_out, xpos, ypos, in_shape, stride):
    """K2/K3."""
    grad_out = np.ascontiguousarray(grad_out)
    suf, _ = _suffix(grad_out.dtype)
    n, c, h, w = in_shape
    xpos = np.ascontiguousarray(xpos, dtype=grad_out.dtype)
    ypos = np.ascontiguousarray(ypos, dtype=grad_out.dtype)
    gin = np.zeros((n, c, h, w), dtype=grad_out.dtype)
    getattr(lib(), f"oracle_shift_bwd_input_{suf}")(_p(grad_out), _p(gin), _p(xpos), _p(ypos),
                                                     n, c, h, w, int(stride))
    return gin


def shift_backward_pos_raw(inp, grad_out, xpos, ypos, stride):
    """K4 + mean(0)/sum(2)/sum(1): raw per-channel position gradients (before K5)."""
    inp = np.ascontiguousarray(inp)
    grad_out = np.ascontiguousarray(grad_out, dtype=inp.dtype)
    suf, _ = _suffix(inp.dtype)
    n, c, h, w = inp.shape
    xpos = np.ascontiguousarray(xpos, dtype=inp.dtype)
    ypos = np.ascontiguousarray(ypos, dtype=inp.dtype)
    gx = np.zeros((c,), dtype=inp.dtype)
    gy = np.zeros((c,), dtype=inp.dtype)
    getattr(lib(), f"oracle_shift_bwd_pos_raw_{suf}")(_p(inp), _p(grad_out), _p(xpos), _p(ypos),
                                                       n, c, h, w, int(stride), _p(gx), _p(gy))
    return gx, gy


def shift_constraint(gx, gy):
    """K5 (returns new arrays)."""
    gx = np.array(gx, copy=True)
    gy = np.array(gy, copy=True)
    suf, _ = _suffix(gy.dtype)
    getattr(lib(), f"oracle_shift_constraint_{suf}")(_p(gx), _p(gy), int(gy.shape[0]))
    return gx, gy


def shift_backward(grad_out, inp, xpos, ypos, stride):
    """Full ``shift_cuda.backward`` restatement: (grad_input, grad_xpos, grad_ypos)."""
    gin = shift_backward_input(grad_out, xpos, ypos, inp.shape, stride)
    gx, gy = shift_backward_pos_raw(inp, grad_out, xpos, ypos, stride)
    gx, gy = shift_constraint(gx, gy)
    return gin, gx, gy
