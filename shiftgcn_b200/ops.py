"""Thin tensor -> pointer wrappers over the C ABI (include/shiftgcn_b200.h).

PyTorch is plumbing only: it owns device memory and the current stream.  Every function validates device /
dtype / contiguity (the checks the reference does with AT_ASSERTM, shift_cuda.cpp:15-17, plus the ones it
omits), launches on the current stream of the tensors' device and never synchronises.
"""
import contextlib
import ctypes
import os
import threading

import torch

from . import _lib
from ._lib import SgcnRowGemm, SgcnSideBwd, SgcnSideFold, SgcnStem, SgcnTShift, SgcnTShiftBwd, SgcnTShiftInBwd, SgcnTShiftInSums, SgcnWgrad

PRO_SPATIAL, PRO_LERP, PRO_PLAIN, PRO_DY = 0, 1, 2, 3
EPI_ROT_RAW, EPI_ROT_FUSED, EPI_LINEAR, EPI_SPATIAL_BWD, EPI_TSHIFT = 0, 1, 2, 3, 4
WG_SPATIAL, WG_TEMPORAL, WG_PLAIN = 0, 1, 2


PREC_TF32, PREC_FP32 = 0, 1
_PRECISION_NAMES = {"tf32": PREC_TF32, "fp32": PREC_FP32}
# Contraction precision of the tensor-core kernels (process-wide; SGCN_PRECISION sets the start value):
#   "tf32"  operands rounded to TF32 (10-bit mantissa), fp32 accumulation -- the fast default (<= 1e-2 contract)
#   "fp32"  3xTF32: both operands split into TF32 head + tail, three MMA groups per block -- fp32-accurate
#           (<= 1e-5 contract; the reference's einsum is a true fp32 contraction, model/shift_gcn.py:131)
_precision = _PRECISION_NAMES.get(os.environ.get("SGCN_PRECISION", "tf32").lower(), PREC_TF32)


def set_precision(mode):
    """"tf32" (default) or "fp32" (3xTF32 operand split, fp32-accurate); returns the previous mode's name.  Forward and
    backward of a step should run under the same mode."""
    global _precision
    if mode not in _PRECISION_NAMES:
        raise ValueError(f"precision must be one of {sorted(_PRECISION_NAMES)}, got {mode!r}")
    prev = get_precision()
    _precision = _PRECISION_NAMES[mode]
    return prev


def get_precision():
    return "fp32" if _precision == PREC_FP32 else "tf32"


class precision:
    """``with ops.precision("fp32"): ...`` -- scoped precision of the tensor-core contractions"""

    def __init__(self, mode):
        self.mode, self.prev = mode, None

    def __enter__(self):
        self.prev = set_precision(self.mode)
        return self

    def __exit__(self, *exc):
        set_precision(self.prev)
        return False


LAUNCHES = 0          # kernels of this library enqueued so far (bench.py reports the per-step count)
PROFILE = None        # when a list: (name, start_event, end_event, algorithmic_bytes) per C-ABI call


class _StreamArg:
    """placeholder in an argument list: replaced by the current stream of the device the op's tensors live on"""


_STREAM = _StreamArg()
_tls = threading.local()          # device of the tensors seen by _p() since the last C-ABI call of this thread
_NULL_CTX = contextlib.nullcontext()


def _call(name, fn, *args):
    """One C-ABI call on the device of its tensors: every pointer handed out by _p()/_d() since the previous call must
    live on ONE device (checked there); the call runs with that device current and on ITS current stream, whatever the
    caller's current device is (the reference extension launches on the legacy default stream of the current device and
    is only correct when that happens to be the tensors' device, shift_cuda_kernel.cu:414)."""
    dev = getattr(_tls, "dev", None)
    _tls.dev = None
    if dev is None:
        dev = torch.cuda.current_device()
    ctx = _NULL_CTX if dev == torch.cuda.current_device() else torch.cuda.device(dev)
    with ctx:
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(fn(*[stream if a is _STREAM else a for a in args]), name)


def _launch(name, nkernels, algo_bytes, fn, *args):
    """call one C-ABI entry point; count its kernels; optionally bracket it with CUDA events (bench profiling pass)"""
    global LAUNCHES
    LAUNCHES += nkernels
    if PROFILE is None:
        _call(name, fn, *args)
        return
    dev = getattr(_tls, "dev", None)
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _call(name, fn, *args)
    e1.record(stream)
    PROFILE.append((name, e0, e1, algo_bytes))


def _nbytes(*tensors):
    """bytes of the distinct tensors (an identity unit passes the same tensor as input and residual: count it once)"""
    seen, total = set(), 0
    for t in tensors:
        if t is None or t.data_ptr() in seen:
            continue
        seen.add(t.data_ptr())
        total += t.numel() * t.element_size()
    return total


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _p(t, dtype=torch.float32, name="tensor"):
    """device pointer of a contiguous CUDA tensor (None -> NULL); all pointers of one call must share a device"""
    if t is None:
        return None
    cur = getattr(_tls, "dev", None)
    bad = None
    if not t.is_cuda:
        bad = f"{name} must be a CUDA tensor"
    elif t.dtype != dtype:
        bad = f"{name} must be {dtype}, got {t.dtype}"
    elif not t.is_contiguous():
        bad = f"{name} must be contiguous"
    elif cur is not None and cur != t.device.index:
        bad = f"{name} is on cuda:{t.device.index} but the other tensors of this call are on cuda:{cur}"
    if bad is not None:
        _tls.dev = None                       # the call is abandoned: do not leak its device into the next one
        raise RuntimeError(bad)
    _tls.dev = t.device.index
    return ctypes.c_void_p(t.data_ptr())


def _d(t, name="buffer"):
    return _p(t, torch.float64, name)


def device_check():
    _lib.check(_lib.load().sgcn_device_check(), "device check")


def set_traversal(snake):
    """Snake traversal (include/shiftgcn_b200.h:sgcn_set_traversal): consecutive kernels walk their tiles in opposite
    orders so that each starts on what is still in L2.  Returns the previous setting."""
    return bool(_lib.load().sgcn_set_traversal(1 if snake else 0))


def set_max_ctas(n):
    """Cap the grid of the persistent tile kernels (include/shiftgcn_b200.h:sgcn_set_max_ctas; 0 = one CTA per SM).
    Parity tests use it to put many tiles on every CTA at small sizes.  Returns the previous cap."""
    return int(_lib.load().sgcn_set_max_ctas(int(n)))


def restart_traversal():
    """restart the snake alternation (top of every step: all steps and a captured graph see the same tile orders)"""
    _lib.load().sgcn_set_traversal(_lib.traversal_mode())


def groups_per_tile(V):
    if V not in (25, 33):
        raise RuntimeError(f"shiftgcn_b200 fused kernels exist for num_point 25 (NTU) and 33 (MediaPipe), got {V}")
    return 128 // V


# ------------------------------------------------------------------------------------------------ self test
def selftest_umma(a, b, mode, K, N, M2=128):
    lib = _lib.load()
    rows = 128
    d = torch.empty(rows, N, device=a.device, dtype=torch.float32)
    _call("selftest_umma", lib.sgcn_selftest_umma, _p(a), _p(b), _p(d), mode, K, N, M2, _STREAM)
    return d


# ------------------------------------------------------------------------------------------------ stand-alone shift (NCHW)
def shift_forward(inp, xpos, ypos, stride):
    """``shift_cuda.forward`` (model/Temporal_shift/cuda/shift_cuda.cpp:19-23): ypos already carries the +0.5."""
    lib = _lib.load()
    if inp.dim() != 4:
        raise RuntimeError("input must be (N, C, H, W)")
    if inp.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("shift supports float32 / float64")
    dt = inp.dtype
    fn = lib.sgcn_shift_fwd_nchw_f32 if dt == torch.float32 else lib.sgcn_shift_fwd_nchw_f64
    n, c, h, w = inp.shape
    stride = int(stride)
    if stride < 1:
        raise RuntimeError("stride must be >= 1")
    xpos = xpos.to(dt).contiguous()
    ypos = ypos.to(dt).contiguous()
    if xpos.numel() != c or ypos.numel() != c:
        raise RuntimeError("xpos / ypos must have one entry per channel")
    out = torch.empty((n, c, h // stride, w), device=inp.device, dtype=dt)
    _count(1)
    _call("shift forward", fn, _p(inp, dt, "input"), _p(out, dt), _p(xpos, dt, "xpos"), _p(ypos, dt, "ypos"), n, c, h, w,
          stride, _STREAM)
    return out


def shift_backward(grad_output, inp, output, xpos, ypos, stride, return_raw=False):
    """``shift_cuda.backward`` (shift_cuda.cpp:25-42): returns [grad_input, grad_xpos, grad_ypos]."""
    lib = _lib.load()
    dt = inp.dtype
    if dt not in (torch.float32, torch.float64):
        raise RuntimeError("shift supports float32 / float64")
    fn = lib.sgcn_shift_bwd_nchw_f32 if dt == torch.float32 else lib.sgcn_shift_bwd_nchw_f64
    n, c, h, w = inp.shape
    stride = int(stride)
    xpos = xpos.to(dt).contiguous()
    ypos = ypos.to(dt).contiguous()
    if output is not None and not output.is_contiguous():
        raise RuntimeError("output must be contiguous")           # the reference checks it (shift_cuda.cpp:34)
    if tuple(grad_output.shape) != (n, c, h // stride, w):
        raise RuntimeError("grad_output has the wrong shape")
    grad_input = torch.empty_like(inp, memory_format=torch.contiguous_format)
    gx = torch.empty(c, device=inp.device, dtype=dt)
    gy = torch.empty(c, device=inp.device, dtype=dt)
    raw = torch.empty(2, c, device=inp.device, dtype=dt) if return_raw else None
    scratch = torch.zeros(2, c, device=inp.device, dtype=torch.float64)
    _count(3)
    _call("shift backward", fn, _p(grad_output, dt, "grad_output"), _p(inp, dt, "input"), _p(xpos, dt), _p(ypos, dt),
          _p(grad_input, dt), _p(gx, dt), _p(gy, dt), _p(raw, dt), _d(scratch), n, c, h, w, stride, _STREAM)
    if return_raw:
        return [grad_input, gx, gy], raw
    return [grad_input, gx, gy]


# ------------------------------------------------------------------------------------------------ small helpers
# ------------------------------------------------------------------------------------------------ frozen tables
# Tables that depend on parameters / buffers only (weight images, the rotated mask multipliers, eval-mode BatchNorm
# scale / shift, folded side-branch weights) are ~70 of the ~100 launches of an inference pass and ~7 % of its device time.
# Outside grad mode they are kept and reused while their sources are unchanged: same storage, same tensor version
# (in-place updates by optimizers and load_state_dict bump it), same parameter epoch.  Whatever changes parameters or
# BatchNorm buffers through raw pointers bumps the epoch (the library's own kernels do: training-mode BatchNorm finalize,
# sgcn_sgd_epilogue; nn.Module.train() / .eval() of the drop-in modules do as well); code that edits ``p.data`` by hand
# calls ``params_changed()``.
_param_epoch = 0
_frozen = {}


def param_epoch():
    """current parameter epoch (cache keys of parameter-derived decisions outside this module)"""
    return _param_epoch


def params_changed():
    """invalidate every cached parameter-derived table (see above)"""
    global _param_epoch
    _param_epoch += 1


_frozen_record = None            # list of (tensors, make) while a capture records which tables it reads


def _frozen_get(key, sources, make):
    if torch.is_grad_enabled():
        return make()
    stamp = (_param_epoch,) + tuple((t.data_ptr(), t._version) for t in sources if t is not None)
    hit = _frozen.get(key)
    if hit is not None and hit[0] == stamp:
        if _frozen_record is not None:
            _frozen_record.append((hit[1], hit[2]))
        return hit[1]
    val = make()
    if torch.cuda.is_current_stream_capturing():
        return val                                                 # graph-pool memory must not outlive its graph in the cache
    if len(_frozen) > 8192:
        _frozen.clear()
    # `make` holds the sources: their storage cannot be freed and handed to another tensor while the entry lives, so an
    # equal (data_ptr, version) stamp really means the same, unchanged tensor
    _frozen[key] = (stamp, val, make)
    return val


class frozen_recording:
    """context: collect the cached tables that are READ inside it (GraphedInference: the captured graph depends on them)"""

    def __enter__(self):
        global _frozen_record
        self._prev, _frozen_record = _frozen_record, []
        self.tables = _frozen_record
        return self

    def __exit__(self, *exc):
        global _frozen_record
        _frozen_record = self._prev
        return False


def _flat(val):
    if isinstance(val, dict):
        return [val[k] for k in sorted(val)]
    return list(val) if isinstance(val, (tuple, list)) else [val]


def frozen_refresh(tables):
    """recompute recorded tables IN PLACE (same storage) from the current parameters -- what an inference graph captured
    with them needs after the weights changed"""
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(True)                                   # no cache look-ups inside make()
    try:
        for old, make in tables:
            for a, b in zip(_flat(old), _flat(make())):
                if a is not None:
                    a.copy_(b)
    finally:
        torch.set_grad_enabled(prev)


def _src_key(*tensors):
    return tuple((t.data_ptr(), tuple(t.shape), tuple(t.stride())) if t is not None else None for t in tensors)


def weight_image(src, ld_n, ld_k, N, K):
    return _frozen_get(("wimg", _src_key(src), ld_n, ld_k, N, K, _precision), (src,),
                       lambda: _weight_image(src, ld_n, ld_k, N, K))


def _weight_image(src, ld_n, ld_k, N, K):
    """canonical TF32 image of B[n][k] = src.flatten()[n*ld_n + k*ld_k]; in "fp32" precision the head image followed
    by the tail image (what sgcn_rowgemm expects under SGCN_PREC_FP32)"""
    _count()
    lib = _lib.load()
    if _precision == PREC_FP32:
        img = torch.empty(2 * N * K, device=src.device, dtype=torch.float32)
        _call("weight image (split)", lib.sgcn_prep_weight_image_split, _p(src, name="weight"), ld_n, ld_k, N, K, _p(img), _STREAM)
        return img
    img = torch.empty(N * K, device=src.device, dtype=torch.float32)
    _call("weight image", lib.sgcn_prep_weight_image, _p(src, name="weight"), ld_n, ld_k, N, K, _p(img), _STREAM)
    return img


def mask_prepare(mask):
    """(V, C) Feature_Mask -> (tanh(mask)+1, the same table indexed by the SOURCE joint of the shift_in gather)"""
    return _frozen_get(("mask", _src_key(mask)), (mask,), lambda: _mask_prepare(mask))


def _mask_prepare(mask):
    _count()
    lib = _lib.load()
    V, C = mask.shape
    out = torch.empty(2, V, C, device=mask.device, dtype=torch.float32)
    _call("mask prepare", lib.sgcn_mask_prepare_rot, _p(mask, name="Feature_Mask"), _p(out[0]), _p(out[1]), V, C, _STREAM)
    return out[0], out[1]


def mask_grad_finalize(raw, mask):
    _count()
    lib = _lib.load()
    dm = torch.empty_like(mask)
    _call("mask grad", lib.sgcn_mask_grad_finalize, _d(raw), _p(mask), _p(dm), mask.numel(), _STREAM)
    return dm


def reduce_export(src, scale=1.0):
    _count()
    lib = _lib.load()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    _call("reduce export", lib.sgcn_reduce_export, _d(src), _p(dst), src.numel(), float(scale), _STREAM)
    return dst


def bn_fwd_finalize(stats, gamma, beta, running_mean, running_var, nbt, features, count, momentum, eps, training):
    """-> (mean, invstd, scale, shift) fp32 [features]; updates running stats / num_batches_tracked when training"""
    if training or stats is not None:
        params_changed()                                           # running statistics are written through raw pointers
        return _bn_fwd_finalize(stats, gamma, beta, running_mean, running_var, nbt, features, count, momentum, eps, training)
    return _frozen_get(("bn_eval", _src_key(gamma, beta, running_mean, running_var), features, float(eps)),
                       (gamma, beta, running_mean, running_var),
                       lambda: _bn_fwd_finalize(None, gamma, beta, running_mean, running_var, None, features, count, momentum,
                                                eps, False))


def _bn_fwd_finalize(stats, gamma, beta, running_mean, running_var, nbt, features, count, momentum, eps, training):
    _count()
    lib = _lib.load()
    dev = gamma.device if gamma is not None else running_mean.device
    out = torch.empty(4, features, device=dev, dtype=torch.float32)
    _call("bn forward finalize", lib.sgcn_bn_fwd_finalize, _d(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(nbt, torch.int64),
        _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), features, float(count), float(momentum), float(eps),
        1 if training else 0, _STREAM)
    return out[0], out[1], out[2], out[3]


def tshift_bwd_finalize(sums, gamma, invstd, C, count, n_batch, training, input_shift=False, want_raw=False):
    """-> dict(dgamma, dbeta, k1, m1, m2, gx, gy[, raw])"""
    _count()
    lib = _lib.load()
    out = torch.empty(8, C, device=gamma.device, dtype=torch.float32)
    fn = lib.sgcn_tshift_in_bwd_finalize if input_shift else lib.sgcn_tshift_bwd_finalize
    _call("tshift backward finalize", fn, _d(sums), _p(gamma), _p(invstd), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _p(out[4]),
                  _p(out[5]), _p(out[6]), _p(out[7]) if want_raw else None, C, float(count), float(n_batch),
                  1 if training else 0, _STREAM)
    return dict(dgamma=out[0], dbeta=out[1], k1=out[2], m1=out[3], m2=out[4], gx=out[5], gy=out[6], raw=out[7])


def bn1d_bwd_finalize(vd_sums, gamma, mean, invstd, V, D, count, training):
    """-> dict(dgamma, dbeta, alpha, beta, gamma, dbias)"""
    _count()
    lib = _lib.load()
    out = torch.empty(5, V * D, device=gamma.device, dtype=torch.float32)
    dbias = torch.empty(D, device=gamma.device, dtype=torch.float32)
    _call("bn1d backward finalize", lib.sgcn_bn1d_bwd_finalize, _d(vd_sums), _p(gamma), _p(mean), _p(invstd), _p(out[0]), _p(out[1]),
                                          _p(out[2]), _p(out[3]), _p(out[4]), _p(dbias), V, D, float(count),
                                          1 if training else 0, _STREAM)
    return dict(dgamma=out[0], dbeta=out[1], alpha=out[2], beta=out[3], gamma=out[4], dbias=dbias)


# ------------------------------------------------------------------------------------------------ tensor-core kernels
def rowgemm(pro, epi, *, in0, out, wimg, groups, V, K, N, T=1, in1=None, pro_a=None, pro_b=None, pro_c=None, bias=None,
            epi_a=None, epi_b=None, res=None, res2=None, res2m=None, xin=None, stats=None, red0=None, relu=0, k0=0,
            in0_gs=0, in1_gs=0, out_gs=0, accum=0):
    lib = _lib.load()
    p = SgcnRowGemm(in0=_p(in0, name="in0"), in1=_p(in1), out=_p(out, name="out"), wimg=_p(wimg), pro_a=_p(pro_a),
                    pro_b=_p(pro_b), pro_c=_p(pro_c), bias=_p(bias), epi_a=_p(epi_a), epi_b=_p(epi_b), res=_p(res),
                    res2=_p(res2), res2m=_p(res2m), xin=_p(xin), stats=_d(stats), red0=_d(red0), groups=int(groups),
                    V=V, G=groups_per_tile(V), T=int(T), K=K, N=N, relu=int(relu), k0=int(k0), in0_gs=int(in0_gs),
                    in1_gs=int(in1_gs), out_gs=int(out_gs), accum=int(accum), prec=_precision)
    if wimg.numel() != (2 if _precision == PREC_FP32 else 1) * K * N:
        raise RuntimeError("rowgemm: the weight image was prepared under a different precision mode")
    name = "rowgemm[%s/%s]" % (("spatial", "lerp", "plain", "dy")[pro], ("rot_raw", "rot_fused", "linear", "spatial_bwd", "tshift")[epi])
    nbytes = int(groups) * V * 4 * (K + N * (2 if accum else 1)) if pro == PRO_PLAIN else _nbytes(in0, in1, out, res, res2, res2m, xin)
    _launch(name, 1, nbytes, lib.sgcn_rowgemm, ctypes.byref(p), pro, epi, _STREAM)


def wgrad(mode, *, a_src, b_src, dw, groups, V, CA, CB, T=1, a_tab0=None, b_src2=None, b_tab0=None, b_tab1=None,
          b_tab2=None, a_gs=1, b_gs=1):
    lib = _lib.load()
    p = SgcnWgrad(a_src=_p(a_src), a_tab0=_p(a_tab0), b_src=_p(b_src), b_src2=_p(b_src2), b_tab0=_p(b_tab0),
                  b_tab1=_p(b_tab1), b_tab2=_p(b_tab2), dw=_p(dw), groups=int(groups), V=V, G=groups_per_tile(V),
                  T=int(T), CA=CA, CB=CB, a_gs=int(a_gs), b_gs=int(b_gs), prec=_precision)
    nbytes = int(groups) * V * 4 * (CA + CB) if mode == WG_PLAIN else _nbytes(a_src, b_src, b_src2)
    _launch("wgrad[%s]" % ("spatial", "temporal", "plain")[mode], 1, nbytes, lib.sgcn_wgrad,
            ctypes.byref(p), mode, _STREAM)


# ------------------------------------------------------------------------------------------------ first spatial unit
_STEM_F32 = ("x", "maskmul", "W", "bias", "Wd", "bd", "sc1", "sh1", "sc2", "sh2", "h", "g", "mean1", "invstd1", "mean2",
             "invstd2", "al", "be", "ga", "a2", "b2", "c2", "dx")
_STEM_F64 = ("stats_vd", "stats_r", "stats_h", "vd_sums", "r_sums", "dw_raw", "dmask_raw")


def _stem_params(groups, V, D, **t):
    kw = {k: _p(t.get(k), name=k) for k in _STEM_F32}
    kw.update({k: _d(t.get(k), name=k) for k in _STEM_F64})
    return SgcnStem(groups=int(groups), V=V, D=D, **kw)


def stem_fwd(mode, *, groups, V, D, **t):
    """l1.gcn1 forward: mode 0 batch statistics of z / of the down conv, mode 1 h = relu(BN1d(z) + BN2d(conv(x)))"""
    lib = _lib.load()
    p = _stem_params(groups, V, D, **t)
    _launch("stem_fwd[%s]" % ("stats", "apply")[mode], 1, _nbytes(t.get("x"), t.get("h") if mode == 1 else None),
            lib.sgcn_stem_fwd, ctypes.byref(p), mode, _STREAM)


def stem_bwd(mode, *, groups, V, D, **t):
    """l1.gcn1 backward: mode 0 BatchNorm backward sums, mode 1 parameter gradients + dx"""
    lib = _lib.load()
    p = _stem_params(groups, V, D, **t)
    _launch("stem_bwd[%s]" % ("stats", "apply")[mode], 1, _nbytes(t.get("x"), t.get("g"), t.get("h"), t.get("dx")),
            lib.sgcn_stem_bwd, ctypes.byref(p), mode, _STREAM)


# ------------------------------------------------------------------------------------------------ SIMT kernels
def bn_res_relu_fwd(z, res, h, scale, shift, stats_out, rows, V, D, relu=1):
    lib = _lib.load()
    _launch("bn_res_relu_fwd", 1, _nbytes(z, res, h), lib.sgcn_bn_res_relu_fwd, _p(z), _p(res), _p(h), _p(scale),
            _p(shift), _d(stats_out), int(rows), V, D, int(relu), _STREAM)


def tshift_fwd(mode, *, q, ypos_eff, n_samples, T_in, T_out, V, C, stride, res=None, out=None, scale=None, shift=None,
               stats=None, relu=0):
    """mode 0: stats [C, 2] f64 += {sum s, sum s^2};  mode 1: out = [relu](BN(s) + res); with stats [n, C] f64 the pooled
    sums over (t, v) of out are accumulated as well and out may be None (include/shiftgcn_b200.h:sgcn_tshift_fwd)"""
    lib = _lib.load()
    p = SgcnTShift(q=_p(q), res=_p(res), out=_p(out), ypos_eff=_p(ypos_eff), scale=_p(scale), shift=_p(shift),
                   stats=_d(stats), n_samples=int(n_samples), T_in=T_in, T_out=T_out, V=V, C=C, stride=stride,
                   relu=int(relu))
    _launch("tshift_fwd[%s]" % ("stats", "apply")[mode], 1, _nbytes(q, res, out), lib.sgcn_tshift_fwd, ctypes.byref(p),
            mode, _STREAM)


def tshift_bwd(mode, *, q, gy, ypos_eff, mean, invstd, n_samples, T_in, T_out, V, C, stride, y=None, relu=0, k1=None,
               m1=None, m2=None, sums=None, dpre=None, dbias=None):
    lib = _lib.load()
    p = SgcnTShiftBwd(q=_p(q), gy=_p(gy), y=_p(y), ypos_eff=_p(ypos_eff), mean=_p(mean), invstd=_p(invstd), k1=_p(k1),
                      m1=_p(m1), m2=_p(m2), sums=_d(sums), dpre=_p(dpre), dbias=_d(dbias), n_samples=int(n_samples),
                      T_in=T_in, T_out=T_out, V=V, C=C, stride=stride, relu=int(relu))
    _launch("tshift_bwd[%s]" % ("stats", "apply")[mode], 1, _nbytes(q, gy, y, dpre), lib.sgcn_tshift_bwd,
            ctypes.byref(p), mode, _STREAM)


def tshift_in_bwd(mode, *, dp, h, ypos_eff, mean, invstd, n_samples, T, V, C, scale=None, shift=None, k1=None, m1=None,
                  m2=None, z=None, zmean=None, zinvstd=None, sums=None, vd_sums=None, gh=None, relu_h=0, pos_sums=None,
                  gate=None):
    lib = _lib.load()
    p = SgcnTShiftInBwd(dp=_p(dp), h=_p(h), z=_p(z), ypos_eff=_p(ypos_eff), mean=_p(mean), invstd=_p(invstd),
                        scale=_p(scale), shift=_p(shift), k1=_p(k1), m1=_p(m1), m2=_p(m2), zmean=_p(zmean),
                        zinvstd=_p(zinvstd), sums=_d(sums), vd_sums=_d(vd_sums), gh=_p(gh), pos_sums=_d(pos_sums),
                        gate=_p(gate, torch.int32, "gate"), n_samples=int(n_samples), T=T, V=V, C=C, relu_h=int(relu_h))
    # a gated statistics pass normally returns at once (its sums came from tshift_in_bwd_sums): no algorithmic bytes
    _launch("tshift_in_bwd[%s]" % ("stats", "apply")[mode] + ("(gated)" if gate is not None else ""), 1,
            0 if gate is not None else _nbytes(dp, h, z, gh), lib.sgcn_tshift_in_bwd, ctypes.byref(p), mode, _STREAM)


def tshift_in_bwd_sums(*, dp, ypos_eff, Wt, dWt, dbt, mean, invstd, scale, shift, sums, gate, n_samples, T, V, C):
    """BN(h) backward sums from dW_t and the conv-bias gradient (include/shiftgcn_b200.h:SgcnTShiftInSums)."""
    lib = _lib.load()
    p = SgcnTShiftInSums(dp=_p(dp), ypos_eff=_p(ypos_eff), Wt=_p(Wt, name="Wt"), dWt=_p(dWt, name="dWt"), dbt=_p(dbt),
                         mean=_p(mean), invstd=_p(invstd), scale=_p(scale), shift=_p(shift), sums=_d(sums),
                         gate=_p(gate, torch.int32, "gate"), n_samples=int(n_samples), T=T, V=V, C=C)
    _launch("tshift_in_bwd[sums]", 3, 0, lib.sgcn_tshift_in_bwd_sums, ctypes.byref(p), _STREAM)


def shift_pos_finalize(pos_sums, C, n_batch, want_raw=False):
    """K5 on the reduced position-gradient sums -> (grad_xpos, grad_ypos, raw means or None)"""
    _count()
    lib = _lib.load()
    out = torch.empty(3, C, device=pos_sums.device, dtype=torch.float32)
    _call("shift position finalize", lib.sgcn_shift_pos_finalize, _d(pos_sums), _p(out[0]), _p(out[1]), _p(out[2]) if want_raw else None, C,
                                           float(n_batch), _STREAM)
    return out[0], out[1], (out[2] if want_raw else None)


def channel_stats(x, stats, rows, C):
    lib = _lib.load()
    _launch("channel_stats", 1, _nbytes(x), lib.sgcn_channel_stats, _p(x), _d(stats), int(rows), C, _STREAM)


def channel_stats_groups(x, stats, groups, V, C, gs):
    """channel sums over `groups` frames of V rows, frame g being frame g*gs of x"""
    lib = _lib.load()
    _launch("channel_stats", 1, int(groups) * V * C * 4, lib.sgcn_channel_stats_groups, _p(x), _d(stats), int(groups), V, C,
            int(gs), _STREAM)


def relu_bn1d_bwd_stats(g, h, z, zmean, zinvstd, gh, vd_sums, groups, V, C):
    lib = _lib.load()
    _launch("relu_bn1d_bwd_stats", 1, _nbytes(g, h, z, gh), lib.sgcn_relu_bn1d_bwd_stats, _p(g), _p(h), _p(z), _p(zmean),
            _p(zinvstd), _p(gh), _d(vd_sums), int(groups), V, C, _STREAM)


def relu_mask_grad(g, y):
    lib = _lib.load()
    out = torch.empty_like(g)
    _launch("relu_mask_grad", 1, _nbytes(g, y, out), lib.sgcn_relu_mask_grad, _p(g), _p(y), _p(out), g.numel(), _STREAM)
    return out


# ------------------------------------------------------------------------------------------------ input streams
def input_stream(joint, parent=None, motion=False, rows=False, scale=None, shift=None):
    """bone / motion / bone-motion stream of a joint batch (N, C, T, V, M) on the device (sgcn_input_stream).

    parent: int32 CUDA tensor [V] of 0-based parent joints or None; rows=True returns the channels-last row tensor
    (N*M, T, V, C), optionally with the input BatchNorm folded into scale / shift [M*V*C]."""
    lib = _lib.load()
    if joint.dim() != 5:
        raise RuntimeError("input_stream expects a joint batch of shape (N, C, T, V, M)")
    N, C, T, V, M = joint.shape
    if parent is not None and (parent.numel() != V or parent.dtype != torch.int32):
        raise RuntimeError("parent must be an int32 tensor with one entry per joint")
    if (scale is None) != (shift is None) or (scale is not None and (not rows or scale.numel() != M * V * C)):
        raise RuntimeError("scale / shift come together, need rows=True and M*V*C entries each")
    out = torch.empty((N * M, T, V, C) if rows else (N, C, T, V, M), device=joint.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    _launch("input_stream", 1, _nbytes(joint, out), lib.sgcn_input_stream, _p(joint, name="joint"), _p(out),
            _p(parent, torch.int32, "parent"), _p(scale), _p(shift), N, C, T, V, M, 1 if motion else 0, 1 if rows else 0,
            _STREAM)
    return out


def window_stream(seq, start, window, parent=None, motion=False, rows=False, scale=None, shift=None):
    """streams of the sliding windows of ONE sequence (C, Ttot, V, M) on the device (sgcn_window_stream): window w = frames
    start[w] .. start[w]+window-1, zero padded past the end, modality derived from the padded window like the reference
    (inference_pipeline.py:252-309).  start: int32 CUDA tensor [W].  Returns (W, C, window, V, M), or with rows=True the
    row tensor (W*M, window, V, C)."""
    lib = _lib.load()
    if seq.dim() != 4:
        raise RuntimeError("window_stream expects a sequence of shape (C, T, V, M)")
    C, Ttot, V, M = seq.shape
    if start.dtype != torch.int32 or start.dim() != 1:
        raise RuntimeError("start must be a 1-D int32 tensor of window start frames")
    if parent is not None and (parent.numel() != V or parent.dtype != torch.int32):
        raise RuntimeError("parent must be an int32 tensor with one entry per joint")
    if (scale is None) != (shift is None) or (scale is not None and (not rows or scale.numel() != M * V * C)):
        raise RuntimeError("scale / shift come together, need rows=True and M*V*C entries each")
    W = start.numel()
    out = torch.empty((W * M, window, V, C) if rows else (W, C, window, V, M), device=seq.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    _launch("window_stream", 1, _nbytes(seq, out), lib.sgcn_window_stream, _p(seq, name="seq"), _p(out),
            _p(start, torch.int32, "start"), _p(parent, torch.int32, "parent"), _p(scale), _p(shift), W, C, Ttot, window, V,
            M, 1 if motion else 0, 1 if rows else 0, _STREAM)
    return out


def window_scores(logits, start, real, total_frames, cls=1):
    """(score [W] f64, per_frame [total_frames] f64): softmax(logits)[cls] per window and its per-frame average over the
    windows covering each frame with real data (sgcn_window_scores; inference_pipeline.py:358-360, 377-386)."""
    lib = _lib.load()
    W, K = logits.shape
    dev = logits.device
    score = torch.empty(W, device=dev, dtype=torch.float64)
    per_frame = torch.empty(total_frames, device=dev, dtype=torch.float64)
    if W == 0 and total_frames == 0:
        return score, per_frame
    _launch("window_scores", 2, _nbytes(logits, per_frame), lib.sgcn_window_scores, _p(logits, name="logits"),
            _p(start, torch.int32, "start"), _p(real, torch.int32, "real"), _d(score), _d(per_frame), W, K, cls, total_frames,
            _STREAM)
    return score, per_frame


def random_move_(data, vals, node):
    """in-place feeders/tools.py:58-101 random_move of a CUDA batch (N, C, T, V, M) (sgcn_random_move).
    vals: fp64 CUDA (N, 4, K+1) node values (angle in degrees, scale, tx, ty); node: int32 CUDA [K+1] frame indices."""
    lib = _lib.load()
    if data.dim() != 5:
        raise RuntimeError("random_move_ expects a batch of shape (N, C, T, V, M)")
    N, C, T, V, M = data.shape
    K = node.numel() - 1
    if vals.shape != (N, 4, K + 1):
        raise RuntimeError(f"vals must have shape {(N, 4, K + 1)}, got {tuple(vals.shape)}")
    if N == 0:
        return data
    _launch("random_move", 1, _nbytes(data), lib.sgcn_random_move, _p(data, name="data"), _d(vals, "vals"),
            _p(node, torch.int32, "node"), N, C, T, V, M, K, _STREAM)
    return data


def data_bn_stats(x, stats):
    """stats [M*V*C, 2] f64 += {sum, sum of squares} of every input feature (m, v, c) of x (N, C, T, V, M) over (N, T)"""
    lib = _lib.load()
    N, C, T, V, M = x.shape
    _launch("data_bn[stats]", 1, _nbytes(x), lib.sgcn_data_bn_stats, _p(x, name="x"), _d(stats), N, C, T, V, M, _STREAM)


def data_bn_bwd(g_rows, x, mean, invstd, sums):
    """sums [M*V*C, 2] f64 += {sum g, sum g*xhat}: g_rows (N*M, T, V, C) = gradient wrt the data_bn output (row layout)"""
    lib = _lib.load()
    N, C, T, V, M = x.shape
    _launch("data_bn[bwd]", 1, _nbytes(g_rows, x), lib.sgcn_data_bn_bwd, _p(g_rows, name="g"), _p(x, name="x"), _p(mean),
            _p(invstd), _d(sums), N, C, T, V, M, _STREAM)


def frame_aggregate(score, start, real, total_frames):
    """per-frame mean of given window scores (fp64 [W]) -- the aggregation half of sgcn_window_scores"""
    lib = _lib.load()
    per_frame = torch.empty(total_frames, device=score.device, dtype=torch.float64)
    if total_frames == 0:
        return per_frame
    _launch("window_scores", 1, _nbytes(score, per_frame), lib.sgcn_window_scores, None, _p(start, torch.int32, "start"),
            _p(real, torch.int32, "real"), _d(score), _d(per_frame), score.numel(), 2, 1, total_frames, _STREAM)
    return per_frame


# ------------------------------------------------------------------------------------------------ conv + BN side branches
def side_fold(Wd, bd, gamma, beta, running_mean, running_var, nbt, rows, eps, momentum, training, sx_sums=None, XX=None,
              counter=None):
    """sgcn_side_fold -> dict(Wf [D,C], bf [D], mean_r [D] f64, invstd [D] f64, sx [C] f64 or None)"""
    if training:
        params_changed()                                           # running statistics are written through raw pointers
        return _side_fold(Wd, bd, gamma, beta, running_mean, running_var, nbt, rows, eps, momentum, training, sx_sums, XX,
                          counter)
    return _frozen_get(("side_fold", _src_key(Wd, bd, gamma, beta, running_mean, running_var), float(eps)),
                       (Wd, bd, gamma, beta, running_mean, running_var),
                       lambda: _side_fold(Wd, bd, gamma, beta, running_mean, running_var, None, rows, eps, momentum, False))


def _side_fold(Wd, bd, gamma, beta, running_mean, running_var, nbt, rows, eps, momentum, training, sx_sums=None, XX=None,
               counter=None):
    _count()
    lib = _lib.load()
    D, C = Wd.shape
    dev = Wd.device
    Wf = torch.empty((D, C), device=dev, dtype=torch.float32)
    bf = torch.empty(D, device=dev, dtype=torch.float32)
    stat = torch.empty((2, D), device=dev, dtype=torch.float64)
    sx = torch.empty(C, device=dev, dtype=torch.float64) if training else None
    p = SgcnSideFold(sx_sums=_d(sx_sums), XX=_p(XX), Wd=_p(Wd, name="Wd"), bd=_p(bd), gamma=_p(gamma), beta=_p(beta),
                     running_mean=_p(running_mean), running_var=_p(running_var),
                     num_batches_tracked=_p(nbt, torch.int64), Wf=_p(Wf), bf=_p(bf), mean_r=_d(stat[0]), invstd=_d(stat[1]),
                     sx=_d(sx), counter=_p(counter, torch.int32), rows=float(rows), eps=float(eps),
                     momentum=float(momentum), C=C, D=D, training=1 if training else 0)
    _call("side fold", lib.sgcn_side_fold, ctypes.byref(p), _STREAM)
    return dict(Wf=Wf, bf=bf, mean_r=stat[0], invstd=stat[1], sx=sx)


def side_bwd(P, sg, Wd, bd, gamma, invstd, mean_r, rows, training, sx=None, XX=None):
    """sgcn_side_bwd -> dict(dgamma, dbeta, dWd [D,C], dbd, Wcat [D+C,C], kvec [C])"""
    _count(2)
    lib = _lib.load()
    D, C = Wd.shape
    dev = Wd.device
    vec = torch.empty((3, D), device=dev, dtype=torch.float32)
    dWd = torch.empty((D, C), device=dev, dtype=torch.float32)
    Wcat = torch.empty((D + C, C), device=dev, dtype=torch.float32)
    kvec = torch.empty(C, device=dev, dtype=torch.float32)
    coef = torch.empty(2 * D, device=dev, dtype=torch.float64)
    p = SgcnSideBwd(P=_p(P, name="P"), sg=_p(sg, name="sg"), XX=_p(XX), sx=_d(sx), Wd=_p(Wd, name="Wd"), bd=_p(bd),
                    gamma=_p(gamma), invstd=_d(invstd), mean_r=_d(mean_r), dgamma=_p(vec[0]), dbeta=_p(vec[1]),
                    dWd=_p(dWd), dbd=_p(vec[2]), Wcat=_p(Wcat), kvec=_p(kvec), coef=_d(coef), rows=float(rows), C=C, D=D,
                    training=1 if training else 0)
    _call("side backward", lib.sgcn_side_bwd, ctypes.byref(p), _STREAM)
    return dict(dgamma=vec[0], dbeta=vec[1], dWd=dWd, dbd=vec[2], Wcat=Wcat, kvec=kvec)


def bcast_rows(g, rows_per_n, scale, mask_y=None):
    """(n, C) -> (n, rows_per_n, C) with every row = g[n] * scale (gradient of a mean over the rows); mask_y: the pooled
    rows themselves (a ReLU output): the result is additionally multiplied by [mask_y > 0]"""
    lib = _lib.load()
    n, C = g.shape
    out = torch.empty((n, rows_per_n, C), device=g.device, dtype=torch.float32)
    if mask_y is not None and (mask_y.numel() != out.numel() or not mask_y.is_contiguous()):
        raise ValueError("bcast_rows: mask_y must be a contiguous tensor of the output's size")
    if out.numel():
        _launch("bcast_rows", 1, _nbytes(out, mask_y), lib.sgcn_bcast_rows, _p(g, name="g"), _p(out), _p(mask_y), n,
                int(rows_per_n), C, ctypes.c_float(scale), _STREAM)
    return out


def head_fwd(pool_sums, W, b, N, M, count):
    """(pooled [N, C], logits [N, K]) from the pooled sums [N*M, C] f64 (sgcn_head_fwd); count = T*V*M rows per sample"""
    lib = _lib.load()
    K, C = W.shape
    pooled = torch.empty((N, C), device=W.device, dtype=torch.float32)
    logits = torch.empty((N, K), device=W.device, dtype=torch.float32)
    if N:
        _launch("head_fwd", 1, _nbytes(pool_sums, logits), lib.sgcn_head_fwd, _d(pool_sums, "pool_sums"), _p(W, name="fc.weight"),
                _p(b), _p(pooled), _p(logits), N, M, C, K, 1.0 / float(count), _STREAM)
    return pooled, logits


def head_bwd(dlogits, pooled, W, N, M, count, want_bias=True):
    """(dW [K, C], db [K] or None, gpool [N*M, C]) -- sgcn_head_bwd"""
    lib = _lib.load()
    K, C = W.shape
    dW = torch.empty((K, C), device=W.device, dtype=torch.float32)
    db = torch.empty(K, device=W.device, dtype=torch.float32) if want_bias else None
    gpool = torch.empty((N * M, C), device=W.device, dtype=torch.float32)
    _launch("head_bwd", 1, _nbytes(dW, gpool), lib.sgcn_head_bwd, _p(dlogits, name="dlogits"), _p(pooled), _p(W), _p(dW),
            _p(db), _p(gpool), N, M, C, K, ctypes.c_float(1.0 / float(count)), _STREAM)
    return dW, db, gpool


# ------------------------------------------------------------------------------------------------ optimizer step
def sgd_epilogue(param, grad, momentum_buf, weight_decay, ypos_src, hyper, n_param, nesterov):
    """scale -> K5 on the reduced raw position sums -> weight decay -> SGD momentum / Nesterov, one kernel over the flat
    buffers (include/shiftgcn_b200.h:sgcn_sgd_epilogue); hyper = device tensor [lr, momentum, gradient scale]"""
    lib = _lib.load()
    params_changed()                                               # parameters are written through raw pointers
    _launch("sgd_epilogue", 1, 0, lib.sgcn_sgd_epilogue, _p(param, name="flat_param"), _p(grad, name="flat_grad"),
            _p(momentum_buf, name="momentum_buf"), _p(weight_decay, name="weight_decay"),
            _p(ypos_src, torch.int32, "ypos_src"), _p(hyper, name="hyper"), int(n_param), 1 if nesterov else 0, _STREAM)
