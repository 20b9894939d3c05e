"""ctypes binding of libshiftgcn_b200.so (the C ABI declared in include/shiftgcn_b200.h).

There is deliberately NO fallback: if the library is missing or the device is not sm_100, every op raises.
"""
import ctypes
import os

from . import build as _build

c_float_p = ctypes.c_void_p   # device pointers travel as integers
c_double_p = ctypes.c_void_p
_i, _ll, _d, _vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p


class SgcnRowGemm(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("in0", "in1", "out", "wimg", "pro_a", "pro_b", "pro_c", "bias", "epi_a", "epi_b",
                                   "res", "res2", "res2m", "xin", "stats", "red0")] + \
               [("groups", _ll), ("V", _i), ("G", _i), ("T", _i), ("K", _i), ("N", _i), ("relu", _i), ("k0", _i),
                ("in0_gs", _i), ("in1_gs", _i), ("out_gs", _i), ("accum", _i), ("prec", _i)]


class SgcnWgrad(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("a_src", "a_tab0", "b_src", "b_src2", "b_tab0", "b_tab1", "b_tab2", "dw")] + \
               [("groups", _ll), ("V", _i), ("G", _i), ("T", _i), ("CA", _i), ("CB", _i), ("a_gs", _i), ("b_gs", _i),
                ("prec", _i)]


class SgcnTShift(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("q", "res", "out", "ypos_eff", "scale", "shift", "stats")] + \
               [("n_samples", _ll), ("T_in", _i), ("T_out", _i), ("V", _i), ("C", _i), ("stride", _i), ("relu", _i)]


class SgcnTShiftBwd(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("q", "gy", "y", "ypos_eff", "mean", "invstd", "k1", "m1", "m2", "sums", "dpre",
                                   "dbias")] + \
               [("n_samples", _ll), ("T_in", _i), ("T_out", _i), ("V", _i), ("C", _i), ("stride", _i), ("relu", _i)]


class SgcnTShiftInBwd(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("dp", "h", "z", "ypos_eff", "mean", "invstd", "scale", "shift", "k1", "m1", "m2",
                                   "zmean", "zinvstd", "sums", "vd_sums", "gh", "pos_sums", "gate")] + \
               [("n_samples", _ll), ("T", _i), ("V", _i), ("C", _i), ("relu_h", _i)]


class SgcnTShiftInSums(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("dp", "ypos_eff", "Wt", "dWt", "dbt", "mean", "invstd", "scale", "shift", "sums",
                                   "gate")] + \
               [("n_samples", _ll), ("T", _i), ("V", _i), ("C", _i)]


class SgcnStem(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("x", "maskmul", "W", "bias", "Wd", "bd", "stats_vd", "stats_r", "sc1", "sh1", "sc2",
                                   "sh2", "h", "stats_h", "g", "mean1", "invstd1", "mean2", "invstd2", "vd_sums",
                                   "r_sums", "al", "be", "ga", "a2", "b2", "c2", "dw_raw", "dmask_raw", "dx")] + \
               [("groups", _ll), ("V", _i), ("D", _i)]


class SgcnSideFold(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("sx_sums", "XX", "Wd", "bd", "gamma", "beta", "running_mean", "running_var",
                                   "num_batches_tracked", "Wf", "bf", "mean_r", "invstd", "sx", "counter")] + \
               [("rows", _d), ("eps", _d), ("momentum", _d), ("C", _i), ("D", _i), ("training", _i)]


class SgcnSideBwd(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("P", "sg", "XX", "sx", "Wd", "bd", "gamma", "invstd", "mean_r", "dgamma", "dbeta",
                                   "dWd", "dbd", "Wcat", "kvec", "coef")] + \
               [("rows", _d), ("C", _i), ("D", _i), ("training", _i)]


# name -> argtypes (restype is always int unless noted); must list every symbol of include/shiftgcn_b200.h
SIGNATURES = {
    "sgcn_abi_version": [],
    "sgcn_device_check": [],
    "sgcn_set_traversal": [_i],
    "sgcn_set_max_ctas": [_i],
    "sgcn_selftest_umma": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "sgcn_selftest_probe": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "sgcn_shift_fwd_nchw_f32": [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "sgcn_shift_fwd_nchw_f64": [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "sgcn_shift_bwd_nchw_f32": [_vp] * 9 + [_ll, _i, _i, _i, _i, _vp],
    "sgcn_shift_bwd_nchw_f64": [_vp] * 9 + [_ll, _i, _i, _i, _i, _vp],
    "sgcn_input_stream": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp],
    "sgcn_window_stream": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "sgcn_window_scores": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "sgcn_head_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _vp],
    "sgcn_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp],
    "sgcn_data_bn_stats": [_vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "sgcn_data_bn_bwd": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "sgcn_random_move": [_vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp],
    "sgcn_side_fold": [ctypes.POINTER(SgcnSideFold), _vp],
    "sgcn_side_bwd": [ctypes.POINTER(SgcnSideBwd), _vp],
    "sgcn_stem_fwd": [ctypes.POINTER(SgcnStem), _i, _vp],
    "sgcn_stem_bwd": [ctypes.POINTER(SgcnStem), _i, _vp],
    "sgcn_rowgemm": [ctypes.POINTER(SgcnRowGemm), _i, _i, _vp],
    "sgcn_wgrad": [ctypes.POINTER(SgcnWgrad), _i, _vp],
    "sgcn_bn_res_relu_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp],
    "sgcn_tshift_fwd": [ctypes.POINTER(SgcnTShift), _i, _vp],
    "sgcn_tshift_bwd": [ctypes.POINTER(SgcnTShiftBwd), _i, _vp],
    "sgcn_tshift_in_bwd": [ctypes.POINTER(SgcnTShiftInBwd), _i, _vp],
    "sgcn_tshift_in_bwd_sums": [ctypes.POINTER(SgcnTShiftInSums), _vp],
    "sgcn_shift_pos_finalize": [_vp, _vp, _vp, _vp, _i, _d, _vp],
    "sgcn_relu_bn1d_bwd_stats": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp],
    "sgcn_bcast_rows": [_vp, _vp, _vp, _ll, _ll, _i, ctypes.c_float, _vp],
    "sgcn_channel_stats": [_vp, _vp, _ll, _i, _vp],
    "sgcn_channel_stats_groups": [_vp, _vp, _ll, _i, _i, _i, _vp],
    "sgcn_relu_mask_grad": [_vp, _vp, _vp, _ll, _vp],
    "sgcn_bn_fwd_finalize": [_vp] * 10 + [_i, _d, _d, _d, _i, _vp],
    "sgcn_tshift_bwd_finalize": [_vp] * 11 + [_i, _d, _d, _i, _vp],
    "sgcn_tshift_in_bwd_finalize": [_vp] * 11 + [_i, _d, _d, _i, _vp],
    "sgcn_bn1d_bwd_finalize": [_vp] * 10 + [_i, _i, _d, _i, _vp],
    "sgcn_mask_prepare": [_vp, _vp, _i, _vp],
    "sgcn_mask_prepare_rot": [_vp, _vp, _vp, _i, _i, _vp],
    "sgcn_mask_grad_finalize": [_vp, _vp, _vp, _i, _vp],
    "sgcn_prep_weight_image": [_vp, _ll, _ll, _i, _i, _vp, _vp],
    "sgcn_prep_weight_image_split": [_vp, _ll, _ll, _i, _i, _vp, _vp],
    "sgcn_reduce_export": [_vp, _vp, _i, _d, _vp],
    "sgcn_sgd_epilogue": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp],
}

_lib = None


def traversal_mode():
    """SGCN_SNAKE: 0 = always ascending, 1 = snake (first kernel of a step descends, default), 2 = snake, first ascends"""
    v = os.environ.get("SGCN_SNAKE", "1")
    return int(v) if v in ("0", "1", "2") else 1


def library_path():
    # SGCN_LIB: developer override for A/B runs of two builds of the same ABI
    return os.environ.get("SGCN_LIB") or _build.LIB_PATH


def load():
    """Load the shared library (never builds implicitly; run ``__graft_entry__.build()`` first)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"shiftgcn_b200: {path} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc 12.9). There is no CPU or PyTorch fallback for the Shift-GCN hot path.")
    lib = ctypes.CDLL(path)
    lib.sgcn_last_error.restype = ctypes.c_char_p
    lib.sgcn_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library drift
        fn.restype = ctypes.c_int
        fn.argtypes = argtypes
    # snake traversal of the full-tensor kernels (on unless SGCN_SNAKE=0), see include/shiftgcn_b200.h
    lib.sgcn_set_traversal(traversal_mode())
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sgcn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"shiftgcn_b200 {what} failed (code {rc}): {msg}")
