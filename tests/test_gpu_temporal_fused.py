"""The eval-mode temporal unit as ONE kernel (sgcn_rowgemm LERP x TSHIFT: BN -> shift -> 1x1 conv -> ReLU -> shift -> BN ->
residual -> ReLU, model/shift_gcn.py:65-74, 160-162) against the two-kernel path it replaces (the same arithmetic up to
single TF32 roundings) and against the fp64 oracle.  Covers frame blocks in front of / behind the sample ends, partial last blocks,
input shift positions outside the staged window (global taps), many tiles per CTA, and the fall-back when the output
shift positions do not fit one window."""
import copy

import pytest
import torch

from oracle import model_ref
from util import fill_pair, rel_err

pytestmark = pytest.mark.gpu


def _narrow_out_positions(mod, ref, lo=-1.0, hi=1.95):
    """output shift positions inside one window of three floor values (the input shift keeps its outliers)"""
    with torch.no_grad():
        t = mod.tcn1 if hasattr(mod, "tcn1") else mod
        r = ref.tcn1 if hasattr(ref, "tcn1") else ref
        C = t.shift_out.ypos.numel()
        g = torch.Generator().manual_seed(77)
        y = torch.rand(C, generator=g) * (hi - lo) + lo
        y[0], y[1] = 0.0, 1.0                                          # exact integers: f = 0 taps
        t.shift_out.ypos.copy_(y)
        r.shift_out.ypos.copy_(y.to(r.shift_out.ypos.dtype))
        t.reset_xpos_check()


def _spy_epilogues():
    from shiftgcn_b200 import ops
    seen = []
    orig = ops.rowgemm

    def spy(pro, epi, **k):
        seen.append((pro, epi))
        return orig(pro, epi, **k)
    return seen, orig, spy


@pytest.mark.parametrize("C,n,T,cap", [(64, 2, 47, 0), (64, 1, 13, 0), (64, 3, 22, 0), (128, 2, 45, 0), (256, 1, 67, 0),
                                       (64, 2, 300, 3), (256, 2, 75, 5)])
def test_fused_eval_temporal_unit(cuda_device, C, n, T, cap):
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(C, C, None, stride=1, residual=True, num_point=25)
    ref = model_ref.RefUnit(C, C, None, stride=1, residual=True, num_point=25)
    fill_pair(mod, ref)
    _narrow_out_positions(mod, ref)
    mod = mod.to(cuda_device).eval()
    ref = copy.deepcopy(ref).double().eval()
    x = torch.randn(n, C, T, 25, generator=torch.Generator().manual_seed(5))
    seen, orig, spy = _spy_epilogues()
    prev = ops.set_max_ctas(cap)
    ops.rowgemm = spy
    try:
        with torch.no_grad():
            got = mod(x.to(cuda_device))
    finally:
        ops.rowgemm = orig
        ops.set_max_ctas(prev)
    assert (ops.PRO_LERP, ops.EPI_TSHIFT) in seen and (ops.PRO_LERP, ops.EPI_LINEAR) not in seen
    # the two-kernel path on the same module
    mod.tcn1.out_window_ok = lambda: False
    try:
        with torch.no_grad():
            two = mod(x.to(cuda_device))
    finally:
        del mod.tcn1.out_window_ok
    # same arithmetic, but the tile geometry decides which rows take the "every tap inside the sample" form of the
    # prologue (one fused multiply-add more or less in front of the TF32 rounding): equal up to single TF32 roundings
    assert rel_err(got, two) < 5e-4, f"max diff {float((got - two).abs().max()):.3e}"
    with torch.no_grad():
        model_ref.TF32_EMULATION = True
        try:
            want = ref(x.double())
        finally:
            model_ref.TF32_EMULATION = False
        exact = ref(x.double())
    assert rel_err(got, want) < 1e-3
    # fp32-accurate mode: the fused kernel against the EXACT oracle
    prev = ops.set_max_ctas(cap)
    try:
        with torch.no_grad(), ops.precision("fp32"):
            got32 = mod(x.to(cuda_device))
    finally:
        ops.set_max_ctas(prev)
    assert rel_err(got32, exact) < 1e-5


@pytest.mark.parametrize("C,T,stride", [(64, 33, 1), (128, 27, 1)])
def test_fused_eval_shift_tcn_module_and_fp32(cuda_device, C, T, stride):
    """Shift_tcn on its own (no residual, no final ReLU), also in the fp32-accurate mode against the exact oracle"""
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(1)
    mod, ref = Shift_tcn(C, C, stride=stride), model_ref.RefShiftTcn(C, C, stride=stride)
    fill_pair(mod, ref)
    _narrow_out_positions(mod, ref)
    mod = mod.to(cuda_device).eval()
    ref = copy.deepcopy(ref).double().eval()
    x = torch.randn(2, C, T, 25, generator=torch.Generator().manual_seed(6))
    seen, orig, spy = _spy_epilogues()
    ops.rowgemm = spy
    try:
        with torch.no_grad(), ops.precision("fp32"):
            got = mod(x.to(cuda_device))
    finally:
        ops.rowgemm = orig
    assert (ops.PRO_LERP, ops.EPI_TSHIFT) in seen
    with torch.no_grad():
        want = ref(x.double())
    assert rel_err(got, want) < 1e-5


def test_wide_output_positions_fall_back(cuda_device):
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(1)
    mod, ref = Shift_tcn(64, 64), model_ref.RefShiftTcn(64, 64)
    fill_pair(mod, ref)                                              # fill values include |ypos| > 8
    mod = mod.to(cuda_device).eval()
    assert not mod.out_window_ok()
    x = torch.randn(1, 64, 20, 25, generator=torch.Generator().manual_seed(7))
    seen, orig, spy = _spy_epilogues()
    ops.rowgemm = spy
    try:
        with torch.no_grad():
            got = mod(x.to(cuda_device))
    finally:
        ops.rowgemm = orig
    assert (ops.PRO_LERP, ops.EPI_LINEAR) in seen and (ops.PRO_LERP, ops.EPI_TSHIFT) not in seen
    with torch.no_grad():
        model_ref.TF32_EMULATION = True
        try:
            want = copy.deepcopy(ref).double().eval()(x.double())
        finally:
            model_ref.TF32_EMULATION = False
    assert rel_err(got, want) < 1e-3
