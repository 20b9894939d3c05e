"""The fp32-accurate mode of the fused path (SGCN_PREC_FP32: 3xTF32 operand split on the tensor cores) against the EXACT
fp64 oracle.  The reference's contraction is a true fp32 einsum (model/shift_gcn.py:131) and BASELINE.json's north star
asks for fp32 results within 1e-5 relative error; the default TF32 mode is held to 1e-2 (test_gpu_units.py).  Here every
unit type runs forward + backward under ``ops.precision("fp32")``:
  outputs, BatchNorm buffers      <= 1e-5 of the largest reference entry
  input / parameter gradients     <= 5e-5 (a backward pass chains 3-5 contractions and fp32 reductions over n*T*V rows)

The gradient bounds hold only while every ReLU takes the same branch in fp32 and in the fp64 oracle.  A pre-activation
that lies within fp32 rounding (~1e-6 of the tensor maximum) of zero flips its mask and changes a few thousand gradient
entries by O(10 %) -- about one such element per 1-2 million ReLU inputs.  The seeds below are draws without such an
element (tools/diag_unit_bwd.py locates the offending entry when a new draw has one: with seed 54 the 256-channel
multi-tile case differs from the oracle in exactly ONE element of the conv-output gradient, (n,t,v,c) = (0,46,12,56)).
"""
import copy

import pytest
import torch

from oracle import model_ref
from util import check_ypos_grad, fill_pair, raw_pos_log, rel_err

pytestmark = pytest.mark.gpu

TOL_OUT, TOL_GRAD = 1e-5, 5e-5


def _oracle(ref, x, go, train):
    ref = copy.deepcopy(ref).double()
    ref.train(train)
    xr = x.double().clone().requires_grad_(True)
    with raw_pos_log(ref) as log:
        out = ref(xr)
        out.backward(go.double())
    grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    bufs = {k: b.clone() for k, b in ref.named_buffers()}
    return out.detach(), xr.grad, grads, bufs, log.raw


def _check_fp32(mod, ref, x, go, train, device):
    from shiftgcn_b200 import ops
    mod = mod.to(device)
    mod.train(train)
    xc = x.to(device).requires_grad_(True)
    with ops.precision("fp32"):
        out = mod(xc)
        out.backward(go.to(device))
    torch.cuda.synchronize()
    assert ops.get_precision() == "tf32"
    out_r, gx_r, grads_r, bufs_r, raw = _oracle(ref, x, go, train)
    assert rel_err(out, out_r) < TOL_OUT, f"output: {rel_err(out, out_r):.2e}"
    assert rel_err(xc.grad, gx_r) < TOL_GRAD, f"input grad: {rel_err(xc.grad, gx_r):.2e}"
    floor = 1e-7 * max(go.abs().sum().item(), 1.0)       # analytically-zero gradients (a bias in front of a train-mode BN)
    for name, p in mod.named_parameters():
        if not p.requires_grad:
            continue
        want = grads_r[name]
        if name.endswith("ypos"):
            check_ypos_grad(name, p.grad, want, raw.get(name), decisive=1e-4)
            continue
        if name.endswith("xpos"):
            assert torch.count_nonzero(p.grad).item() == 0
            continue
        err = (p.grad.double().cpu() - want).abs().max().item()
        assert err < TOL_GRAD * want.abs().max().item() + floor, f"grad {name}: {err:.3e} of {want.abs().max().item():.3e}"
    if train:
        for name, b in mod.named_buffers():
            if b.dtype.is_floating_point:
                assert rel_err(b, bufs_r[name]) < TOL_OUT, f"buffer {name}"


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,D,V,n,T", [(64, 64, 25, 2, 23), (64, 128, 25, 2, 17), (128, 128, 33, 1, 19), (128, 256, 25, 1, 16),
                                       (256, 256, 25, 1, 17), (256, 256, 33, 1, 11), (3, 64, 25, 2, 15)])
def test_shift_gcn_fp32(cuda_device, C, D, V, n, T, train):
    from shiftgcn_b200.modules import Shift_gcn
    torch.manual_seed(1)
    mod = Shift_gcn(C, D, None, num_point=V)
    ref = model_ref.RefShiftGcn(C, D, None, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(51)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T, V, generator=g)
    _check_fp32(mod, ref, x, go, train, cuda_device)


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,V,n,T,stride", [(64, 25, 2, 33, 1), (128, 25, 2, 26, 2), (256, 25, 1, 31, 1), (128, 33, 1, 27, 1),
                                            (256, 33, 1, 24, 2)])
def test_shift_tcn_fp32(cuda_device, C, V, n, T, stride, train):
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(1)
    mod = Shift_tcn(C, C, stride=stride)
    ref = model_ref.RefShiftTcn(C, C, stride=stride)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(52)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, C, T // stride, V, generator=g)
    _check_fp32(mod, ref, x, go, train, cuda_device)


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,D,V,n,T,stride,residual", [
    (64, 64, 25, 2, 27, 1, True), (128, 128, 33, 1, 22, 1, True), (256, 256, 25, 1, 21, 1, True),
    (64, 128, 25, 2, 24, 2, True), (128, 256, 25, 1, 26, 2, True), (3, 64, 25, 2, 20, 1, False)])
def test_tcn_gcn_unit_fp32(cuda_device, C, D, V, n, T, stride, residual, train):
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(C, D, None, stride=stride, residual=residual, num_point=V)
    ref = model_ref.RefUnit(C, D, None, stride=stride, residual=residual, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(53)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T // stride, V, generator=g)
    _check_fp32(mod, ref, x, go, train, cuda_device)


@pytest.mark.parametrize("C,D,V,n,T,stride", [(64, 64, 25, 2, 43, 1), (256, 256, 25, 1, 47, 1), (64, 128, 25, 2, 38, 2),
                                              (128, 256, 25, 1, 42, 2), (128, 128, 33, 1, 37, 1)])
def test_tcn_gcn_unit_fp32_many_tiles_per_cta(cuda_device, C, D, V, n, T, stride):
    """the multi-tile steady state of the 3xTF32 pipelines (head / tail operand halves, four streamed weight blocks per
    chunk, one-deep weight ring at 256 channels): 3 persistent CTAs, partial last tile, training mode"""
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(C, D, None, stride=stride, residual=True, num_point=V)
    ref = model_ref.RefUnit(C, D, None, stride=stride, residual=True, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(53)             # see the module docstring: seed 54 draws a ReLU input at the kink
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T // stride, V, generator=g)
    prev = ops.set_max_ctas(3)
    try:
        _check_fp32(mod, ref, x, go, True, cuda_device)
    finally:
        ops.set_max_ctas(prev)


@pytest.mark.parametrize("num_class,V,M", [(60, 25, 2), (2, 33, 1)])
def test_model_eval_fp32(cuda_device, num_class, V, M):
    """whole model, inference: ten chained units -> 1e-4; separates TF32 rounding (1.6e-2 on the 2-class MediaPipe head in
    the default mode, test_gpu_model.py) from defects of the kernels"""
    from shiftgcn_b200 import ops
    from test_gpu_model import _build, _calibrate
    mod, ref = _build(num_class, V, M, cuda_device)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 32, V, M, generator=g)
    _calibrate(ref, mod, x)
    mod.eval(), ref.eval()
    with torch.no_grad(), ops.precision("fp32"):
        out = mod(x.to(cuda_device))
    with torch.no_grad():
        want = ref(x.double())
    assert rel_err(out, want) < 1e-4, f"{rel_err(out, want):.2e}"
    assert torch.equal(out.argmax(1).cpu(), want.argmax(1))


def test_precision_mode_guards(cuda_device):
    from shiftgcn_b200 import ops
    with pytest.raises(ValueError):
        ops.set_precision("bf16")
    W = torch.randn(64, 64, device=cuda_device)
    img = ops.weight_image(W, 1, 64, 64, 64)                     # TF32 image ...
    x = torch.randn(1, 5, 25, 64, device=cuda_device)
    with ops.precision("fp32"), pytest.raises(RuntimeError):     # ... is refused by an fp32-mode launch
        ops.rowgemm(ops.PRO_PLAIN, ops.EPI_LINEAR, in0=x, out=torch.empty_like(x), wimg=img, groups=5, V=25, K=64, N=64)
