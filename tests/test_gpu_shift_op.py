"""cuda.shift.Shift / ShiftFunction (stand-alone NCHW op) against the oracle restatement of K1-K5."""
import pytest
import torch

from oracle import shift_torch
from util import rel_err

pytestmark = pytest.mark.gpu

CASES = [  # n, c, h, w, stride
    (2, 5, 8, 4, 2), (2, 5, 8, 4, 1), (3, 7, 13, 25, 1), (3, 7, 13, 25, 2), (1, 64, 300, 25, 1), (2, 16, 75, 33, 2),
]


def _ypos(c, gen):
    y = (torch.rand(c, generator=gen, dtype=torch.float64) * 6 - 3)
    if c >= 5:
        y[0], y[1], y[2] = 1.0, -2.0, 0.0          # exact integers: floor() edge
        y[3], y[4] = 9.25, -11.5                    # beyond the tensor for small H
    return y


@pytest.mark.parametrize("n,c,h,w,stride", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_shift_forward_backward(cuda_device, n, c, h, w, stride, dtype):
    from shiftgcn_b200.shift import ShiftFunction
    gen = torch.Generator().manual_seed(1 + n + 10 * c + 100 * h + stride)
    x = torch.randn(n, c, h, w, generator=gen, dtype=torch.float64)
    xpos = (torch.rand(c, generator=gen, dtype=torch.float64) * 2 - 1) * 1e-8
    ypos = _ypos(c, gen)
    go = torch.randn(n, c, h // stride, w, generator=gen, dtype=torch.float64)

    xr, ypr, xpr = x.clone().requires_grad_(True), ypos.clone().requires_grad_(True), xpos.clone().requires_grad_(True)
    shift_torch.RAW_POS_LOG = {}
    out_ref = shift_torch.OracleShiftFunction.apply(xr, xpr, ypr, stride)
    out_ref.backward(go)
    raw_y = shift_torch.RAW_POS_LOG[id(xpr)][1]
    shift_torch.RAW_POS_LOG = None

    xc = x.to(cuda_device, dtype).requires_grad_(True)
    ypc = ypos.to(cuda_device, dtype).requires_grad_(True)
    xpc = xpos.to(cuda_device, dtype).requires_grad_(True)
    out = ShiftFunction.apply(xc, xpc, ypc, stride)
    out.backward(go.to(cuda_device, dtype))
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert out.shape == out_ref.shape
    assert rel_err(out, out_ref) < tol
    assert rel_err(xc.grad, xr.grad) < tol
    assert torch.count_nonzero(xpc.grad).item() == 0            # grad_xpos is an explicit zero tensor
    sure = raw_y.abs() > 1e-4 * raw_y.abs().max()
    assert torch.equal(ypc.grad.double().cpu()[sure].sign(), ypr.grad[sure].sign())
    assert rel_err(ypc.grad.double().cpu()[sure], ypr.grad[sure]) < 1e-6


def test_shift_raw_position_sums(cuda_device):
    """raw (pre-K5) sums agree numerically, not only in sign"""
    from shiftgcn_b200 import ops
    gen = torch.Generator().manual_seed(7)
    n, c, h, w, stride = 4, 32, 50, 25, 2
    x = torch.randn(n, c, h, w, generator=gen, dtype=torch.float64)
    ypos = torch.rand(c, generator=gen, dtype=torch.float64) * 4 - 2 + 0.5
    xpos = torch.zeros(c, dtype=torch.float64)
    go = torch.randn(n, c, h // stride, w, generator=gen, dtype=torch.float64)
    gx_ref, gy_ref = shift_torch.shift_backward_pos_raw(x, go, xpos, ypos, stride)
    out = ops.shift_forward(x.to(cuda_device, torch.float32), xpos.to(cuda_device), ypos.to(cuda_device), stride)
    (_, _, _), raw = ops.shift_backward(go.to(cuda_device, torch.float32).contiguous(), x.to(cuda_device, torch.float32),
                                        out, xpos.to(cuda_device), ypos.to(cuda_device), stride, return_raw=True)
    assert rel_err(raw[1], gy_ref) < 1e-5


def test_reference_error_behaviour(cuda_device):
    """non-CUDA / non-contiguous inputs raise RuntimeError like AT_ASSERTM in shift_cuda.cpp:15-17"""
    from shiftgcn_b200.shift import shift_cuda
    x = torch.randn(1, 4, 6, 5)
    p = torch.zeros(4)
    with pytest.raises(RuntimeError):
        shift_cuda.forward(x, p, p, 1)
    xc = torch.randn(1, 4, 6, 10, device=cuda_device)[..., ::2]
    with pytest.raises(RuntimeError):
        shift_cuda.forward(xc, p.to(cuda_device), p.to(cuda_device), 1)


def test_demo_case(cuda_device):
    """model/Temporal_shift/demo.py:13-29: Shift(5, stride 2) on ones(1,5,8,4); interior outputs are 1, ypos.grad in {+-0.01, 1e-4}"""
    from shiftgcn_b200.shift import Shift
    torch.manual_seed(0)
    layer = Shift(channel=5, stride=2)
    x = torch.ones(1, 5, 8, 4, device=cuda_device, requires_grad=True)
    out = layer(x)
    out.sum().backward()
    assert out.shape == (1, 5, 4, 4)
    assert layer.ypos.grad is not None and layer.xpos.grad is not None
    vals = set(round(v, 6) for v in layer.ypos.grad.abs().cpu().tolist())
    assert vals <= {0.01, 0.0001}
    assert x.grad.shape == x.shape
