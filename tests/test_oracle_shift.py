"""CPU tests of the temporal-shift oracle: C restatement == torch restatement == independent autograd formulation,
adjoint identities, K5 cases, committed vectors (tests/golden/shift_op.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import shift_c, shift_torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _case(seed, n=2, c=6, h=9, w=5):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, c, h, w))
    xpos = rng.uniform(-1e-8, 1e-8, c)
    ypos = np.array([0.3, -1.7, 2.0, -3.0, 5.5, 0.0])[:c]
    return x, xpos, ypos


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_oracle_matches_torch_oracle(stride, dtype):
    x, xpos, ypos = _case(3)
    ypos = ypos + (0.5 if stride != 1 else 0.0)
    x, xpos, ypos = x.astype(dtype), xpos.astype(dtype), ypos.astype(dtype)
    go = np.random.default_rng(4).standard_normal((2, 6, 9 // stride, 5)).astype(dtype)
    t = lambda a: torch.from_numpy(a)
    tol = 1e-5 if dtype == np.float32 else 1e-13
    out_c = shift_c.shift_forward(x, xpos, ypos, stride)
    out_t = shift_torch.shift_forward(t(x), t(xpos), t(ypos), stride).numpy()
    assert np.abs(out_c - out_t).max() <= tol
    gin_c, gx_c, gy_c = shift_c.shift_backward(go, x, xpos, ypos, stride)
    gin_t = shift_torch.shift_backward_input(t(go), t(xpos), t(ypos), 9, stride).numpy()
    assert np.abs(gin_c - gin_t).max() <= tol
    rx_c, ry_c = shift_c.shift_backward_pos_raw(x, go, xpos, ypos, stride)
    rx_t, ry_t = shift_torch.shift_backward_pos_raw(t(x), t(go), t(xpos), t(ypos), stride)
    assert np.abs(ry_c - ry_t.numpy()).max() <= tol * 50
    _, gy_t = shift_torch.shift_constraint(rx_t, ry_t)
    assert np.array_equal(np.sign(gy_c), np.sign(gy_t.numpy()))
    assert np.all(gx_c == 0)


def _independent_shift(x, ypos, stride):
    """differentiable re-derivation (no floor tables shared with the oracle): dense interpolation matrix per channel"""
    n, c, h, w = x.shape
    ho = h // stride
    rows = torch.arange(ho, dtype=x.dtype)[:, None] * stride
    src = torch.arange(h, dtype=x.dtype)[None, :]
    outs = []
    for ch in range(c):
        pos = rows + ypos[ch]                                    # fractional source row per output row
        wgt = torch.clamp(1 - (src - pos).abs(), min=0)          # hat function == linear interpolation, zero padded
        outs.append(torch.einsum("os,nsw->now", wgt, x[:, ch]))
    return torch.stack(outs, 1)


@pytest.mark.parametrize("stride", [1, 2])
def test_oracle_matches_independent_autograd(stride):
    torch.manual_seed(5)
    x = torch.randn(2, 5, 10, 4, dtype=torch.float64, requires_grad=True)
    ypos = torch.tensor([0.25, -1.6, 2.4, -0.4, 3.75], dtype=torch.float64, requires_grad=True)
    eff = ypos + (0.5 if stride != 1 else 0.0)
    go = torch.randn(2, 5, 10 // stride, 4, dtype=torch.float64)
    out_i = _independent_shift(x, eff, stride)
    out_i.backward(go)
    xpos = torch.zeros(5, dtype=torch.float64)
    out_o = shift_torch.shift_forward(x.detach(), xpos, eff.detach(), stride)
    assert (out_o - out_i.detach()).abs().max() < 1e-13
    gin_o = shift_torch.shift_backward_input(go, xpos, eff.detach(), 10, stride)
    assert (gin_o - x.grad).abs().max() < 1e-13                 # K2/K3 are the exact adjoint of K1
    _, raw_y = shift_torch.shift_backward_pos_raw(x.detach(), go, xpos, eff.detach(), stride)
    assert (raw_y * x.shape[0] - ypos.grad).abs().max() < 1e-12  # raw sum == autograd dL/dypos / batch (mean over N)


@pytest.mark.parametrize("stride", [1, 2, 3])
@pytest.mark.parametrize("h", [8, 13])
def test_adjoint_identity(stride, h):
    """<Shift(x), g> == <x, Shift^T(g)> for fractional, integer and out-of-range positions, even and odd H"""
    rng = np.random.default_rng(h * 10 + stride)
    n, c, w = 2, 7, 3
    x = rng.standard_normal((n, c, h, w))
    g = rng.standard_normal((n, c, h // stride, w))
    xpos = np.zeros(c)
    ypos = np.array([0.3, -1.7, 2.0, -3.0, 20.5, 0.0, -0.999])
    lhs = (shift_c.shift_forward(x, xpos, ypos, stride) * g).sum()
    rhs = (x * shift_c.shift_backward_input(g, xpos, ypos, x.shape, stride)).sum()
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))


def test_constraint_cases():
    gx = np.array([3.0, -2.0, 0.5, 7.0])
    gy = np.array([4.0, -1e-30, 0.0, -2.5])
    ox, oy = shift_c.shift_constraint(gx, gy)
    assert np.all(ox == 0)
    assert np.allclose(oy, [0.01, -0.01, 0.0001, -0.01], rtol=0, atol=1e-15)
    oxf, oyf = shift_c.shift_constraint(gx.astype(np.float32), gy.astype(np.float32))
    # in fp32 (-1e-30)^2 underflows to 0 -> the 1e-4 branch, exactly like sqrt(dy*dy) in the reference kernel
    assert np.allclose(oyf, np.array([0.01, 0.0001, 0.0001, -0.01], dtype=np.float32))


def test_half_frame_rule():
    """stride != 1 samples at ypos + 0.5: ypos = 0 averages rows 2h and 2h+1 (cuda/shift.py:14-19)"""
    x = torch.arange(2 * 1 * 8 * 3, dtype=torch.float64).reshape(2, 1, 8, 3)
    out = shift_torch.OracleShiftFunction.apply(x, torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 2)
    want = 0.5 * (x[:, :, 0::2] + x[:, :, 1::2])
    assert torch.equal(out, want)


@pytest.mark.parametrize("stride", [1, 2])
def test_committed_vectors(stride):
    z = np.load(os.path.join(GOLD, "shift_op.npz"))
    g = lambda k: z[f"s{stride}/{k}"]
    t = torch.from_numpy
    out = shift_torch.shift_forward(t(g("x")), t(g("xpos")), t(g("ypos")), stride)
    assert np.abs(out.numpy() - g("out")).max() < 1e-13
    gin = shift_torch.shift_backward_input(t(g("go")), t(g("xpos")), t(g("ypos")), g("x").shape[2], stride)
    assert np.abs(gin.numpy() - g("gin")).max() < 1e-13
    _, ry = shift_torch.shift_backward_pos_raw(t(g("x")), t(g("go")), t(g("xpos")), t(g("ypos")), stride)
    assert np.abs(ry.numpy() - g("raw_y")).max() < 1e-12
