"""The reference's OWN kernels (shift_cuda_kernel.cu compiled into oracle/_ref/ by oracle/build_ref_ext.py) against
the oracle restatement and against the product kernels -- this is what pins the oracle for the temporal shift."""
import pytest
import torch

from oracle import build_ref_ext, shift_torch
from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_ext():
    if build_ref_ext.built_library() is None:
        pytest.skip("oracle/_ref/shift_cuda_ref.so not built (needs /root/reference at build time)")
    return build_ref_ext.load()


@pytest.mark.parametrize("n,c,h,w,stride", [(2, 5, 8, 4, 2), (3, 7, 13, 25, 1), (3, 7, 13, 25, 2), (2, 64, 40, 25, 1),
                                             (2, 64, 40, 33, 2)])
def test_reference_kernels_vs_oracle_and_product(cuda_device, ref_ext, n, c, h, w, stride):
    from shiftgcn_b200 import ops
    g = torch.Generator().manual_seed(n + 10 * c + 100 * h + stride)
    x = torch.randn(n, c, h, w, generator=g)
    xpos = (torch.rand(c, generator=g) * 2 - 1) * 1e-8
    ypos = torch.rand(c, generator=g) * 6 - 3
    if c >= 5:
        ypos[0], ypos[1], ypos[2] = 1.0, -2.0, 0.0
    if stride != 1:
        ypos = ypos + 0.5
    go = torch.randn(n, c, h // stride, w, generator=g)
    xc, xpc, ypc, goc = (t.to(cuda_device).contiguous() for t in (x, xpos, ypos, go))

    out_ref = ref_ext.forward(xc, xpc, ypc, stride)
    gin_ref, gx_ref, gy_ref = ref_ext.backward(goc, xc, out_ref, xpc, ypc, stride)
    torch.cuda.synchronize()

    # (a) oracle restatement (fp64 on CPU) == reference kernels (fp32 on the B200)
    out_o = shift_torch.shift_forward(x.double(), xpos.double(), ypos.double(), stride)
    gin_o = shift_torch.shift_backward_input(go.double(), xpos.double(), ypos.double(), h, stride)
    _, raw_y = shift_torch.shift_backward_pos_raw(x.double(), go.double(), xpos.double(), ypos.double(), stride)
    _, gy_o = shift_torch.shift_constraint(torch.zeros_like(raw_y), raw_y)
    assert rel_err(out_ref, out_o) < 1e-5
    assert rel_err(gin_ref, gin_o) < 1e-5
    sure = raw_y.abs() > 1e-3 * raw_y.abs().max()
    assert rel_err(gy_ref.double().cpu()[sure], gy_o[sure]) < 1e-6
    assert torch.count_nonzero(gx_ref).item() == 0

    # (b) product kernels == reference kernels
    out = ops.shift_forward(xc, xpc, ypc, stride)
    gin, gx, gy = ops.shift_backward(goc, xc, out, xpc, ypc, stride)
    assert rel_err(out, out_ref) < 1e-5
    assert rel_err(gin, gin_ref) < 1e-5
    assert torch.equal(gy.cpu()[sure], gy_ref.cpu()[sure])
    assert torch.count_nonzero(gx).item() == 0
