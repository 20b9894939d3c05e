"""sgcn_side_fold / sgcn_side_bwd (csrc/side.cu) against the same parameter arithmetic in torch fp64: the conv + BatchNorm
side branches (`down`, model/shift_gcn.py:82-86; strided `tcn` residual, :31-45) as autograd sees them."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _branch(x, Wd, bd, gamma, beta, training, rmean, rvar, eps=1e-5):
    """BatchNorm2d(Conv1x1(x)) on rows in fp64 (the arbiter)"""
    r = x @ Wd.t() + bd
    if training:
        mean, var = r.mean(0), r.var(0, unbiased=False)
    else:
        mean, var = rmean, rvar
    return (r - mean) / torch.sqrt(var + eps) * gamma + beta


@pytest.mark.parametrize("C,D,training", [(64, 128, True), (128, 256, True), (64, 128, False)])
def test_side_kernels_match_autograd_of_conv_bn(cuda_device, C, D, training):
    from shiftgcn_b200 import ops
    g = torch.Generator().manual_seed(C + D)
    rows = 600
    x = torch.randn(rows, C, generator=g, dtype=torch.float64)
    Wd = (torch.randn(D, C, generator=g, dtype=torch.float64) / C ** 0.5).requires_grad_(True)
    bd = (0.1 * torch.randn(D, generator=g, dtype=torch.float64)).requires_grad_(True)
    gamma = (0.5 + torch.rand(D, generator=g, dtype=torch.float64)).requires_grad_(True)
    beta = (0.1 * torch.randn(D, generator=g, dtype=torch.float64)).requires_grad_(True)
    rmean, rvar = 0.1 * torch.randn(D, generator=g, dtype=torch.float64), 0.5 + torch.rand(D, generator=g, dtype=torch.float64)
    G = torch.randn(rows, D, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    out = _branch(xr, Wd, bd, gamma, beta, training, rmean, rvar)
    out.backward(G)

    dev = cuda_device
    f32 = lambda t: t.detach().float().contiguous().to(dev)
    XX = f32(x.t() @ x) if training else None
    sums = torch.stack([x.sum(0), (x * x).sum(0)], 1).contiguous().to(dev) if training else None      # [C][2] fp64
    counter = torch.zeros(1, dtype=torch.int32, device=dev) if training else None
    rm32, rv32 = f32(rmean), f32(rvar)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    f = ops.side_fold(f32(Wd), f32(bd), f32(gamma), f32(beta), rm32, rv32, nbt if training else None, rows, 1e-5, 0.1,
                      training, sx_sums=sums, XX=XX, counter=counter)
    folded = x @ f["Wf"].double().cpu().t() + f["bf"].double().cpu()           # the forward GEMM the product runs
    assert rel_err(folded, out) < 2e-5
    if training:
        assert sums.abs().max().item() == 0.0 and counter.item() == 0           # scratch handed back zeroed
        assert nbt.item() == 1
        r = x @ Wd.detach().t() + bd.detach()
        assert rel_err(rm32, 0.9 * rmean + 0.1 * r.mean(0)) < 1e-5
        assert rel_err(rv32, 0.9 * rvar + 0.1 * r.var(0, unbiased=True)) < 1e-4

    P = f32(x.t() @ G)
    b = ops.side_bwd(P, f32(G.sum(0)), f32(Wd), f32(bd), f32(gamma), f["invstd"], f["mean_r"], rows, training,
                     sx=f["sx"], XX=XX)
    assert rel_err(b["dgamma"], gamma.grad) < 1e-4
    assert rel_err(b["dbeta"], beta.grad) < 1e-5
    assert rel_err(b["dWd"], Wd.grad) < 1e-4
    # a conv bias in front of a training-mode BatchNorm has no gradient; the kernel returns rounding noise of the sums
    assert (b["dbd"].double().cpu() - bd.grad).abs().max().item() < 1e-3 * max(1.0, G.abs().sum(0).max().item())
    dx = torch.cat([G, x], 1) @ b["Wcat"].double().cpu() + b["kvec"].double().cpu()    # the [G | x] input-gradient GEMM
    assert rel_err(dx, xr.grad) < 1e-4
