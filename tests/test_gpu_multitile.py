"""Parity of the STEADY STATE of the persistent tensor-core kernels (fused_gemm.cuh, spatial_bwd.cu, wgrad.cu).

Every GEMM-type kernel launches min(#SM, #tiles) persistent CTAs that walk tile = blockIdx.x + i * gridDim.x.  The
bench runs ~52 tiles per CTA; the small parity cases of test_gpu_units.py give every CTA exactly one tile.  Here every
CTA processes many tiles, so the second TMEM accumulator, the accumulator / operand / raw / weight ring wrap across
tiles, the mbarrier phase parities beyond the first use, wgrad's cross-tile accumulation in TMEM, partial last tiles
and both traversal directions are all compared against the oracle:

  * capped grids (sgcn_set_max_ctas: 3 and 7 CTAs) on small tensors -> 5..60 tiles per CTA, incl. a partial last tile;
  * the benchmark's own sequence lengths (n = 8 samples, T = 300 / 150 / 75 -> up to 2400 groups = 480 tiles of 5
    groups, >= 3 per CTA on 148 SMs) with the full grid;
  * self-consistency: the SAME module and input with 3 CTAs (hundreds of tiles each), 7 CTAs and the full grid.  In
    inference mode nothing depends on how tiles are spread over CTAs, so outputs and input gradients must be
    BIT-IDENTICAL and the weight gradients (summed per CTA in TMEM, then atomically) equal to accumulation-order noise.

Against the oracle the gradients are compared in the robust form of test_gpu_units._compare (a few ReLU flips among
millions of entries are expected, see there).
"""
import pytest
import torch

from oracle import model_ref
from test_gpu_units import _compare
from util import fill_pair

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture
def capped(request, cuda_device):
    """run the test body with the persistent grid capped to request.param CTAs and a chosen first traversal direction"""
    from shiftgcn_b200 import _lib, ops
    ctas, first = request.param
    lib = _lib.load()
    prev = ops.set_max_ctas(ctas)
    prev_snake = lib.sgcn_set_traversal(first)            # 1: first kernel descends, 2: first kernel ascends
    try:
        yield ctas
    finally:
        ops.set_max_ctas(prev)
        lib.sgcn_set_traversal(_lib.traversal_mode() if prev_snake else 0)


CAPS = [(3, 1), (7, 2)]


@pytest.mark.parametrize("capped", CAPS, indirect=True)
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,D,V,n,T", [(64, 64, 25, 2, 41), (64, 128, 25, 2, 33), (128, 128, 25, 2, 27), (128, 256, 25, 1, 43),
                                       (256, 256, 25, 1, 41), (64, 64, 33, 2, 25), (128, 128, 33, 1, 31),
                                       (256, 256, 33, 1, 28)])
def test_shift_gcn_many_tiles_per_cta(cuda_device, capped, C, D, V, n, T, train):
    """n*T is never a multiple of the tile's group count: the last tile is partial"""
    from shiftgcn_b200.modules import Shift_gcn
    torch.manual_seed(1)
    mod = Shift_gcn(C, D, None, num_point=V)
    ref = model_ref.RefShiftGcn(C, D, None, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T, V, generator=g)
    _compare(mod, ref, x, go, train, cuda_device, robust=True)


@pytest.mark.parametrize("capped", CAPS, indirect=True)
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,V,n,T,stride", [(64, 25, 2, 61, 1), (128, 25, 2, 44, 2), (256, 25, 1, 53, 1), (64, 33, 2, 35, 2),
                                            (256, 33, 1, 47, 1)])
def test_shift_tcn_many_tiles_per_cta(cuda_device, capped, C, V, n, T, stride, train):
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(1)
    mod = Shift_tcn(C, C, stride=stride)
    ref = model_ref.RefShiftTcn(C, C, stride=stride)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(22)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, C, T // stride, V, generator=g)
    _compare(mod, ref, x, go, train, cuda_device, robust=True)


@pytest.mark.parametrize("capped", CAPS, indirect=True)
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,D,V,n,T,stride,residual", [
    (64, 64, 25, 2, 43, 1, True),        # identity unit -> UnitFn
    (128, 128, 25, 1, 57, 1, True),
    (256, 256, 25, 1, 38, 1, True),
    (64, 128, 25, 2, 34, 2, True),       # strided units -> ConvUnitFn (conv side branches, [g | x] input-gradient GEMM)
    (128, 256, 25, 1, 46, 2, True),
    (128, 128, 33, 1, 40, 1, True),
    (64, 128, 33, 1, 38, 2, True),
    (3, 64, 25, 2, 30, 1, False),        # first layer (stem kernels + temporal unit)
])
def test_tcn_gcn_unit_many_tiles_per_cta(cuda_device, capped, C, D, V, n, T, stride, residual, train):
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(C, D, None, stride=stride, residual=residual, num_point=V)
    ref = model_ref.RefUnit(C, D, None, stride=stride, residual=residual, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(23)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T // stride, V, generator=g)
    _compare(mod, ref, x, go, train, cuda_device, robust=True)


# ---------------------------------------------------------------------------------------------- benchmark-length sequences
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("C,D,V,n,T", [(64, 64, 25, 8, 300), (128, 128, 25, 8, 150), (256, 256, 25, 16, 75),
                                       (64, 64, 33, 6, 300)])
def test_shift_gcn_full_length(cuda_device, C, D, V, n, T, train):
    """SURVEY section 8d C1 shape family: >= 3 tiles per persistent CTA with the full 148-CTA grid"""
    from shiftgcn_b200.modules import Shift_gcn
    torch.manual_seed(1)
    mod = Shift_gcn(C, D, None, num_point=V)
    ref = model_ref.RefShiftGcn(C, D, None, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(31)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T, V, generator=g)
    _compare(mod, ref, x, go, train, cuda_device, robust=True)


@pytest.mark.parametrize("C,D,V,n,T,stride", [(64, 64, 25, 8, 300, 1), (64, 128, 25, 8, 300, 2), (128, 128, 25, 8, 150, 1),
                                              (256, 256, 25, 16, 75, 1), (128, 128, 33, 6, 150, 1)])
def test_tcn_gcn_unit_full_length_train(cuda_device, C, D, V, n, T, stride):
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(C, D, None, stride=stride, residual=True, num_point=V)
    ref = model_ref.RefUnit(C, D, None, stride=stride, residual=True, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T // stride, V, generator=g)
    _compare(mod, ref, x, go, True, cuda_device, robust=True)


@pytest.mark.parametrize("V,n,T", [(25, 16, 300), (33, 6, 300), (25, 3, 41)])
def test_first_unit_full_length_train(cuda_device, V, n, T):
    """l1 (3 -> 64 channels, residual=False): the stem backward streams g and h through a two-stage bulk-copy ring;
    4800 groups = 6 chunks of 8 groups per block on the full grid (V=33: chunks of 6 with a partial last one)"""
    from shiftgcn_b200.modules import TCN_GCN_unit
    torch.manual_seed(1)
    mod = TCN_GCN_unit(3, 64, None, residual=False, num_point=V)
    ref = model_ref.RefUnit(3, 64, None, residual=False, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(34)
    x = torch.randn(n, 3, T, V, generator=g)
    go = torch.randn(n, 64, T, V, generator=g)
    _compare(mod, ref, x, go, True, cuda_device, robust=True)


@pytest.mark.parametrize("C,V,n,T,stride", [(64, 25, 8, 300, 1), (128, 25, 8, 300, 2), (256, 25, 16, 75, 1)])
def test_shift_tcn_full_length_eval(cuda_device, C, V, n, T, stride):
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(1)
    mod = Shift_tcn(C, C, stride=stride)
    ref = model_ref.RefShiftTcn(C, C, stride=stride)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(33)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, C, T // stride, V, generator=g)
    _compare(mod, ref, x, go, False, cuda_device, robust=True)


# ---------------------------------------------------------------------------------------------- self-consistency
def _run(mod, x, go, cap, train, first):
    from shiftgcn_b200 import _lib, ops
    lib = _lib.load()
    prev = ops.set_max_ctas(cap)
    lib.sgcn_set_traversal(first)
    try:
        mod.train(train)
        mod.zero_grad(set_to_none=True)
        xc = x.clone().requires_grad_(True)
        out = mod(xc)
        out.backward(go)
        torch.cuda.synchronize()
    finally:
        ops.set_max_ctas(prev)
        lib.sgcn_set_traversal(_lib.traversal_mode())
    res = {"out": out.detach().clone(), "gx": xc.grad.clone()}
    for k, p in mod.named_parameters():
        if p.grad is not None and not k.endswith("pos"):
            res["grad:" + k] = p.grad.clone()
    return res


@pytest.mark.parametrize("kind,C,D,V,n,T,stride", [
    ("gcn", 64, 64, 25, 8, 300, 1), ("gcn", 64, 128, 25, 4, 150, 1), ("gcn", 128, 256, 25, 4, 75, 1),
    ("gcn", 256, 256, 33, 2, 75, 1), ("tcn", 64, 64, 25, 8, 300, 1), ("tcn", 128, 128, 25, 4, 150, 2),
    ("tcn", 256, 256, 33, 4, 75, 1), ("unit", 64, 64, 25, 8, 300, 1), ("unit", 64, 128, 25, 4, 300, 2),
    ("unit", 128, 128, 33, 4, 150, 1), ("unit", 128, 256, 25, 4, 150, 2), ("unit", 256, 256, 25, 16, 75, 1),
    ("unit", 3, 64, 25, 4, 300, 1)])
def test_tiles_per_cta_do_not_change_results(cuda_device, kind, C, D, V, n, T, stride):
    from shiftgcn_b200.modules import Shift_gcn, Shift_tcn, TCN_GCN_unit
    torch.manual_seed(1)
    if kind == "gcn":
        mod = Shift_gcn(C, D, None, num_point=V)
    elif kind == "tcn":
        mod = Shift_tcn(C, C, stride=stride)
        D = C
    else:
        mod = TCN_GCN_unit(C, D, None, stride=stride, residual=C != 3, num_point=V)
    model_ref.fill_module_(mod)
    mod = mod.to(cuda_device)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(n, C, T, V, generator=g).to(cuda_device)
    go = torch.randn(n, D, T // stride, V, generator=g).to(cuda_device)
    base = _run(mod, x, go, 0, False, 1)
    for cap, first in ((3, 1), (3, 2), (7, 1), (61, 2)):
        got = _run(mod, x, go, cap, False, first)
        assert torch.equal(got["out"], base["out"]), f"output differs with {cap} CTAs"
        assert torch.equal(got["gx"], base["gx"]), f"input gradient differs with {cap} CTAs"
        for k in base:
            if k.startswith("grad:"):
                scale = base[k].abs().max().item()
                err = (got[k] - base[k]).abs().max().item()
                assert err <= 2e-4 * scale + 1e-6, f"{k} differs with {cap} CTAs: {err:.3e} of {scale:.3e}"
    # training mode: the per-CTA fp32 partial sums of the batch statistics depend on the split (1e-7 relative), which
    # moves a few TF32 roundings and ReLU decisions -- compared as vectors
    base = _run(mod, x, go, 0, True, 1)
    got = _run(mod, x, go, 5, True, 2)
    assert rel_l2_t(got["out"], base["out"]) < 1e-3
    assert rel_l2_t(got["gx"], base["gx"]) < 3e-2


def rel_l2_t(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_max_ctas_hook_is_restored(cuda_device):
    from shiftgcn_b200 import ops
    assert ops.set_max_ctas(0) == 0
