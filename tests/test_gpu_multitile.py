"""Parity of the STEADY STATE of the persistent tensor-core kernels (fused_gemm.cuh, spatial_bwd.cu, wgrad.cu).

Every GEMM-type kernel launches min(#SM, #tiles) persistent CTAs that walk tile = blockIdx.x + i * gridDim.x.  The
bench runs ~52 tiles per CTA; the small parity cases of test_gpu_units.py give every CTA exactly one tile.  Here every
CTA processes many tiles, so the second TMEM accumulator, the accumulator / operand / raw / weight ring wrap across
tiles, the mbarrier phase parities beyond the first use, wgrad's cross-tile accumulation in TMEM, partial last tiles
and both traversal directions are all compared against the oracle:

  * capped grids (sgcn_set_max_ctas: 3 and 7 CTAs) on small tensors -> 5..60 tiles per CTA, incl. a partial last tile;
  * the benchmark's own sequence lengths (n = 8 samples, T = 300 / 150 / 75 -> up to 2400 groups = 480 tiles of 5
    groups, >= 3 per CTA on 148 SMs) with the full grid.
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
    _compare(mod, ref, x, go, train, cuda_device)


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
    _compare(mod, ref, x, go, train, cuda_device)


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
    _compare(mod, ref, x, go, train, cuda_device)


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
    _compare(mod, ref, x, go, train, cuda_device)


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
    _compare(mod, ref, x, go, True, cuda_device)


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
    _compare(mod, ref, x, go, False, cuda_device)


def test_max_ctas_hook_is_restored(cuda_device):
    from shiftgcn_b200 import ops
    assert ops.set_max_ctas(0) == 0
