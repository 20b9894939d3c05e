"""Pins the tcgen05 descriptor / canonical-tile conventions of csrc/common.cuh on real hardware."""
import pytest
import torch

from oracle.model_ref import tf32_round
from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 32), (32, 64), (64, 64), (64, 128), (128, 128), (64, 256), (128, 256), (256, 64)])
def test_row_tile_times_weight(cuda_device, K, N):
    from shiftgcn_b200 import ops
    torch.manual_seed(K * 1000 + N)
    a = torch.randn(128, K, device=cuda_device)
    b = torch.randn(N, K, device=cuda_device)
    d = ops.selftest_umma(a, b, 0, K, N)
    want = tf32_round(a).double().cpu() @ tf32_round(b).double().cpu().t()
    assert rel_err(d, want) < 2e-6
    exact = a.double().cpu() @ b.double().cpu().t()
    assert rel_err(d, exact) < 1e-2      # TF32 tolerance of the north star


@pytest.mark.parametrize("M,N", [(64, 64), (128, 64), (128, 128), (64, 128), (32, 64)])
def test_weight_gradient_shape(cuda_device, M, N):
    from shiftgcn_b200 import ops
    torch.manual_seed(M * 1000 + N)
    a = torch.randn(128, M, device=cuda_device)
    b = torch.randn(128, N, device=cuda_device)
    d = ops.selftest_umma(a, b, 1, 0, N, M)
    want = tf32_round(a).double().cpu().t() @ tf32_round(b).double().cpu()
    assert rel_err(d[:M], want) < 2e-6
