"""Feeder -> GPU pipeline on the device: sgcn_random_move against golden vectors made by the
reference's own feeders/tools.py random_move, and DeviceFeeder against the reference's synchronous loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feeder.npz")


@pytest.mark.parametrize("tag", ["ntu", "mp", "two_seg"])
def test_random_move_kernel_against_reference_golden(cuda_device, tag):
    from shiftgcn_b200 import ops
    g = np.load(GOLDEN)
    x = torch.from_numpy(g[f"{tag}/x"]).to(cuda_device)
    vals = torch.from_numpy(g[f"{tag}/vals"]).to(cuda_device)
    node = torch.from_numpy(g[f"{tag}/node"]).to(cuda_device)
    got = ops.random_move_(x.clone(), vals, node).cpu().numpy()
    want = g[f"{tag}/out"]
    # fp64 arithmetic, one rounding to fp32: identical up to the last bit of sin / cos of the two math libraries
    assert np.abs(got - want).max() <= 2.4e-7 * np.abs(want).max()
    assert np.mean(got == want) > 0.99
    assert np.array_equal(got[:, 2], g[f"{tag}/x"][:, 2])           # the z channel is untouched


def test_device_feeder_matches_the_synchronous_loop(cuda_device):
    from oracle import feeder_tools
    from shiftgcn_b200.feeder import DeviceFeeder
    rng = np.random.default_rng(4)
    batches = [(torch.from_numpy(rng.standard_normal((3, 3, 12, 25, 2)).astype(np.float32)),
                torch.from_numpy(rng.integers(0, 60, 3)), torch.arange(3) + 3 * i) for i in range(4)]
    # reference order of operations: per-sample random_move in the Dataset, then data.float().cuda() (main.py:400-402)
    np.random.seed(77)
    want = []
    for data, label, _ in batches:
        moved = []
        for n in range(data.shape[0]):
            node, vals = feeder_tools.move_nodes(12, 1)
            moved.append(feeder_tools.apply_move(data[n].numpy(), node, vals))
        want.append((np.stack(moved), label.numpy()))
    np.random.seed(77)
    feeder = DeviceFeeder(batches, cuda_device, random_move=True, depth=2)
    seen = 0
    for (d, l, idx), (wd, wl), (_, _, widx) in zip(feeder, want, batches):
        d = d * 1.0                                                   # consume on the current stream
        assert d.is_cuda and d.dtype == torch.float32 and l.dtype == torch.int64
        assert np.abs(d.cpu().numpy() - wd).max() <= 2.4e-7 * np.abs(wd).max()
        assert np.array_equal(l.cpu().numpy(), wl) and torch.equal(idx, widx)
        seen += 1
    assert seen == 4 and feeder.batches == 4
    plain = DeviceFeeder(batches, cuda_device, depth=3)
    for (d, l, _), (data, label, _) in zip(plain, batches):
        assert torch.equal(d.cpu(), data) and torch.equal(l.cpu(), label)
