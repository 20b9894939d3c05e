"""Feeder -> GPU pipeline, host side: the oracle restatement of feeders/tools.py random_move
against golden vectors made by the reference's own function, the node draws, and DeviceFeeder's batching logic on CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import feeder_tools
from shiftgcn_b200 import feeder as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feeder.npz")
TAGS = {"ntu": 1, "mp": 1, "two_seg": 2}


@pytest.mark.parametrize("tag", sorted(TAGS))
def test_oracle_random_move_equals_reference_golden(tag):
    g = np.load(GOLDEN)
    x, node, vals = g[f"{tag}/x"], g[f"{tag}/node"], g[f"{tag}/vals"]
    for n in range(x.shape[0]):
        got = feeder_tools.apply_move(x[n], node, vals[n])
        assert got.dtype == np.float32 and np.array_equal(got, g[f"{tag}/out"][n])


@pytest.mark.parametrize("tag", sorted(TAGS))
def test_node_draws_replay_the_reference_rng(tag):
    g = np.load(GOLDEN)
    T = g[f"{tag}/x"].shape[2]
    for n in range(4):
        np.random.seed(100 + n)
        node, vals = F.draw_move_nodes(1, T, TAGS[tag])
        assert np.array_equal(node, g[f"{tag}/node"]) and np.array_equal(vals[0], g[f"{tag}/vals"][n])
        np.random.seed(100 + n)
        node2, vals2 = feeder_tools.move_nodes(T, TAGS[tag])
        assert np.array_equal(node2, node) and np.array_equal(vals2, vals[0])


def test_device_feeder_yields_every_batch_in_order_on_cpu():
    batches = [(torch.full((2, 3, 4, 5, 1), float(i), dtype=torch.float64), torch.tensor([i, i + 1], dtype=torch.int32),
                torch.tensor([2 * i, 2 * i + 1])) for i in range(5)]
    for depth in (1, 2, 4, 9):
        f = F.DeviceFeeder(batches, "cpu", depth=depth)
        got = list(f)
        assert len(got) == 5 and f.batches == 5 and len(f) == 5
        for i, (d, l, idx) in enumerate(got):
            assert d.dtype == torch.float32 and l.dtype == torch.int64
            assert torch.equal(d, batches[i][0].float()) and torch.equal(l, batches[i][1].long())
            assert torch.equal(idx, batches[i][2])
    assert list(F.DeviceFeeder([], "cpu")) == []
    with pytest.raises(RuntimeError):
        F.DeviceFeeder(batches, "cpu", random_move=True)
    with pytest.raises(ValueError):
        F.DeviceFeeder(batches, "cpu", depth=0)
