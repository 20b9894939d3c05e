"""Sliding-window inference, host side: the oracle restatement and the product's window plan /
interval detection against golden vectors made by EXECUTING the reference's own create_sliding_windows,
aggregate_per_frame and detect_fall_intervals (inference_pipeline.py:252-281, 377-424; oracle/make_golden.py step 8)."""
import json
import os

import numpy as np
import pytest

from oracle import modalities
from shiftgcn_b200 import pipeline as P

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "windows.npz")
TAGS = ("short", "exact", "ragged", "one_over")


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_windows_equal_reference_golden(tag):
    g = np.load(GOLDEN)
    wins = modalities.sliding_windows(g[f"{tag}/seq"], window_size=8, stride=4)
    assert np.array_equal(np.stack([w[0] for w in wins]), g[f"{tag}/windows"])
    assert [[w[1], w[2], w[3]] for w in wins] == g[f"{tag}/meta"].tolist()
    results = [(float(s), w[1], w[2], w[3]) for s, w in zip(g[f"{tag}/scores"], wins)]
    assert np.array_equal(modalities.aggregate_per_frame(results, g[f"{tag}/seq"].shape[1]), g[f"{tag}/per_frame"])
    streams = [modalities.derive(w[0][None])for w in wins]
    for name in modalities.MODALITIES[1:]:
        assert np.array_equal(np.concatenate([s[name] for s in streams]), g[f"{tag}/{name}"]), name


@pytest.mark.parametrize("tag", TAGS)
def test_product_window_plan_and_windows_equal_reference_golden(tag):
    g = np.load(GOLDEN)
    seq = g[f"{tag}/seq"]
    assert [list(p) for p in P.window_plan(seq.shape[1], 8, 4)] == g[f"{tag}/meta"].tolist()
    wins = P.create_sliding_windows(seq, window_size=8, stride=4)
    assert np.array_equal(np.stack([w[0] for w in wins]), g[f"{tag}/windows"])
    assert [[w[1], w[2], w[3]] for w in wins] == g[f"{tag}/meta"].tolist()
    assert all(w[0].dtype == np.float32 for w in wins)


def test_window_plan_edge_cases():
    assert P.window_plan(300) == [(0, 300, 300)]                       # exactly one window, nothing padded
    assert P.window_plan(1) == [(0, 1, 1)]
    assert P.window_plan(301) == [(0, 300, 300), (150, 301, 151)]
    assert P.window_plan(450) == [(0, 300, 300), (150, 450, 300)]       # second window ends exactly at T: loop stops
    plan = P.window_plan(10_000)
    assert plan[0] == (0, 300, 300) and plan[-1][1] == 10_000 and all(b[0] - a[0] == 150 for a, b in zip(plan, plan[1:]))


@pytest.mark.parametrize("tag", TAGS)
def test_detect_fall_intervals_equal_reference_golden(tag):
    g = np.load(GOLDEN)
    want = json.load(open(os.path.join(HERE, "golden", "windows_detections.json")))[tag]
    got = P.detect_fall_intervals(g[f"{tag}/per_frame"], 0.5, 30.0)
    assert got == want


def test_detect_fall_intervals_edges():
    s = np.array([0.9, 0.9, 0.1, 0.6, 0.7, 0.2, 0.8])
    got = P.detect_fall_intervals(s, 0.5, 10.0)
    assert [(d["start_frame"], d["end_frame"], d["peak_frame"]) for d in got] == [(0, 2, 0), (3, 5, 4), (6, 7, 6)]
    assert P.detect_fall_intervals(np.zeros(5), 0.5, 10.0) == []
    assert P.detect_fall_intervals(np.zeros(0), 0.5, 10.0) == []
