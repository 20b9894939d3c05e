"""Sliding-window inference on the GPU: sgcn_window_stream / sgcn_window_scores against the
golden vectors made by the reference's own functions, and the batched WindowedEnsemble against the reference's
window-by-window loop (inference_pipeline.py:342-366) restated with the oracle models."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref, modalities

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "windows.npz")
TAGS = ("short", "exact", "ragged", "one_over")


@pytest.mark.parametrize("tag", TAGS)
def test_window_stream_bit_exact_against_reference_golden(cuda_device, tag):
    from shiftgcn_b200 import ops
    from shiftgcn_b200.ensemble import bone_parents, stream_flags
    g = np.load(GOLDEN)
    seq = torch.from_numpy(g[f"{tag}/seq"]).to(cuda_device)
    meta = g[f"{tag}/meta"]
    start = torch.tensor(meta[:, 0], dtype=torch.int32, device=cuda_device)
    parents = torch.tensor(bone_parents(33), dtype=torch.int32, device=cuda_device)
    for name in modalities.MODALITIES:
        bone, motion = stream_flags(name)
        want = g[f"{tag}/windows"] if name == "joint" else g[f"{tag}/{name}"]
        got = ops.window_stream(seq, start, 8, parent=parents if bone else None, motion=motion)
        assert np.array_equal(got.cpu().numpy(), want), name
        rows = ops.window_stream(seq, start, 8, parent=parents if bone else None, motion=motion, rows=True)
        W, C, T, V, M = want.shape
        want_rows = np.transpose(want, (0, 4, 2, 3, 1)).reshape(W * M, T, V, C)
        assert np.array_equal(rows.cpu().numpy(), want_rows), name + " rows"


@pytest.mark.parametrize("tag", TAGS)
def test_window_scores_and_aggregation(cuda_device, tag):
    from shiftgcn_b200 import ops, pipeline as P
    g = np.load(GOLDEN)
    meta, T = g[f"{tag}/meta"], g[f"{tag}/seq"].shape[1]
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((len(meta), 2)).astype(np.float32) * 3
    start = torch.tensor(meta[:, 0], dtype=torch.int32, device=cuda_device)
    real = torch.tensor(meta[:, 2], dtype=torch.int32, device=cuda_device)
    score, per_frame = ops.window_scores(torch.from_numpy(logits).to(cuda_device), start, real, T)
    want_score = modalities.fall_scores(logits.astype(np.float64))
    assert np.abs(score.cpu().numpy() - want_score).max() < 1e-14
    results = [(float(s), int(m[0]), int(m[1]), int(m[2])) for s, m in zip(want_score, meta)]
    want_pf = modalities.aggregate_per_frame(results, T)
    assert np.abs(per_frame.cpu().numpy() - want_pf).max() < 1e-14
    # the drop-in aggregate_per_frame on the reference's own scores: equal to the golden per-frame vector
    results = [(float(s), int(m[0]), int(m[1]), int(m[2])) for s, m in zip(g[f"{tag}/scores"], meta)]
    got = P.aggregate_per_frame(results, T, device=cuda_device)
    assert np.abs(got - g[f"{tag}/per_frame"]).max() < 1e-15


def test_window_ops_empty_and_guards(cuda_device):
    from shiftgcn_b200 import ops
    seq = torch.randn(3, 5, 33, 1, device=cuda_device)
    none = ops.window_stream(seq, torch.zeros(0, dtype=torch.int32, device=cuda_device), 8)
    assert none.shape == (0, 3, 8, 33, 1)
    with pytest.raises(RuntimeError):
        ops.window_stream(seq[0], torch.zeros(1, dtype=torch.int32, device=cuda_device), 8)
    with pytest.raises(RuntimeError):
        ops.window_stream(seq, torch.zeros(1, dtype=torch.int64, device=cuda_device), 8)


def _four_models(device, V=33):
    from shiftgcn_b200.modules import Model
    mods, refs = {}, {}
    for k, name in enumerate(modalities.MODALITIES):
        torch.manual_seed(20 + k)
        m = Model(num_class=2, num_point=V, num_person=1, graph="graph.mediapipe_pose.Graph",
                  graph_args=dict(labeling_mode="spatial"))
        r = model_ref.RefModel(num_class=2, num_point=V, num_person=1)
        model_ref.fill_module_(m, prefix=name), model_ref.fill_module_(r, prefix=name)
        mods[name], refs[name] = m.to(device).eval(), r.double().eval()
    return mods, refs


def test_windowed_ensemble_matches_the_reference_loop(cuda_device):
    """whole sequence -> per-window fall scores and per-frame averages: one batched pass per stream on the GPU against
    the reference's loop (one window, one stream at a time) over the oracle models"""
    from shiftgcn_b200 import pipeline as P
    mods, refs = _four_models(cuda_device)
    rng = np.random.default_rng(3)
    T, win, stride = 53, 24, 12
    seq = rng.standard_normal((3, T, 33, 1)).astype(np.float32)
    ens = P.WindowedEnsemble(mods, window_size=win, stride=stride)
    results, per_frame = ens.score(seq)
    # reference loop (inference_pipeline.py:342-366) with the oracle pieces
    want = []
    with torch.no_grad():
        for w, start, end, real in modalities.sliding_windows(seq, win, stride):
            streams = modalities.derive(w[None])
            logits = [refs[name](torch.from_numpy(streams[name]).double()).numpy()[0] for name in modalities.MODALITIES]
            tot = modalities.ensemble_logits(logits)
            want.append((float(modalities.fall_scores(tot[None])[0]), start, end, real))
    assert [r[1:] for r in results] == [w[1:] for w in want]
    got_s, want_s = np.array([r[0] for r in results]), np.array([w[0] for w in want])
    assert np.abs(got_s - want_s).max() < 2e-2, (got_s, want_s)               # TF32 contractions, ten layers
    assert ((got_s > 0.5) == (want_s > 0.5)).all()
    assert np.abs(per_frame - modalities.aggregate_per_frame(want, T)).max() < 2e-2
    # the drop-in functions on materialised windows give the same numbers as the window-free path
    wins = P.create_sliding_windows(seq, win, stride)
    res2 = P.run_ensemble_inference(wins, mods, device=cuda_device)
    assert np.abs(np.array([r[0] for r in res2]) - got_s).max() < 1e-5
    pf2 = P.aggregate_per_frame(res2, T, device=cuda_device)
    assert np.abs(pf2 - per_frame).max() < 1e-5
