"""Batched sliding-window inference: the caller side of the hot path in ``inference_pipeline.py``.

The reference slices a pose sequence into overlapping windows on the host (``create_sliding_windows``, :252-281), derives
the four streams per window in numpy (:284-309), runs FOUR batch-1 forward passes per window with a ``.cpu()`` sync after
each (``run_ensemble_inference``, :342-366) and averages the window scores per frame in numpy (``aggregate_per_frame``,
:377-386).  Here the sequence is copied to the device ONCE; every stream of every window comes out of one kernel
(``sgcn_window_stream``: windowing + zero padding + bone / motion derivation + the model's input BatchNorm + row layout),
each model runs once on the whole window batch, the alpha-weighted logits are summed on the device (one all-reduce
when the streams are placed on several ranks, ``ensemble.StreamEnsemble``), and ``sgcn_window_scores`` turns them into
window scores and per-frame averages in fp64.

The module keeps the reference's function names, argument meaning and return values so that ``inference_pipeline.py``
can switch over by import:

    create_sliding_windows(data, window_size=300, stride=150) -> [(window, start, end, num_real), ...]
    run_ensemble_inference(windows, models, ensemble_weights, progress_callback=None) -> [(score, start, end, num_real)]
    aggregate_per_frame(window_results, total_frames) -> float64 array
    detect_fall_intervals(per_frame_scores, threshold, fps) -> [dict(start_frame, end_frame, start_time, end_time, ...)]

plus ``WindowedEnsemble``, which never materialises the windows at all.
"""
import numpy as np
import torch

from . import ops
from .ensemble import ENSEMBLE_WEIGHTS_DEFAULT, MODALITIES, bone_parents, derive_modality, stream_flags


def window_plan(total_frames, window_size=300, stride=150):
    """[(start, end, num_real)] of ``create_sliding_windows`` (inference_pipeline.py:252-281) without touching data:
    one zero-padded window for a short sequence, otherwise windows every ``stride`` frames up to and including the first
    one that reaches the end of the sequence."""
    T = int(total_frames)
    if T <= window_size:
        return [(0, T, T)]
    plan, start = [], 0
    while start < T:
        end = start + window_size
        if end <= T:
            plan.append((start, end, window_size))
        else:
            plan.append((start, T, T - start))
        start += stride
        if end >= T:
            break
    return plan


def create_sliding_windows(data, window_size=300, stride=150):
    """Drop-in for inference_pipeline.py:252-281: data (C, T, V, M) numpy -> list of (window, start, end, num_real)."""
    data = np.asarray(data)
    C, T, V, M = data.shape
    out = []
    for start, end, real in window_plan(T, window_size, stride):
        if real == window_size:
            w = data[:, start:start + window_size].copy()
        else:
            w = np.zeros((C, window_size, V, M), dtype=np.float32)
            w[:, :real] = data[:, start:start + real]
        out.append((w, start, end, real))
    return out


def _stream_batch(joint, name, parents):
    return derive_modality(joint, name, parents)


def _ensemble_logits(models, streams, weights):
    """sum_k alpha_k * model_k(stream_k) on the device; ``streams``: name -> stream batch"""
    total = None
    for name, alpha in zip(MODALITIES, weights):
        out = models[name](streams[name]).float()
        total = out * alpha if total is None else total.add_(out, alpha=alpha)
    return total


@torch.no_grad()
def run_ensemble_inference(windows, models, ensemble_weights=ENSEMBLE_WEIGHTS_DEFAULT, progress_callback=None, device=None):
    """Drop-in for inference_pipeline.py:342-366, batched: ALL windows go through each model in one call.

    windows: the list ``create_sliding_windows`` returns; models: modality name -> model (on ``device``)."""
    if not windows:
        return []
    device = torch.device(device) if device is not None else _model_device(models)
    joint = torch.from_numpy(np.stack([w[0] for w in windows]).astype(np.float32)).to(device, non_blocking=True)
    streams = {name: _stream_batch(joint, name, None) for name in MODALITIES}
    logits = _ensemble_logits(models, streams, ensemble_weights)
    start = torch.tensor([w[1] for w in windows], dtype=torch.int32, device=device)
    real = torch.tensor([w[3] for w in windows], dtype=torch.int32, device=device)
    score, _ = ops.window_scores(logits.contiguous(), start, real, 0, cls=1)
    score = score.cpu().numpy()
    results = [(float(s), w[1], w[2], w[3]) for s, w in zip(score, windows)]
    if progress_callback:
        progress_callback(len(windows), len(windows))
    return results


def _model_device(models):
    for m in models.values():
        for p in m.parameters():
            return p.device
    raise RuntimeError("run_ensemble_inference: cannot infer the device of the models; pass device=")


def aggregate_per_frame(window_results, total_frames, device="cuda"):
    """Drop-in for inference_pipeline.py:377-386 on the device (sgcn_window_scores' aggregation kernel): the per-frame
    mean of the scores of the windows that cover the frame with real data; 0 where no window does."""
    W = len(window_results)
    dev = torch.device(device)
    score = torch.tensor([r[0] for r in window_results], dtype=torch.float64, device=dev)
    start = torch.tensor([r[1] for r in window_results], dtype=torch.int32, device=dev)
    real = torch.tensor([r[3] for r in window_results], dtype=torch.int32, device=dev)
    return ops.frame_aggregate(score, start, real, int(total_frames)).cpu().numpy()


def detect_fall_intervals(per_frame_scores, threshold, fps):
    """inference_pipeline.py:389-424: contiguous runs of frames whose score exceeds the threshold, with the reference's
    report fields (host side; the run edges are found with one vectorised comparison instead of a Python frame loop)."""
    s = np.asarray(per_frame_scores, dtype=np.float64)
    above = np.concatenate([[False], s > threshold, [False]])
    edges = np.flatnonzero(above[1:] != above[:-1])

    def fmt_time(frame):
        secs = frame / fps
        mins = int(secs // 60)
        return f"{mins}:{secs % 60:05.2f}"

    out = []
    for a, b in zip(edges[0::2].tolist(), edges[1::2].tolist()):
        seg = s[a:b]
        out.append({"start_frame": int(a), "end_frame": int(b), "start_time": fmt_time(a), "end_time": fmt_time(b),
                    "mean_confidence": float(seg.mean()), "peak_confidence": float(seg.max()),
                    "peak_frame": int(a + int(np.argmax(seg)))})
    return out


class WindowedEnsemble:
    """Whole-sequence fall scoring without materialised windows.

    models: modality name -> Model (eval mode) on ``device``.  ``score(sequence)``: sequence (C, T, V, M) numpy or
    tensor -> (window_results, per_frame float64 numpy) with the reference's meaning.  Each stream of all windows is
    produced by ONE ``sgcn_window_stream`` launch from the resident sequence and consumed by ONE forward call.
    """

    def __init__(self, models, weights=ENSEMBLE_WEIGHTS_DEFAULT, window_size=300, stride=150, fall_class=1):
        if len(weights) != 4:
            raise ValueError("the ensemble has four stream weights")
        self.models, self.weights = models, tuple(float(w) for w in weights)
        self.window_size, self.stride, self.fall_class = int(window_size), int(stride), int(fall_class)
        self.device = _model_device(models)
        self._parents = {}

    def _parent_table(self, V):
        if V not in self._parents:
            self._parents[V] = torch.tensor(bone_parents(V), dtype=torch.int32, device=self.device)
        return self._parents[V]

    @torch.no_grad()
    def logits(self, sequence):
        seq = torch.as_tensor(sequence, dtype=torch.float32)
        if seq.dim() != 4:
            raise RuntimeError("WindowedEnsemble expects a sequence of shape (C, T, V, M)")
        pinned = seq.device.type == "cpu"
        seq = seq.contiguous()
        if pinned:
            seq = seq.pin_memory().to(self.device, non_blocking=True)
        C, T, V, M = seq.shape
        plan = window_plan(T, self.window_size, self.stride)
        start = torch.tensor([p[0] for p in plan], dtype=torch.int32, device=self.device)
        total = None
        for name, alpha in zip(MODALITIES, self.weights):
            use_bone, motion = stream_flags(name)
            stream = ops.window_stream(seq, start, self.window_size, parent=self._parent_table(V) if use_bone else None,
                                       motion=motion)
            out = self.models[name](stream).float()
            total = out * alpha if total is None else total.add_(out, alpha=alpha)
        return total, plan, T

    @torch.no_grad()
    def score(self, sequence):
        logits, plan, T = self.logits(sequence)
        start = torch.tensor([p[0] for p in plan], dtype=torch.int32, device=self.device)
        real = torch.tensor([p[2] for p in plan], dtype=torch.int32, device=self.device)
        score, per_frame = ops.window_scores(logits.contiguous(), start, real, T, cls=self.fall_class)
        score = score.cpu().numpy()
        results = [(float(s), a, b, r) for s, (a, b, r) in zip(score, plan)]
        return results, per_frame.cpu().numpy()
