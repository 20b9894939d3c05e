"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's stream derivation and ensemble sum.

Follows, statement by statement:
  * ``data_gen/gen_bone_data.py:5-31`` (the ``paris`` table, 1-based (joint, parent) pairs, identical for every NTU
    benchmark) and ``:52-58`` (copy the joints, then ``bone[v1] = joint[v1] - joint[v2]`` for every pair);
  * ``data_gen/gen_motion_data.py:30-34`` (``motion[t] = data[t+1] - data[t]`` for ``t < T-1``, last frame 0);
  * ``inference_pipeline.py:16-22`` (``BONE_PAIRS``, 0-based, MediaPipe) and ``:284-309`` (``derive_modalities``);
  * ``ensemble.py:18-27`` / ``inference_pipeline.py:352-360`` (``sum_k alpha_k * logits_k``, softmax after the sum).
Pinned by tests/golden/modalities.npz, which oracle/make_golden.py writes by EXECUTING the reference's own
``derive_modalities`` (extracted from its source file, because importing the module needs cv2 / mediapipe) and the
reference's own pair tables.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import ast
import os

import numpy as np

NTU_PAIRS = ((1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10), (12, 11),
             (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 23), (21, 21), (23, 8),
             (24, 25), (25, 12))
MEDIAPIPE_PAIRS = ((0, 0), (1, 0), (2, 1), (3, 2), (4, 0), (5, 4), (6, 5), (7, 3), (8, 6), (9, 0), (10, 9), (11, 0),
                   (12, 11), (13, 11), (14, 12), (15, 13), (16, 14), (17, 15), (18, 16), (19, 15), (20, 16), (21, 15),
                   (22, 16), (23, 11), (24, 12), (25, 23), (26, 24), (27, 25), (28, 26), (29, 27), (30, 28), (31, 27),
                   (32, 28))
MODALITIES = ("joint", "bone", "joint_motion", "bone_motion")
ALPHA = (0.6, 0.6, 0.4, 0.4)


def pairs_0based(V):
    if V == 25:
        return tuple((a - 1, b - 1) for a, b in NTU_PAIRS)
    if V == 33:
        return MEDIAPIPE_PAIRS
    raise ValueError(V)


def bone(data, pairs):
    """data (N, C, T, V, M); gen_bone_data.py:52-58 / inference_pipeline.py:292-294"""
    out = data.copy()                                           # fp_sp[:, :C] = data
    for v1, v2 in pairs:
        out[:, :, :, v1, :] = data[:, :, :, v1, :] - data[:, :, :, v2, :]
    return out


def motion(data):
    """gen_motion_data.py:30-34 / inference_pipeline.py:297-298"""
    out = np.zeros_like(data)
    T = data.shape[2]
    for t in range(T - 1):
        out[:, :, t] = data[:, :, t + 1] - data[:, :, t]
    return out


def derive(joint):
    """all four streams of a joint batch (N, C, T, V, M)"""
    b = bone(joint, pairs_0based(joint.shape[3]))
    return {"joint": joint, "bone": b, "joint_motion": motion(joint), "bone_motion": motion(b)}


def ensemble_logits(logits, alpha=ALPHA):
    """ensemble.py:26 -- r11*alpha[0] + r22*alpha[1] + r33*alpha[2] + r44*alpha[3]"""
    return sum(a * l for a, l in zip(alpha, logits))


def fall_scores(logits):
    """inference_pipeline.py:358-360 on a batch: softmax of the ensemble logits, class 1"""
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return e[:, 1] / e.sum(axis=1)


def sliding_windows(data, window_size=300, stride=150):
    """inference_pipeline.py:252-281 -- (C, T, V, M) -> [(window, start, end, num_real)], short sequences and the last
    ragged window zero padded; the loop stops after the first window that reaches the end of the sequence"""
    C, T, V, M = data.shape
    if T <= window_size:
        padded = np.zeros((C, window_size, V, M), dtype=np.float32)
        padded[:, :T] = data
        return [(padded, 0, T, T)]
    windows, start = [], 0
    while start < T:
        end = start + window_size
        if end <= T:
            windows.append((data[:, start:end].copy(), start, end, window_size))
        else:
            padded = np.zeros((C, window_size, V, M), dtype=np.float32)
            padded[:, :T - start] = data[:, start:T]
            windows.append((padded, start, T, T - start))
        start += stride
        if end >= T:
            break
    return windows


def aggregate_per_frame(window_results, total_frames):
    """inference_pipeline.py:377-386 -- per-frame mean of the window scores over the REAL frames of each window"""
    score_sum = np.zeros(total_frames, dtype=np.float64)
    score_count = np.zeros(total_frames, dtype=np.float64)
    for fall_score, start, end, num_real in window_results:
        score_sum[start:start + num_real] += fall_score
        score_count[start:start + num_real] += 1.0
    return score_sum / np.maximum(score_count, 1.0)


# ------------------------------------------------------------------ the reference's own code (build container only)
def reference_window_functions(root="/root/reference"):
    """(create_sliding_windows, aggregate_per_frame, detect_fall_intervals) executed from the reference SOURCE with ast
    (see reference_objects for why it cannot be imported); used by oracle/make_golden.py only"""
    src = open(os.path.join(root, "inference_pipeline.py")).read()
    want = ("create_sliding_windows", "aggregate_per_frame", "detect_fall_intervals", "_add_detection")
    keep = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "inference_pipeline.py", "exec"), ns)
    return ns["create_sliding_windows"], ns["aggregate_per_frame"], ns["detect_fall_intervals"]


# ------------------------------------------------------------------ the reference's own code (build container only)
def reference_objects(root="/root/reference"):
    """(derive_modalities function, BONE_PAIRS, NTU `paris` dict) taken from the reference SOURCE FILES with ast --
    ``inference_pipeline.py`` cannot be imported here (cv2 / mediapipe are absent) and ``gen_bone_data.py`` is a script
    that rewrites dataset files on import.  Nothing is copied into this repository; the objects live in memory only."""
    src = open(os.path.join(root, "inference_pipeline.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body
            if (isinstance(n, ast.Assign) and any(getattr(t, "id", None) in ("BONE_PAIRS", "MODALITIES") for t in n.targets))
            or (isinstance(n, ast.FunctionDef) and n.name == "derive_modalities")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "inference_pipeline.py", "exec"), ns)
    src2 = open(os.path.join(root, "data_gen", "gen_bone_data.py")).read()
    paris = None
    for n in ast.parse(src2).body:
        if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "paris":
            paris = ast.literal_eval(n.value)
    return ns["derive_modalities"], ns["BONE_PAIRS"], paris
