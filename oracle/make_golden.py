"""Generate tests/golden/* from the REAL reference (TEST INFRASTRUCTURE ONLY; build container only).

    python -m oracle.make_golden

imports /root/reference through oracle/ref_import.py (CPU, fp64), gives every module the deterministic
non-degenerate weights of oracle.model_ref.fill_value, and stores inputs + outputs + gradients.  The reference
ships no fixtures of its own, so these files are what pins parity on the GPU box, where
/root/reference does not exist.  The temporal shift inside Shift_tcn / TCN_GCN_unit / Model comes from the oracle
restatement (the reference's CUDA extension cannot run here); that restatement is pinned separately against the
reference's compiled kernels on the GPU (tests/test_gpu_reference_ext.py) and against the C oracle
(``shift_op.npz`` below is produced by oracle/shift_oracle.c).
"""
import json
import os

import numpy as np
import torch

from . import modalities, model_ref, ref_import, shift_c, shift_torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _run(module, x, go, train):
    module.train(train)
    xr = x.clone().requires_grad_(True)
    shift_torch.RAW_POS_LOG = {}
    out = module(xr)
    out.backward(go)
    log, shift_torch.RAW_POS_LOG = shift_torch.RAW_POS_LOG, None
    rec = {"x": _np(x).astype(np.float32), "go": _np(go).astype(np.float32), "out": _np(out), "gx": _np(xr.grad)}
    for k, p in module.named_parameters():
        if p.grad is not None:
            rec["grad/" + k] = _np(p.grad)
        if k.endswith("xpos") and id(p) in log:
            rec["raw/" + k[:-4] + "ypos"] = _np(log[id(p)][1])
    if train:
        for k, b in module.named_buffers():
            rec["buf/" + k] = _np(b)
    return rec


def _f32_inputs(shape_x, shape_go, seed):
    g = torch.Generator().manual_seed(seed)
    # inputs are stored (and used) as exactly-representable fp32 values
    return (torch.randn(*shape_x, generator=g).float().double(), torch.randn(*shape_go, generator=g).float().double())


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    torch.manual_seed(1)

    # ---- 1. integer gather tables, straight from reference Shift_gcn instances (bit-exact contract)
    tables = {}
    for (V, C, D) in [(25, 3, 64), (25, 64, 64), (25, 64, 128), (33, 64, 64), (33, 128, 256)]:
        m = ref_import.build(ns.Shift_gcn, C, D, None, num_point=V)
        tables[f"in_{V}_{C}_{D}"] = m.shift_in.data.numpy().astype(np.int64)
        tables[f"out_{V}_{C}_{D}"] = m.shift_out.data.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tables)

    # ---- 2. state_dict contract of the full models
    contract = {}
    for tag, kw in (("ntu60", dict(num_class=60, num_point=25, num_person=2, graph=ns.GRAPH_NTU)),
                    ("mediapipe", dict(num_class=2, num_point=33, num_person=1, graph=ns.GRAPH_MEDIAPIPE))):
        m = ref_import.build(ns.Model, graph_args=dict(labeling_mode="spatial"), **kw)
        contract[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    with open(os.path.join(OUT, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f)

    # ---- 3. Shift_gcn: pure reference code
    cases = {"gcn_64_64_v25": (64, 64, 25, 2, 6), "gcn_64_128_v25": (64, 128, 25, 1, 5), "gcn_128_128_v33": (128, 128, 33, 1, 4)}
    for tag, (C, D, V, n, T) in cases.items():
        for train in (True, False):
            m = model_ref.fill_module_(ref_import.build(ns.Shift_gcn, C, D, None, num_point=V)).double()
            x, go = _f32_inputs((n, C, T, V), (n, D, T, V), 100 + C + D + V)
            rec = _run(m, x, go, train)
            rec["meta"] = np.array([C, D, V, n, T, int(train)])
            np.savez_compressed(os.path.join(OUT, f"{tag}_{'train' if train else 'eval'}.npz"), **rec)

    # ---- 4. TCN_GCN_unit (reference module code + oracle shift)
    ucases = {"unit_64_64_s1": (64, 64, 25, 2, 8, 1, True), "unit_64_128_s2": (64, 128, 25, 2, 8, 2, True)}
    for tag, (C, D, V, n, T, s, res) in ucases.items():
        m = model_ref.fill_module_(ref_import.build(ns.TCN_GCN_unit, C, D, None, stride=s, residual=res, num_point=V)).double()
        x, go = _f32_inputs((n, C, T, V), (n, D, T // s, V), 200 + C + D + s)
        rec = _run(m, x, go, True)
        rec["meta"] = np.array([C, D, V, n, T, s, int(res)])
        np.savez_compressed(os.path.join(OUT, f"{tag}_train.npz"), **rec)

    # ---- 5. full NTU-60 model, eval, with calibrated BN buffers stored alongside
    m = model_ref.fill_module_(ref_import.build(ns.Model, num_class=60, num_point=25, num_person=2, graph=ns.GRAPH_NTU,
                                                graph_args=dict(labeling_mode="spatial"))).double()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 16, 25, 2, generator=g).float().double()
    bns = [mod for mod in m.modules() if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm)]
    for b in bns:
        b.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(x)
    for b in bns:
        b.momentum = 0.1
    with torch.no_grad():                      # the fixture stores the buffers in fp32: evaluate with exactly those values
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.copy_(b.float().double())
    m.eval()
    with torch.no_grad():
        logits = m(x)
    rec = {"x": _np(x).astype(np.float32), "logits": _np(logits)}
    for k, b in m.named_buffers():
        if b.dtype.is_floating_point:
            rec["buf/" + k] = _np(b).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "model_ntu60_eval.npz"), **rec)

    # ---- 6. the stand-alone op, from the scalar C oracle
    rng = np.random.default_rng(6)
    rec = {}
    for stride in (1, 2):
        n, c, h, w = 2, 6, 9, 5
        x = rng.standard_normal((n, c, h, w)).astype(np.float32).astype(np.float64)
        go = rng.standard_normal((n, c, h // stride, w)).astype(np.float32).astype(np.float64)
        xpos = (rng.uniform(-1e-8, 1e-8, c)).astype(np.float32).astype(np.float64)
        ypos = np.array([0.3, -1.7, 2.0, -3.0, 5.5, 0.0]) + (0.5 if stride != 1 else 0.0)
        out = shift_c.shift_forward(x, xpos, ypos, stride)
        gin = shift_c.shift_backward_input(go, xpos, ypos, x.shape, stride)
        rx, ry = shift_c.shift_backward_pos_raw(x, go, xpos, ypos, stride)
        _, gy = shift_c.shift_constraint(rx, ry)
        for k, v in dict(x=x, go=go, xpos=xpos, ypos=ypos, out=out, gin=gin, raw_y=ry, gy=gy).items():
            rec[f"s{stride}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "shift_op.npz"), **rec)
    # ---- 7. ensemble streams: the reference's own derive_modalities (inference_pipeline.py:284-309, executed from its
    #         source) on MediaPipe windows; for NTU the reference only has file-rewriting scripts, so its pair table
    #         (gen_bone_data.py:5-31) is read from the source and applied with the statements of :52-58 / :30-34
    ref_derive, ref_pairs, ref_paris = modalities.reference_objects()
    assert tuple(map(tuple, ref_pairs)) == modalities.MEDIAPIPE_PAIRS
    assert all(tuple(v) == modalities.NTU_PAIRS for v in ref_paris.values())
    rng = np.random.default_rng(7)
    rec = {}
    jm = rng.standard_normal((3, 3, 9, 33, 1)).astype(np.float32)          # three windows (C, T, V, M)
    rec["mp/joint"] = jm
    per_window = [ref_derive(w) for w in jm]
    for name in modalities.MODALITIES[1:]:
        rec["mp/" + name] = np.stack([d[name] for d in per_window])
    jn = rng.standard_normal((2, 3, 7, 25, 2)).astype(np.float32)
    bone = jn.copy()
    for v1, v2 in ref_paris["ntu/xview"]:
        v1 -= 1
        v2 -= 1
        bone[:, :, :, v1, :] = jn[:, :, :, v1, :] - jn[:, :, :, v2, :]

    def _motion(data):
        out = np.zeros_like(data)
        T = data.shape[2]
        for t in range(T - 1):
            out[:, :, t, :, :] = data[:, :, t + 1, :, :] - data[:, :, t, :, :]
        out[:, :, T - 1, :, :] = 0
        return out

    rec["ntu/joint"], rec["ntu/bone"] = jn, bone
    rec["ntu/joint_motion"], rec["ntu/bone_motion"] = _motion(jn), _motion(bone)
    np.savez_compressed(os.path.join(OUT, "modalities.npz"), **rec)
    # ---- 8. sliding windows and per-frame aggregation: the reference's own create_sliding_windows / aggregate_per_frame /
    #         detect_fall_intervals (inference_pipeline.py:252-281, 377-424, executed from its source) on a short, an
    #         exact and a ragged sequence
    ref_windows, ref_aggregate, ref_detect = modalities.reference_window_functions()
    rng = np.random.default_rng(11)
    rec, det = {}, {}
    for tag, T in (("short", 7), ("exact", 20), ("ragged", 23), ("one_over", 9)):
        seq = rng.standard_normal((3, T, 33, 1)).astype(np.float32)
        wins = ref_windows(seq, window_size=8, stride=4)
        rec[f"{tag}/seq"] = seq
        rec[f"{tag}/windows"] = np.stack([w[0] for w in wins])
        rec[f"{tag}/meta"] = np.array([[w[1], w[2], w[3]] for w in wins], dtype=np.int64)
        scores = rng.uniform(0, 1, len(wins))
        results = [(float(s), w[1], w[2], w[3]) for s, w in zip(scores, wins)]
        rec[f"{tag}/scores"] = scores
        rec[f"{tag}/per_frame"] = ref_aggregate(results, T)
        det[tag] = ref_detect(rec[f"{tag}/per_frame"], 0.5, 30.0)
        # the four streams of the padded windows, from the reference's own derive_modalities
        per_window = [ref_derive(w[0]) for w in wins]
        for name in modalities.MODALITIES[1:]:
            rec[f"{tag}/{name}"] = np.stack([d[name] for d in per_window])
    np.savez_compressed(os.path.join(OUT, "windows.npz"), **rec)
    with open(os.path.join(OUT, "windows_detections.json"), "w") as f:
        json.dump(det, f, indent=1)
    # ---- 9. feeder augmentation: the reference's own feeders/tools.py random_move (imported; it needs only numpy) on
    #         seeded draws; the node values are recovered by replaying the same np.random calls with the same seed
    import importlib.util
    from . import feeder_tools
    spec = importlib.util.spec_from_file_location("ref_feeder_tools", "/root/reference/feeders/tools.py")
    ref_tools = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_tools)
    rng = np.random.default_rng(13)
    rec = {}
    for tag, (T, V, M, move_time) in {"ntu": (20, 25, 2, [1]), "mp": (13, 33, 1, [1]), "two_seg": (17, 25, 2, [2])}.items():
        x = rng.standard_normal((4, 3, T, V, M)).astype(np.float32)
        outs, nodes, vals = [], None, []
        for n in range(4):
            np.random.seed(100 + n)
            outs.append(ref_tools.random_move(x[n].copy(), move_time_candidate=move_time))
            np.random.seed(100 + n)
            nodes, v = feeder_tools.move_nodes(T, move_time[0])
            vals.append(v)
        rec[f"{tag}/x"], rec[f"{tag}/out"] = x, np.stack(outs)
        rec[f"{tag}/node"], rec[f"{tag}/vals"] = nodes.astype(np.int32), np.stack(vals)
    np.savez_compressed(os.path.join(OUT, "feeder.npz"), **rec)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {total / 1e6:.2f} MB -> {OUT}")


if __name__ == "__main__":
    main()
