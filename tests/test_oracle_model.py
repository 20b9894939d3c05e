"""CPU tests that pin oracle/model_ref.py: against the committed reference outputs (tests/golden, always) and
against the live reference import (only where /root/reference exists)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_ref, ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _check(module, z, train, tol=1e-11):
    module = module.double()
    module.train(train)
    x = torch.from_numpy(z["x"]).double().requires_grad_(True)
    out = module(x)
    out.backward(torch.from_numpy(z["go"]).double())
    assert np.abs(out.detach().numpy() - z["out"]).max() < tol * max(1.0, np.abs(z["out"]).max())
    assert np.abs(x.grad.numpy() - z["gx"]).max() < tol * max(1.0, np.abs(z["gx"]).max())
    for k, p in module.named_parameters():
        if p.grad is None:
            continue
        want = z["grad/" + k]
        assert np.abs(p.grad.numpy() - want).max() < tol * max(1.0, np.abs(want).max()), k
    if train:
        for k, b in module.named_buffers():
            assert np.abs(b.numpy() - z["buf/" + k]).max() < 1e-10, k


@pytest.mark.parametrize("tag", ["gcn_64_64_v25", "gcn_64_128_v25", "gcn_128_128_v33"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_shift_gcn_against_committed_reference_output(tag, mode):
    z = _load(f"{tag}_{mode}.npz")
    C, D, V = (int(v) for v in z["meta"][:3])
    m = model_ref.fill_module_(model_ref.RefShiftGcn(C, D, None, num_point=V))
    _check(m, z, mode == "train")


@pytest.mark.parametrize("tag", ["unit_64_64_s1", "unit_64_128_s2"])
def test_unit_against_committed_reference_output(tag):
    z = _load(f"{tag}_train.npz")
    C, D, V, n, T, s, res = (int(v) for v in z["meta"])
    m = model_ref.fill_module_(model_ref.RefUnit(C, D, None, stride=s, residual=bool(res), num_point=V))
    _check(m, z, True)


def test_model_against_committed_reference_logits():
    z = _load("model_ntu60_eval.npz")
    m = model_ref.fill_module_(model_ref.RefModel(60, 25, 2)).double()
    sd = m.state_dict()
    for k in z.files:
        if k.startswith("buf/"):
            sd[k[4:]] = torch.from_numpy(z[k]).double()
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        out = m(torch.from_numpy(z["x"]).double())
    assert np.abs(out.numpy() - z["logits"]).max() < 1e-9 * np.abs(z["logits"]).max()


def test_index_tables_bit_exact():
    z = _load("tables.npz")
    from shiftgcn_b200.modules import shift_tables
    for key in z.files:
        kind, V, C, D = key.split("_")
        V, C, D = int(V), int(C), int(D)
        closed = model_ref.shift_tables_closed_form(V, C, D)
        loop = model_ref.shift_tables_loop(V, C, D)
        prod = shift_tables(V, C, D)
        i = 0 if kind == "in" else 1
        for name, tab in (("closed form", closed[i]), ("loop", loop[i]), ("product", prod[i])):
            assert tab.dtype == np.int64 and np.array_equal(tab, z[key]), f"{key}: {name}"
        assert np.array_equal(np.sort(z[key]), np.arange(z[key].size))     # both tables are permutations


@pytest.mark.parametrize("tag,kw", [("ntu60", dict(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph")),
                                    ("mediapipe", dict(num_class=2, num_point=33, num_person=1, graph="graph.mediapipe_pose.Graph"))])
def test_state_dict_contract(tag, kw):
    """keys, shapes and dtypes of the drop-in Model equal the reference's"""
    from shiftgcn_b200.modules import Model
    with open(os.path.join(GOLD, "state_dict_contract.json")) as f:
        contract = json.load(f)[tag]
    sd = Model(graph_args=dict(labeling_mode="spatial"), **kw).state_dict()
    got = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
    assert sorted(got) == sorted(contract)
    oracle_sd = model_ref.RefModel(kw["num_class"], kw["num_point"], kw["num_person"]).state_dict()
    assert sorted(oracle_sd) == sorted(sd)


def test_tf32_rounding_helper():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -(1.0 + 3 * 2 ** -12), 3.14159265], dtype=torch.float32)
    r = model_ref.tf32_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[2] == 1.0 + 2 ** -10     # ties away from zero
    assert r[3] == -(1.0 + 2 ** -10 * 1.0) or r[3] == -(1.0 + 2 ** -9)
    assert (r.view(torch.int32) & 0x1FFF).abs().max() == 0


# ------------------------------------------------------------------ live reference (build container only)
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")


@needs_ref
def test_oracle_equals_live_reference_unit():
    ns = ref_import.load()
    torch.manual_seed(1)
    for (C, D, s, res) in [(64, 64, 1, True), (16, 32, 2, True), (3, 16, 1, False)]:
        u = model_ref.fill_module_(ref_import.build(ns.TCN_GCN_unit, C, D, None, stride=s, residual=res, num_point=25)).double()
        m = model_ref.fill_module_(model_ref.RefUnit(C, D, None, stride=s, residual=res, num_point=25)).double()
        x = torch.randn(2, C, 12, 25, dtype=torch.float64, requires_grad=True)
        x2 = x.detach().clone().requires_grad_(True)
        a, b = u(x), m(x2)
        assert torch.equal(a, b)
        a.square().sum().backward(), b.square().sum().backward()
        assert torch.equal(x.grad, x2.grad)
        for (k, p), (_, q) in zip(u.named_parameters(), m.named_parameters()):
            if p.grad is not None:
                assert torch.equal(p.grad, q.grad), k


@needs_ref
def test_graph_adjacency_equals_live_reference():
    ns = ref_import.load()
    import graph.mediapipe_pose as mp
    import graph.ntu_rgb_d as ntu
    for mine, ref in ((ntu, ns.graph_ntu), (mp, ns.graph_mediapipe)):
        a, b = mine.Graph("spatial"), ref.Graph("spatial")
        assert np.array_equal(a.A, b.A) and a.num_node == b.num_node and sorted(a.inward) == sorted(b.inward)


def test_output_window_choice_follows_the_shift_positions():
    """host logic of Shift_tcn.out_window_ok (which inference kernel serves the unit): re-derived after an in-place
    update of ypos (optimizer step), after a raw .data edit + ops.params_changed(), and after load_state_dict"""
    from shiftgcn_b200 import ops
    from shiftgcn_b200.modules import Shift_tcn
    torch.manual_seed(3)
    m = Shift_tcn(64, 64)
    assert m.out_window_ok()                              # U(-1, 1): two floor values
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        m.shift_out.ypos[::4] += 2.7                      # four floor values
    assert not m.out_window_ok()
    m.load_state_dict(sd)
    assert m.out_window_ok()
    m.shift_out.ypos.data[0] = 5.3                        # raw edit: the version counter of the parameter does not move
    ops.params_changed()
    assert not m.out_window_ok()
    with torch.no_grad():
        m.shift_out.ypos[0] = 0.3
        m.shift_out.ypos[1] = 9.5                         # outside the clamp range of the kernel's window base
    assert not m.out_window_ok()
