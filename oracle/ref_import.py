"""Import the REAL reference modules on CPU (TEST INFRASTRUCTURE ONLY; build container only).

/root/reference does not exist on the GPU box, so nothing that runs there may import this file;
it is used by oracle/make_golden.py (to write tests/golden/*.npz) and by the ``not gpu`` tests that
pin oracle/model_ref.py against the reference (they skip when the tree is absent).

Shims:
  1. the reference hard-codes ``device='cuda'`` when it builds parameters
     (model/shift_gcn.py:90,93,96) -> torch.zeros/ones are wrapped to drop the kwarg while the
     reference code constructs modules;
  2. ``from cuda.shift import Shift`` (model/shift_gcn.py:11) needs the compiled extension -> a
     ``cuda.shift`` module backed by oracle/shift_torch.py is registered in sys.modules first;
  3. the reference's ``model``/``graph`` top-level names collide with this repo's drop-in packages
     -> reference files are loaded under ``ref_*`` aliases.
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

from . import shift_torch

REFERENCE_ROOT = os.environ.get("SHIFTGCN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "shift_gcn.py"))


@contextlib.contextmanager
def cpu_construction():
    """Drop ``device=...`` from torch.zeros / torch.ones while reference code builds parameters."""
    real_zeros, real_ones = torch.zeros, torch.ones

    def zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    def ones(*a, **k):
        k.pop("device", None)
        return real_ones(*a, **k)

    torch.zeros, torch.ones = zeros, ones
    try:
        yield
    finally:
        torch.zeros, torch.ones = real_zeros, real_ones


def _load_file(alias, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """Returns a namespace with the reference's Shift_gcn, Shift_tcn, TCN_GCN_unit, tcn, Model and graphs."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")

    # (2) stand-in for the compiled op; the class bodies restate cuda/shift.py, the math is oracle/shift_torch.py
    stub = types.ModuleType("cuda.shift")
    stub.Shift = shift_torch.OracleShift
    stub.ShiftFunction = shift_torch.OracleShiftFunction
    saved_stub = sys.modules.get("cuda.shift")
    sys.modules["cuda.shift"] = stub

    # (3) reference graph modules do ``from graph import tools``: load them with the reference's
    # own ``graph`` package temporarily bound to that name.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "graph" or k.startswith("graph.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import graph.ntu_rgb_d as ref_ntu            # noqa: E402  (reference's, via sys.path[0])
        import graph.mediapipe_pose as ref_mp        # noqa: E402
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "graph" or k.startswith("graph.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    sys.modules["ref_graph_ntu_rgb_d"] = ref_ntu
    sys.modules["ref_graph_mediapipe_pose"] = ref_mp

    with cpu_construction():
        ref_model = _load_file("ref_shift_gcn", os.path.join("model", "shift_gcn.py"))

    if saved_stub is not None:
        sys.modules["cuda.shift"] = saved_stub
    else:
        sys.modules.pop("cuda.shift", None)

    ns = types.SimpleNamespace(
        module=ref_model,
        Shift_gcn=ref_model.Shift_gcn,
        Shift_tcn=ref_model.Shift_tcn,
        TCN_GCN_unit=ref_model.TCN_GCN_unit,
        tcn=ref_model.tcn,
        Model=ref_model.Model,
        graph_ntu=ref_ntu,
        graph_mediapipe=ref_mp,
        GRAPH_NTU="ref_graph_ntu_rgb_d.Graph",
        GRAPH_MEDIAPIPE="ref_graph_mediapipe_pose.Graph",
    )
    _cache["ns"] = ns
    return ns


def build(factory, *args, **kwargs):
    """Construct a reference module on CPU: ``build(ns.Shift_gcn, 64, 64, A)``."""
    with cpu_construction():
        return factory(*args, **kwargs)
