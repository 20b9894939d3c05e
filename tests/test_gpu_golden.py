"""The CUDA modules against the committed golden fixtures DIRECTLY (tests/golden/*.npz were produced by executing the
real reference on CPU in fp64, oracle/make_golden.py): the fixture's input and upstream gradient go through the
sm_100a kernels, outputs / input gradients / parameter gradients / BatchNorm buffers are compared with the stored
reference results.  Tolerances: the north star's TF32 bar (1e-2 relative; gradients L2-relative 5e-2, see util.rel_l2)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref
from util import check_ypos_grad, rel_err, rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL, TOL_GRAD = 1e-2, 5e-2


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _check(mod, z, train, device):
    mod = mod.to(device)
    mod.train(train)
    x = torch.from_numpy(z["x"]).to(device).requires_grad_(True)
    go = torch.from_numpy(z["go"]).to(device)
    out = mod(x)
    out.backward(go)
    torch.cuda.synchronize()
    assert rel_err(out, z["out"]) < TOL
    assert rel_l2(x.grad, z["gx"]) < TOL_GRAD
    floor = 1e-6 * max(float(np.abs(z["go"]).sum()), 1.0)
    seen = 0
    for name, p in mod.named_parameters():
        key = "grad/" + name
        if not p.requires_grad or key not in z.files:
            continue
        seen += 1
        want = torch.from_numpy(z[key])
        if name.endswith("ypos"):
            raw = torch.from_numpy(z["raw/" + name]) if "raw/" + name in z.files else None
            check_ypos_grad(name, p.grad, want, raw, decisive=5e-2, min_sure=0.3)   # exact oracle vs TF32 sums
        elif name.endswith("xpos"):
            assert torch.count_nonzero(p.grad).item() == 0
        else:
            err = (p.grad.double().cpu() - want.double()).norm().item()
            assert err < TOL_GRAD * want.double().norm().item() + floor, f"grad {name}: l2 err {err:.3e}"
    assert seen >= 5
    if train:
        for name, b in mod.named_buffers():
            if b.dtype.is_floating_point:
                assert rel_err(b, z["buf/" + name]) < TOL, f"buffer {name}"
            else:
                assert np.array_equal(b.cpu().numpy(), z["buf/" + name]), f"buffer {name}"


@pytest.mark.parametrize("tag", ["gcn_64_64_v25", "gcn_64_128_v25", "gcn_128_128_v33"])
@pytest.mark.parametrize("train", [True, False])
def test_shift_gcn_golden(cuda_device, tag, train):
    from shiftgcn_b200.modules import Shift_gcn
    z = _load(f"{tag}_{'train' if train else 'eval'}.npz")
    C, D, V, n, T, tr = (int(v) for v in z["meta"])
    assert bool(tr) == train
    mod = model_ref.fill_module_(Shift_gcn(C, D, None, num_point=V))
    _check(mod, z, train, cuda_device)


@pytest.mark.parametrize("tag", ["unit_64_64_s1", "unit_64_128_s2"])
def test_tcn_gcn_unit_golden(cuda_device, tag):
    from shiftgcn_b200.modules import TCN_GCN_unit
    z = _load(f"{tag}_train.npz")
    C, D, V, n, T, s, res = (int(v) for v in z["meta"])
    mod = model_ref.fill_module_(TCN_GCN_unit(C, D, None, stride=s, residual=bool(res), num_point=V))
    _check(mod, z, True, cuda_device)


def test_model_ntu60_golden_logits(cuda_device):
    """full NTU-60 model in inference on the fixture's input with the fixture's BatchNorm buffers"""
    from shiftgcn_b200.modules import Model
    z = _load("model_ntu60_eval.npz")
    mod = model_ref.fill_module_(Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph",
                                       graph_args=dict(labeling_mode="spatial")))
    sd = mod.state_dict()
    for k in z.files:
        if k.startswith("buf/"):
            sd[k[4:]] = torch.from_numpy(z[k])
    mod.load_state_dict(sd)
    mod = mod.to(cuda_device).eval()
    with torch.no_grad():
        out = mod(torch.from_numpy(z["x"]).to(cuda_device))
    want = torch.from_numpy(z["logits"])
    assert rel_err(out, want) < TOL
    assert torch.equal(out.argmax(1).cpu(), want.argmax(1))
