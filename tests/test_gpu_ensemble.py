"""Input streams on the device (sgcn_input_stream) and the 4-stream ensemble through the real models (config 5)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref, modalities
from util import fill_pair, rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modalities.npz")


@pytest.mark.parametrize("tag", ["mp", "ntu"])
def test_streams_bit_exact_against_reference_golden(cuda_device, tag):
    from shiftgcn_b200 import ensemble as E
    g = np.load(GOLDEN)
    joint = torch.from_numpy(g[f"{tag}/joint"]).to(cuda_device)
    got = E.derive_modalities(joint)
    for name in E.MODALITIES:
        assert np.array_equal(got[name].cpu().numpy(), g[f"{tag}/{name}"]), name


@pytest.mark.parametrize("shape", [(5, 3, 300, 25, 2), (3, 3, 64, 33, 1), (1, 3, 1, 25, 2), (2, 3, 2, 33, 1)])
def test_streams_bit_exact_against_oracle(cuda_device, shape):
    from shiftgcn_b200 import ensemble as E
    rng = np.random.default_rng(shape[2])
    joint = rng.standard_normal(shape).astype(np.float32)
    want = modalities.derive(joint)
    got = E.derive_modalities(torch.from_numpy(joint).to(cuda_device))
    for name in E.MODALITIES:
        assert np.array_equal(got[name].cpu().numpy(), want[name]), name


def test_rows_layout_with_folded_input_bn(cuda_device):
    """rows != 0: (N*M, T, V, C) channels-last rows with data_bn in inference form (model/shift_gcn.py:193-198)"""
    from shiftgcn_b200 import ensemble as E, ops
    N, C, T, V, M = 3, 3, 10, 25, 2
    rng = np.random.default_rng(3)
    joint = rng.standard_normal((N, C, T, V, M)).astype(np.float32)
    bn = torch.nn.BatchNorm1d(M * V * C).double().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(0, 0.1), bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.5, 1.5)
    for name in E.MODALITIES:
        stream = torch.from_numpy(modalities.derive(joint)[name]).double()
        x = stream.permute(0, 4, 3, 1, 2).contiguous().view(N, M * V * C, T)          # reference :195-198
        want = bn(x).view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
        bone, motion = E.stream_flags(name)
        par = torch.tensor(E.bone_parents(V), dtype=torch.int32, device=cuda_device) if bone else None
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach()
        shift = (bn.bias - bn.running_mean * scale).detach()
        rows = ops.input_stream(torch.from_numpy(joint).to(cuda_device), parent=par, motion=motion, rows=True,
                                scale=scale.float().to(cuda_device), shift=shift.float().to(cuda_device))
        assert rows.shape == (N * M, T, V, C)
        assert rel_err(rows.permute(0, 3, 1, 2), want) < 1e-6, name


def test_input_stream_argument_errors(cuda_device):
    from shiftgcn_b200 import ops
    x = torch.zeros(2, 3, 4, 25, 2, device=cuda_device)
    with pytest.raises(RuntimeError):
        ops.input_stream(x[..., 0])                                             # not a 5-D joint batch
    with pytest.raises(RuntimeError):
        ops.input_stream(x, parent=torch.zeros(24, dtype=torch.int32, device=cuda_device))
    with pytest.raises(RuntimeError):
        ops.input_stream(x, rows=False, scale=torch.ones(150, device=cuda_device), shift=torch.ones(150, device=cuda_device))
    assert ops.input_stream(x[:0].contiguous()).numel() == 0                     # empty batch


@pytest.mark.parametrize("num_class,V,M", [(2, 33, 1), (60, 25, 2)])
def test_four_stream_ensemble_matches_oracle(cuda_device, num_class, V, M):
    from shiftgcn_b200 import ensemble as E
    from shiftgcn_b200.modules import Model
    graph = "graph.ntu_rgb_d.Graph" if V == 25 else "graph.mediapipe_pose.Graph"
    g = torch.Generator().manual_seed(9)
    joint = torch.randn(4, 3, 32, V, M, generator=g)
    streams = modalities.derive(joint.numpy())
    models, want = {}, []
    for k, name in enumerate(E.MODALITIES):
        mod = Model(num_class=num_class, num_point=V, num_person=M, graph=graph, graph_args=dict(labeling_mode="spatial"))
        ref = model_ref.RefModel(num_class=num_class, num_point=V, num_person=M)
        fill_pair(mod, ref, prefix=f"s{k}.")
        x = torch.from_numpy(streams[name]).double()
        ref = ref.double().train()
        for m in ref.modules():                                    # realistic running statistics (one momentum-1 pass)
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 1.0
        with torch.no_grad():
            ref(x)
        ref.eval()
        mod.load_state_dict({k2: v.float() if v.dtype.is_floating_point else v for k2, v in ref.state_dict().items()})
        mod = mod.to(cuda_device).eval()
        models[name] = (lambda m, n: (lambda jb: m.forward_stream(jb, n)))(mod, name)
        with torch.no_grad():
            want.append(ref(x).numpy())
    want = modalities.ensemble_logits(want)
    ens = E.StreamEnsemble(models, num_class=num_class, stream_fn=lambda jb, name: jb)   # models derive their own stream
    got = ens.logits(joint.to(cuda_device)).cpu().numpy()
    assert rel_err(got, want) < (1e-2 if num_class > 2 else 2.5e-2)
    assert np.array_equal(got.argmax(1), want.argmax(1))
    scores = ens.scores(joint.to(cuda_device)).cpu().numpy()
    assert np.allclose(scores.sum(1), 1.0, atol=1e-5)
