"""4-stream ensemble: the oracle of the stream derivation against the reference-made golden
vectors, the host-side placement logic, and the sharded logits reduction over gloo (world 2 and 4, CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import modalities
from shiftgcn_b200 import ensemble as E

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modalities.npz")


@pytest.mark.parametrize("tag", ["mp", "ntu"])
def test_oracle_streams_equal_reference_golden(tag):
    g = np.load(GOLDEN)
    got = modalities.derive(g[f"{tag}/joint"])
    for name in modalities.MODALITIES:
        assert np.array_equal(got[name], g[f"{tag}/{name}"]), name      # fp32 subtractions: bit-exact


def test_bone_tables_match_the_oracle_tables():
    for V in (25, 33):
        parents = E.bone_parents(V)
        assert len(parents) == V
        assert dict(modalities.pairs_0based(V)) == {v: p for v, p in enumerate(parents)}
    assert E.MODALITIES == modalities.MODALITIES and E.ENSEMBLE_WEIGHTS_DEFAULT == modalities.ALPHA
    with pytest.raises(ValueError):
        E.bone_parents(18)


def test_placement_covers_every_stream_and_sample_once():
    for world in (1, 2, 4, 8, 16):
        for n in (1, 7, 64):
            seen = np.zeros((4, n), dtype=int)
            for rank in range(world):
                for k, shard, shards in E.placement(world, rank):
                    lo, hi = E.shard_bounds(n, shard, shards)
                    seen[k, lo:hi] += 1
            assert (seen == 1).all(), (world, n)
    for bad in (3, 5, 6):
        with pytest.raises(ValueError):
            E.placement(bad, 0)


def _models(num_class, feat):
    """four small deterministic stand-in models: stream batch (n, C, T, V, M) -> logits (n, num_class)"""
    out = {}
    for k, name in enumerate(E.MODALITIES):
        g = torch.Generator().manual_seed(100 + k)
        W = torch.randn(feat, num_class, generator=g) / feat ** 0.5
        out[name] = (lambda W: (lambda x: torch.tanh(x.reshape(x.shape[0], -1)) @ W))(W)
    return out


def _cpu_stream(joint, name):
    return torch.from_numpy(modalities.derive(joint.numpy())[name])


def _joint():
    g = torch.Generator().manual_seed(11)
    return torch.randn(6, 3, 5, 25, 2, generator=g)


def _expected(joint, num_class):
    models = _models(num_class, joint[0].numel())
    streams = modalities.derive(joint.numpy())
    logits = [models[name](torch.from_numpy(streams[name])).numpy() for name in E.MODALITIES]
    return modalities.ensemble_logits(logits)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    joint = _joint()
    all_models = _models(4, joint[0].numel())
    mine = {E.MODALITIES[k]: all_models[E.MODALITIES[k]] for k, _, _ in E.placement(world, rank)}
    ens = E.StreamEnsemble(mine, num_class=4, world_size=world, rank=rank, stream_fn=_cpu_stream)
    out = ens.logits(joint)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ensemble_equals_the_reference_sum(world):
    want = _expected(_joint(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):                                     # every rank ends up with the full ensemble logits
        assert np.allclose(got[rank], want, rtol=1e-5, atol=1e-6), rank


def test_single_process_ensemble_and_fall_score():
    joint = _joint()
    ens = E.StreamEnsemble(_models(4, joint[0].numel()), num_class=4, stream_fn=_cpu_stream)
    want = _expected(joint, 4)
    assert np.allclose(ens.logits(joint).numpy(), want, rtol=1e-5, atol=1e-6)
    assert np.allclose(ens.scores(joint)[:, 1].numpy(), modalities.fall_scores(want), rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        E.StreamEnsemble({"joint": None}, num_class=4)           # a single process owns all four streams
