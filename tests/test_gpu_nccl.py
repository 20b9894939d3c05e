"""Data-parallel step over NCCL: two ranks, one GPU each, against ONE process that runs both
shards itself.  BatchNorm statistics are local to a rank (the reference's DataParallel semantics, main.py:294-299), so
the reduced gradient of the 2-rank step must equal the mean of the two per-shard gradients, the K5 constraint acts on
the reduced raw position sums, and both ranks end the step with identical parameters.  Needs two visible GPUs
(skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py`)."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(device):
    from oracle import model_ref
    from shiftgcn_b200.modules import Model
    torch.manual_seed(1)
    mod = Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph", graph_args=dict(labeling_mode="spatial"))
    model_ref.fill_module_(mod)
    return mod.to(device).train()


def _batch():
    g = torch.Generator().manual_seed(31)
    return torch.randn(4, 3, 32, 25, 2, generator=g), torch.randint(0, 60, (4,), generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from shiftgcn_b200.dp import FlatSGDTrainer
    x, y = _batch()
    tr = FlatSGDTrainer(_model(dev), lr=0.05)
    lo, hi = rank * 2, rank * 2 + 2
    loss = tr.train_step(x[lo:hi].to(dev), y[lo:hi].to(dev))
    torch.cuda.synchronize()
    out[rank] = dict(grad=tr.flat_grad.cpu(), param=tr.flat_param.cpu(), loss=float(loss))
    dist.barrier()
    dist.destroy_process_group()


def test_two_nccl_ranks_equal_one_process_running_both_shards(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from shiftgcn_b200.dp import FlatSGDTrainer
    from util import rel_l2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    r0, r1 = out[0], out[1]
    assert torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["param"], r1["param"])   # one all-reduce, same update
    # one process: each shard through its own copy of the model (local BatchNorm statistics), gradients averaged
    x, y = _batch()
    base = _model(cuda_device)
    shards = []
    for lo in (0, 2):
        tr = FlatSGDTrainer(copy.deepcopy(base), lr=0.05)
        tr.zero_grad()
        loss = torch.nn.functional.cross_entropy(tr.model(x[lo:lo + 2].to(cuda_device)), y[lo:lo + 2].to(cuda_device))
        loss.backward()
        tr.gather_gradients()
        tr.reduce_gradients()                            # world 1: collects the raw position sums only
        shards.append((tr, float(loss)))
    want = (shards[0][0].flat_grad + shards[1][0].flat_grad) / 2
    n = shards[0][0].n_param
    pos = shards[0][0].ypos_src >= 0
    got = r0["grad"].to(cuda_device)
    assert abs(r0["loss"] - shards[0][1]) < 2e-3 * abs(shards[0][1]) and abs(r1["loss"] - shards[1][1]) < 2e-3 * abs(shards[1][1])
    # raw position sums (behind the parameters) and all non-position gradients: the NCCL average of the two shards
    assert rel_l2(got[n:], want[n:]) < 2e-2
    assert rel_l2(got[:n][~pos], want[:n][~pos]) < 2e-2
    # position gradients after the step: K5 of the REDUCED raw sums (+-0.01 by sign), not the sum of per-rank +-0.01
    raw = want[n:]
    k5 = torch.where(raw != 0, torch.sign(raw) * 0.01, torch.full_like(raw, 1e-4))
    sure = raw.abs() > 1e-3 * raw.abs().max()
    src = shards[0][0].ypos_src[pos].long()
    assert torch.equal(got[:n][pos][sure[src]], k5[src][sure[src]])
