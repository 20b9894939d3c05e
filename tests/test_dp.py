"""Flat-buffer data-parallel SGD (shiftgcn_b200/dp.py) on CPU: world_size 2 over gloo.

Checks the host-side logic of the multi-GPU path without a GPU: one all-reduce of the flat gradient buffer, the K5 sign
constraint applied AFTER the reduction on the reduced raw sums, and SGD-Nesterov with the reference's
per-parameter weight decay (main.py:307-322) -- two ranks with half a batch each must follow exactly the trajectory of
one process that sees the whole batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from shiftgcn_b200.dp import FlatSGDTrainer, reference_weight_decay


class FakeShift(nn.Module):
    """stands in for cuda.shift.Shift: a sign-constrained ypos gradient whose raw sum is exported on the module"""

    def __init__(self, c):
        super().__init__()
        self.xpos = nn.Parameter(torch.zeros(c))
        self.ypos = nn.Parameter(torch.linspace(-1, 1, c))

    def forward(self, x):
        ypos, xpos, mod = self.ypos, self.xpos, self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, xp, yp):
                ctx.save_for_backward(x)
                return x * (1 + yp)

            @staticmethod
            def backward(ctx, g):
                (x,) = ctx.saved_tensors
                raw = (g * x).mean(0)                                     # mean over the (local) batch, like K4 + at::mean
                if getattr(mod, "_export_raw", False):
                    mod._raw_ypos_grad = raw
                gy = torch.where(raw != 0, torch.sign(raw) * 0.01, torch.full_like(raw, 0.0001))
                return g * (1 + ypos.detach()), torch.zeros_like(xpos), gy

        return Fn.apply(x, xpos, ypos)


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.Linear_weight = nn.Parameter(torch.randn(6, 5) * 0.3)
        self.Feature_Mask = nn.Parameter(torch.randn(5) * 0.1)
        self.shift = FakeShift(5)
        self.fc = nn.Linear(5, 3)

    def forward(self, x):
        h = x @ self.Linear_weight * (torch.tanh(self.Feature_Mask) + 1)
        return self.fc(self.shift(torch.relu(h)))


def _data():
    g = torch.Generator().manual_seed(7)
    return torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)


def _trajectory(trainer, x, y, steps=3):
    out = []
    for _ in range(steps):
        trainer.train_step(x, y)
        out.append(trainer.flat_param.clone())
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    net = Net()
    trainer = FlatSGDTrainer(net, lr=0.1, momentum=0.9, nesterov=True)
    x, y = _data()
    shard = slice(rank * 4, rank * 4 + 4)
    traj = _trajectory(trainer, x[shard], y[shard])
    if rank == 0:
        q.put([t.numpy() for t in traj])
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_match_one_process_on_the_whole_batch():
    torch.manual_seed(3)
    net = Net()
    single = FlatSGDTrainer(net, lr=0.1, momentum=0.9, nesterov=True)
    x, y = _data()
    want = _trajectory(single, x, y)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step, (a, b) in enumerate(zip(want, got)):
        assert torch.allclose(a, torch.from_numpy(b), rtol=1e-5, atol=1e-6), f"step {step}"


def test_flat_views_and_reference_sgd_semantics():
    """p.data / p.grad are views of the flat buffers; the update equals torch.optim.SGD with the reference's groups"""
    torch.manual_seed(5)
    net, ref = Net(), Net()
    ref.load_state_dict(net.state_dict())
    trainer = FlatSGDTrainer(net, lr=0.05, momentum=0.9, nesterov=True)
    for _, _, _, shift in trainer.ypos_slices:
        shift._export_raw = False                     # plain SGD comparison: use the in-graph K5 values on both sides
    groups = [{"params": [p], "weight_decay": reference_weight_decay(n)} for n, p in ref.named_parameters()]
    opt = torch.optim.SGD(groups, lr=0.05, momentum=0.9, nesterov=True)
    x, y = _data()
    for _ in range(3):
        trainer.train_step(x, y)
        opt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        opt.step()
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert p.data.untyped_storage().data_ptr() == trainer.flat_param.untyped_storage().data_ptr(), n
        assert p.grad.untyped_storage().data_ptr() == trainer.flat_grad.untyped_storage().data_ptr(), n
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), n
    assert reference_weight_decay("l1.gcn1.Linear_weight") == 1e-3 and reference_weight_decay("l1.gcn1.Feature_Mask") == 0.0
    assert reference_weight_decay("l1.tcn1.bn.weight") == 1e-4


def test_host_prefetcher_order_and_errors():
    """HostPrefetcher (CPU path: no streams): batches come back in put() order, slots are reused, get() needs a put()"""
    import pytest
    from shiftgcn_b200.dp import HostPrefetcher
    pf = HostPrefetcher("cpu")
    batches = [(torch.full((3,), float(i)), torch.tensor([i])) for i in range(5)]
    pf.put(*batches[0])
    for i in range(5):
        x, y = pf.get()
        if i + 1 < 5:
            pf.put(*batches[i + 1])
        assert torch.equal(x, batches[i][0]) and torch.equal(y, batches[i][1])
    with pytest.raises(RuntimeError):
        pf.get()
