"""Feeder -> GPU pipeline (the caller side of the hot path in training: main.py:231-251 builds the DataLoader, :400-402
moves every batch with a synchronous ``data.float().cuda()`` right in front of the forward pass; the augmentation
``random_move`` runs per sample in numpy inside the Dataset, feeders/feeder.py:70-87, feeders/tools.py:58-101).

``DeviceFeeder`` wraps any iterable of ``(data, label, index)`` batches (a ``torch.utils.data.DataLoader`` over the
reference's ``Feeder``, or plain tensors):
  * batches are staged through pinned host buffers and copied on a side stream ``depth`` batches ahead, so the H2D copy of
    batch i+1 overlaps the step on batch i; the consumer's stream only waits on the copy event of its own batch;
  * ``data.float()`` / ``label.long()`` happen on the device;
  * ``random_move=True`` applies the reference's augmentation to the whole batch on the device (``sgcn_random_move``);
    the random node values are drawn on the host with the reference's own ``np.random.choice`` calls, sample by sample,
    so a seeded run sees the same augmentation (switch the Dataset's own ``random_move`` off).
"""
import numpy as np
import torch

from . import ops

ANGLE_CANDIDATE = [-10., -5., 0., 5., 10.]
SCALE_CANDIDATE = [0.9, 1.0, 1.1]
TRANSFORM_CANDIDATE = [-0.2, -0.1, 0.0, 0.1, 0.2]


def draw_move_nodes(n_samples, T, move_time=1):
    """node [K+1] int32 and vals [n, 4, K+1] float64 of feeders/tools.py:65-73, one sample after the other (each sample
    consumes np.random exactly like one call of the reference's ``random_move``)."""
    node = np.append(np.arange(0, T, T * 1.0 / move_time).round().astype(int), T)
    num_node = len(node)
    vals = np.empty((n_samples, 4, num_node), dtype=np.float64)
    for i in range(n_samples):
        vals[i, 0] = np.random.choice(ANGLE_CANDIDATE, num_node)
        vals[i, 1] = np.random.choice(SCALE_CANDIDATE, num_node)
        vals[i, 2] = np.random.choice(TRANSFORM_CANDIDATE, num_node)
        vals[i, 3] = np.random.choice(TRANSFORM_CANDIDATE, num_node)
    return node.astype(np.int32), vals


class DeviceFeeder:
    """for data, label, index in DeviceFeeder(loader, device): ...   (data fp32 and label int64 on ``device``)"""

    def __init__(self, loader, device, random_move=False, move_time=1, depth=2):
        if depth < 1:
            raise ValueError("depth must be at least 1")
        self.loader, self.device = loader, torch.device(device)
        self.random_move, self.move_time, self.depth = bool(random_move), int(move_time), int(depth)
        self.cuda = self.device.type == "cuda"
        if self.random_move and not self.cuda:
            raise RuntimeError("DeviceFeeder(random_move=True) runs the augmentation kernel and needs a CUDA device")
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._pinned = {}                                         # slot -> pinned host staging tensors
        self.batches = 0

    def __len__(self):
        return len(self.loader)

    def _pin(self, slot, k, t):
        """pinned staging copy of host tensor t (reused per slot; a DataLoader with pin_memory=True passes through)"""
        if not self.cuda or t.is_pinned():
            return t
        buf = self._pinned.get((slot, k))
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[(slot, k)] = buf
        buf.copy_(t)
        return buf

    def _stage(self, slot, batch):
        data, label = torch.as_tensor(batch[0]), torch.as_tensor(batch[1])
        index = batch[2] if len(batch) > 2 else None
        if not self.cuda:
            return data.float(), label.long(), index, None
        moves = draw_move_nodes(data.shape[0], data.shape[2], self.move_time) if self.random_move else None
        self.stream.wait_stream(torch.cuda.current_stream(self.device))    # the slot's previous consumer is enqueued
        with torch.cuda.stream(self.stream):
            d = self._pin(slot, 0, data).to(self.device, non_blocking=True).float().contiguous()
            l = self._pin(slot, 1, label).to(self.device, non_blocking=True).long()
            if moves is not None:
                node = torch.from_numpy(moves[0]).to(self.device, non_blocking=True)
                vals = torch.from_numpy(moves[1]).to(self.device, non_blocking=True)
                ops.random_move_(d, vals, node)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return d, l, index, ev

    def __iter__(self):
        it = iter(self.loader)
        queue, slot = [], 0
        while True:
            while it is not None and len(queue) < self.depth:
                try:
                    batch = next(it)
                except StopIteration:
                    it = None
                    break
                queue.append(self._stage(slot % (self.depth + 1), batch))
                slot += 1
            if not queue:
                return
            d, l, index, ev = queue.pop(0)
            if ev is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                d.record_stream(cur), l.record_stream(cur)
            self.batches += 1
            yield d, l, index
