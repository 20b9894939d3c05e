"""Data-parallel training of the Shift-GCN path: one process per GPU, ONE NCCL all-reduce per step.

Replaces the reference's ``nn.DataParallel`` wrapper (main.py:294-299: per-iteration scatter / replicate / gather)
and its per-parameter SGD groups (main.py:301-322):

  * parameters and gradients live in two flat fp32 buffers (``p.data`` / ``p.grad`` are views), so the gradient
    exchange is a single ``all_reduce(AVG)`` of ~2.8 MB over NVLink instead of ~300 small ones;
  * BatchNorm statistics stay local to each GPU, exactly like DataParallel (no SyncBN in the reference);
  * the shift-position gradient is a SIGN (kernel K5, shift_cuda_kernel.cu:371-395).  The raw per-channel sums
    travel in the same flat buffer and the constraint is applied AFTER the reduction, which gives the
    single-GPU large-batch semantics (SURVEY.md App. E-7 documents how DataParallel differs: it sums the
    already-constrained per-replica +-0.01 values);
  * SGD with momentum / Nesterov and the reference's decay rules: 1e-4 everywhere (BN and biases included),
    1e-3 for ``Linear_weight``, 0 for ``Feature_Mask`` (main.py:307-317): scale, K5, decay and the momentum update
    are ONE kernel over the flat buffers (``sgcn_sgd_epilogue``);
  * the learning rate lives in device memory (``set_lr`` / ``reference_lr`` = main.py:342-351), so the captured CUDA
    graph of a whole step follows the warm-up / step-decay schedule without re-capture.

The batch is sharded by the caller (each rank feeds its own samples); there is no collective on the data path.
"""
import torch
import torch.distributed as dist

from . import ops


def reference_lr(epoch, base_lr=0.1, warm_up_epoch=0, step=(60, 80, 100)):
    """``Processor.adjust_learning_rate`` (main.py:342-351): linear warm-up, then x0.1 at every epoch in ``step``"""
    if epoch < warm_up_epoch:
        return base_lr * (epoch + 1) / warm_up_epoch
    return base_lr * (0.1 ** sum(1 for s in step if epoch >= s))


def reference_weight_decay(name):
    """main.py:307-317"""
    if "Linear_weight" in name:
        return 1e-3
    if "Mask" in name:
        return 0.0
    return 1e-4


class FlatSGDTrainer:
    """Flat-buffer data-parallel SGD around a ``Model`` (or any module using this package's Shift modules)."""

    def __init__(self, model, lr=0.1, momentum=0.9, nesterov=True, process_group=None, weight_decay_fn=None):
        self.model = model
        self.lr, self.momentum, self.nesterov = lr, momentum, nesterov
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        wd_fn = weight_decay_fn or reference_weight_decay
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        # every parameter starts on a 256-byte boundary inside the flat buffers (library kernels such as cuDNN's
        # convolutions and our own vectorised loads assume at least 16-byte aligned tensors)
        ALIGN = 64
        offsets, off = [], 0
        for sz in sizes:
            offsets.append(off)
            off += (sz + ALIGN - 1) // ALIGN * ALIGN
        self.n_param = off
        # raw shift-position sums ride behind the gradients: one slot per ypos element
        self.ypos_slices = []                       # (grad_offset, raw_offset, count, owning Shift module)
        shifts = {id(m.ypos): m for m in model.modules() if hasattr(m, "ypos") and hasattr(m, "xpos")}
        raw_total = 0
        for p, sz, off in zip(self.params, sizes, offsets):
            if id(p) in shifts:
                self.ypos_slices.append((off, self.n_param + raw_total, sz, shifts[id(p)]))
                raw_total += sz
        self.flat_param = torch.zeros(self.n_param, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(self.n_param + raw_total, device=dev, dtype=torch.float32)
        self.momentum_buf = torch.zeros(self.n_param, device=dev, dtype=torch.float32)
        self.weight_decay = torch.zeros(self.n_param, device=dev, dtype=torch.float32)
        self.grad_views = []
        for n, p, sz, off in zip(self.names, self.params, sizes, offsets):
            self.flat_param[off:off + sz].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + sz].view_as(p.data)
            self.grad_views.append(self.flat_grad[off:off + sz].view_as(p.data))
            p.grad = None
            self.weight_decay[off:off + sz] = wd_fn(n)
        # element i of the parameter buffer is a ypos entry whose reduced raw sum sits at flat_grad[n_param + ypos_src[i]]
        self.ypos_src = torch.full((self.n_param,), -1, device=dev, dtype=torch.int32)
        for g_off, raw_off, cnt, _ in self.ypos_slices:
            self.ypos_src[g_off:g_off + cnt] = torch.arange(raw_off - self.n_param, raw_off - self.n_param + cnt,
                                                            device=dev, dtype=torch.int32)
        # [lr, momentum, gradient scale] in device memory: the captured graph reads them at replay time
        self.hyper = torch.tensor([lr, momentum, 1.0], device=dev, dtype=torch.float32)
        self.steps = 0
        if self.world > 1:                           # replicas start from rank 0's weights and buffers
            dist.broadcast(self.flat_param, 0, group=self.group)
            for b in model.buffers():
                dist.broadcast(b, 0, group=self.group)
        for _, _, _, shift in self.ypos_slices:
            shift._export_raw = True                 # functional.py stores the raw sums on the module

    def zero_grad(self):
        """Gradients are NOT pre-assigned views: autograd then hands over its freshly produced tensors instead of
        launching one accumulation kernel per parameter (~180 tiny adds per step); ``gather_gradients`` packs them
        into the flat buffer with a multi-tensor copy."""
        for p in self.params:
            p.grad = None

    def gather_gradients(self):
        grads, views = [], []
        for p, v in zip(self.params, self.grad_views):
            if p.grad is None:
                v.zero_()
            else:
                grads.append(p.grad.reshape(v.shape))
                views.append(v)
        torch._foreach_copy_(views, grads)
        for p, v in zip(self.params, self.grad_views):
            p.grad = v                                   # callers (and tests) read p.grad as a view of the flat buffer

    def _collect_raw(self):
        torch._foreach_copy_([self.flat_grad[raw_off:raw_off + cnt] for _, raw_off, cnt, _ in self.ypos_slices],
                             [shift._raw_ypos_grad.reshape(-1) for *_, shift in self.ypos_slices])

    def set_lr(self, lr):
        """takes effect on the next step, eager or replayed (the value is read from device memory)"""
        self.lr = float(lr)
        self.hyper[0:1].fill_(self.lr)

    def adjust_learning_rate(self, epoch, base_lr=0.1, warm_up_epoch=0, step=(60, 80, 100)):
        """the reference's schedule (main.py:342-351) applied to this trainer; returns the learning rate"""
        lr = reference_lr(epoch, base_lr, warm_up_epoch, step)
        self.set_lr(lr)
        return lr

    def reduce_gradients(self):
        """one collective for gradients and raw shift-position sums alike"""
        self._have_raw = all(getattr(s, "_raw_ypos_grad", None) is not None for *_, s in self.ypos_slices)
        if self._have_raw and self.ypos_slices:
            self._collect_raw()
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG if self.flat_grad.is_cuda else dist.ReduceOp.SUM,
                            group=self.group)
            if not self.flat_grad.is_cuda:           # gloo has no AVG
                self.flat_grad.div_(self.world)

    def step(self):
        """K5 on the reduced raw sums, weight decay, SGD momentum / Nesterov (torch.optim.SGD semantics: the momentum
        buffer starts as the first gradient, which a zero-initialised buffer reproduces)."""
        have_raw = getattr(self, "_have_raw", False) and bool(self.ypos_slices)
        if self.flat_param.is_cuda:
            ops.sgd_epilogue(self.flat_param, self.flat_grad, self.momentum_buf, self.weight_decay,
                             self.ypos_src if have_raw else None, self.hyper, self.n_param, self.nesterov)
        else:
            self._step_host(have_raw)
        self.steps += 1

    def _step_host(self, have_raw):
        """The same arithmetic with tensor ops for CPU tensors: used ONLY by the gloo tests of the multi-process logic
        (tests/test_dp.py); CUDA buffers always take the kernel above."""
        g = self.flat_grad[:self.n_param]
        if have_raw:
            raw = self.flat_grad[self.n_param:]
            k5 = torch.where(raw != 0, torch.sign(raw) * 0.01, torch.full_like(raw, 0.0001))
            sel = self.ypos_src >= 0
            g[sel] = k5[self.ypos_src[sel].long()]
        lr, mom = float(self.hyper[0]), float(self.hyper[1])
        d = torch.addcmul(g, self.weight_decay, self.flat_param)
        if mom != 0:
            self.momentum_buf.mul_(mom).add_(d)
            d = d.add(self.momentum_buf, alpha=mom) if self.nesterov else self.momentum_buf
        self.flat_param.add_(d, alpha=-lr)

    def train_step(self, x, label, loss_fn=torch.nn.functional.cross_entropy):
        """forward + backward + all-reduce + SGD on this rank's shard; returns the (local) loss tensor"""
        self.zero_grad()
        loss = loss_fn(self.model(x), label)
        loss.backward()
        self.gather_gradients()
        self.reduce_gradients()
        self.step()
        return loss

    # ------------------------------------------------------------------------------------------------ CUDA graph
    def capture(self, x, label, loss_fn=torch.nn.functional.cross_entropy, warmup=3):
        """Capture one whole training step (zero-grad, forward, loss, backward, all-reduce, K5, SGD) into a CUDA
        graph.  Every kernel of this library is stream-ordered and allocation-free, so the step replays with a
        single launch; ``x`` / ``label`` fix the shapes.  ``warmup`` eager steps run first (they are real steps)."""
        if not x.is_cuda:
            raise RuntimeError("capture needs CUDA tensors")
        self._static_x = x.clone()
        self._static_y = label.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self.train_step(self._static_x, self._static_y, loss_fn)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self.train_step(self._static_x, self._static_y, loss_fn)
        return self

    def replay(self, x, label):
        """one captured step on new data (copied into the static buffers on the current stream)"""
        self._static_x.copy_(x, non_blocking=True)
        self._static_y.copy_(label, non_blocking=True)
        self._graph.replay()
        ops.params_changed()                                       # the captured step rewrites parameters and BatchNorm buffers
        return self._static_loss


class HostPrefetcher:
    """Double-buffered pinned-host -> device staging on a side stream, so that the copy of batch i+1 overlaps the step
    on batch i.  (The reference moves every batch synchronously inside the training loop, ``main.py:400-414``:
    ``data.float().cuda(...)`` right before the forward pass.)

        pf = HostPrefetcher(device)
        pf.put(x0, y0)
        for i in range(n):
            x, y = pf.get()                 # current stream waits for the copy of batch i
            if i + 1 < n: pf.put(x1, y1)    # starts once the work enqueued so far has released the other slot
            step(x, y)
    """

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.slots = [None] * slots
        self.events = [None] * slots
        self._w = self._r = 0

    def put(self, *host_tensors):
        k = self._w % len(self.slots)
        self._w += 1
        if self.stream is None:                                   # CPU (tests of the host logic): plain copies
            self.slots[k] = tuple(t.clone() for t in host_tensors)
            return
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)                              # the slot's previous consumer has been enqueued before
        with torch.cuda.stream(self.stream):
            old = self.slots[k]
            if old is None or any(o.shape != t.shape or o.dtype != t.dtype for o, t in zip(old, host_tensors)) \
                    or len(old) != len(host_tensors):
                old = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors)
            for o, t in zip(old, host_tensors):
                o.copy_(t, non_blocking=True)
            self.slots[k] = old
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.events[k] = ev

    def get(self):
        if self._r >= self._w:
            raise RuntimeError("HostPrefetcher.get() without a matching put()")
        k = self._r % len(self.slots)
        self._r += 1
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.events[k])
        return self.slots[k]


class GraphedInference:
    """One inference call of a module captured in a CUDA graph (every kernel of this library is stream-ordered and
    allocation-free): ``replay(x)`` copies the batch into the static input and launches the whole forward pass at once,
    which removes the ~100 host launches per batch of the eager call (batched window inference,
    ``inference_pipeline.py:342-366``).  ``fn`` maps the static input to the output (default: ``module(x)``)."""

    def __init__(self, module, example, fn=None, warmup=2):
        if not example.is_cuda:
            raise RuntimeError("GraphedInference needs a CUDA example batch")
        self.module = module.eval()
        self._fn = fn if fn is not None else (lambda x: self.module(x))
        self._static_x = example.clone()
        self._warmup = max(warmup, 1)
        self._capture()

    def _decisions(self):
        """host-side kernel choices that depend on parameter VALUES and are baked into a captured graph: whether a
        temporal unit's output shift positions fit the one-kernel inference path (Shift_tcn.out_window_ok)"""
        return tuple(m.out_window_ok() for m in self.module.modules() if hasattr(m, "out_window_ok"))

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self._warmup):
                self._fn(self._static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._choices = self._decisions()
        self._graph = torch.cuda.CUDAGraph()
        # The warm-up calls left the parameter-derived tables (weight images, folded BatchNorm tables, ...) in the cache of
        # ops._frozen_get, so the graph holds only the activation kernels and READS those tables: record them and keep them
        # alive with the graph.  The weights are frozen as of the capture; after changing them call refresh().
        with ops.frozen_recording() as rec, torch.no_grad(), torch.cuda.graph(self._graph):
            self._static_out = self._fn(self._static_x)
        self._tables = rec.tables

    def refresh(self):
        """re-derive the parameter tables the captured graph reads, in place, from the module's current parameters; if the
        new parameters change a kernel choice the graph has baked in (shift positions that left the window of the
        one-kernel temporal unit), the graph is captured again (``replay`` then returns a NEW static output tensor)"""
        if self._decisions() != self._choices:
            self._capture()
            return
        ops.frozen_refresh(self._tables)

    def replay(self, x):
        """x: same shape / dtype as the example (device or pinned host tensor); returns the static output tensor"""
        self._static_x.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._static_out
