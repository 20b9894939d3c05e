"""Drop-in nn.Modules for the reference's model/shift_gcn.py: ``tcn``, ``Shift_tcn``, ``Shift_gcn``,
``TCN_GCN_unit`` and ``Model`` -- same constructor signatures, attribute names and state_dict keys / dtypes
(SURVEY.md App. D), same train / eval semantics, computed by the fused sm_100a kernels of this package.

Logical tensors stay (N, C, T, V) as in the reference; physically the network runs channels-last
(``torch.channels_last`` strides), which is the reference's own internal ``(n, t, v, c)`` layout
(model/shift_gcn.py:123), so chained units exchange activations without any copy.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as FN
from . import ops
from .shift import Shift

FUSED_CHANNELS = (64, 128, 256)
XPOS_LIMIT = 1e-6     # the fused temporal kernels treat xpos (init U(-1e-8, 1e-8), zero gradient) as 0


def import_class(name):
    components = name.split('.')
    mod = __import__(components[0])
    for comp in components[1:]:
        mod = getattr(mod, comp)
    return mod


def conv_init(conv):
    nn.init.kaiming_normal_(conv.weight, mode='fan_out')
    nn.init.constant_(conv.bias, 0)


def bn_init(bn, scale):
    nn.init.constant_(bn.weight, scale)
    nn.init.constant_(bn.bias, 0)


def _param_device():
    # the reference creates these parameters with device='cuda' (model/shift_gcn.py:90,93,96)
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: shiftgcn_b200 has no CPU path; move the module and its input to a B200 (sm_100a)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{who}: expected float32 activations, got {x.dtype}")


def to_rows(x):
    """logical (n, C, T, V) -> contiguous (n, T, V, C); zero-copy for channels_last inputs"""
    return x.permute(0, 2, 3, 1).contiguous()


def from_rows(r):
    """(n, T, V, C) rows -> logical (n, C, T, V) view (channels_last strides)"""
    return r.permute(0, 3, 1, 2)


def side_supported(conv, num_point):
    """1x1 conv + BatchNorm side branches served by the tensor-core path (FN.side_forward)"""
    return (isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.padding == (0, 0)
            and (conv.in_channels, conv.out_channels) in ((64, 128), (128, 256)) and num_point in (25, 33))


def shift_tables(num_point, in_channels, out_channels):
    """int64 gather tables of the spatial shift (model/shift_gcn.py:108-118), vectorised:
    shift_in[v*C + c] = (v*C + c + c*C) mod (V*C),  shift_out[v*D + d] = (v*D + d - d*D) mod (V*D)."""
    v = np.arange(num_point, dtype=np.int64)[:, None]
    c = np.arange(in_channels, dtype=np.int64)[None, :]
    d = np.arange(out_channels, dtype=np.int64)[None, :]
    tab_in = np.mod(v * in_channels + c + c * in_channels, in_channels * num_point).reshape(-1)
    tab_out = np.mod(v * out_channels + d - d * out_channels, out_channels * num_point).reshape(-1)
    return tab_in, tab_out


class tcn(nn.Module):
    """Strided (k x 1) convolution + BatchNorm: the block residual of the strided units (reference :31-45)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, stride=1):
        super().__init__()
        pad = int((kernel_size - 1) / 2)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(kernel_size, 1), padding=(pad, 0),
                              stride=(stride, 1))
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU()          # present in the reference, never applied
        conv_init(self.conv)
        bn_init(self.bn, 1)

    def forward(self, x):
        return self.bn(self.conv(x))


class _FrozenTables:
    """nn.Module.train() / .eval() of the drop-in modules invalidate the parameter-derived tables that inference keeps
    between calls (ops._frozen_get): a mode switch is where weights usually changed hands."""

    def train(self, mode=True):
        ops.params_changed()
        return super().train(mode)


class Shift_tcn(_FrozenTables, nn.Module):
    """bn -> Shift(stride 1) -> 1x1 conv -> ReLU -> Shift(stride) -> bn2  (reference :48-74)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, stride=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.bn = nn.BatchNorm2d(in_channels)
        self.bn2 = nn.BatchNorm2d(in_channels)
        bn_init(self.bn2, 1)
        self.relu = nn.ReLU(inplace=True)
        self.shift_in = Shift(channel=in_channels, stride=1, init_scale=1)
        self.shift_out = Shift(channel=out_channels, stride=stride, init_scale=1)
        self.temporal_linear = nn.Conv2d(in_channels, out_channels, 1)
        nn.init.kaiming_normal_(self.temporal_linear.weight, mode='fan_out')
        self._ws = FN.Workspace()
        self._xpos_ok = None
        self._xpos_gen = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._xpos_ok = None
        return super()._load_from_state_dict(*args, **kwargs)

    def fused_supported(self, x):
        c, v = x.shape[1], x.shape[3]
        if not (self.in_channels == self.out_channels == c and c in FUSED_CHANNELS and v in (25, 33)):
            return False
        if x.shape[2] // self.shift_out.stride < 1:
            return False
        # one host sync, repeated only after a load_state_dict into this module, an ancestor or the Shift modules
        # themselves.  Editing xpos in place to a non-negligible value is not noticed (call reset_xpos_check()): the
        # reference initialises it to U(-1e-8, 1e-8) and gives it a zero gradient (shift_cuda_kernel.cu:371-395).
        gen = (self.shift_in._load_generation, self.shift_out._load_generation)
        if self._xpos_ok is None or self._xpos_gen != gen:
            with torch.no_grad():
                m = torch.maximum(self.shift_in.xpos.abs().max(), self.shift_out.xpos.abs().max())
                self._xpos_ok = bool(m.item() <= XPOS_LIMIT)
            self._xpos_gen = gen
        return self._xpos_ok

    def reset_xpos_check(self):
        self._xpos_ok = None
        self._out_win = None

    def out_window_ok(self):
        """Can the fused inference kernel (LERP x TSHIFT) serve this unit?  It recomputes the conv output for a halo of
        kWo = 3 frame offsets, so every floor(ypos) of the OUTPUT shift must lie in one window of three values (true at
        the reference's initialisation, ypos ~ U(-1, 1), and for as long as training leaves the positions near it).
        Training MOVES the positions (K5: +-0.01 * lr per step, four floor values after 25 steps in some units), so the
        answer is re-derived (one host sync) whenever ypos may have changed: another storage or tensor version (in-place
        optimizer updates, load_state_dict) or another parameter epoch (raw-pointer writers such as sgcn_sgd_epilogue,
        Module.train() / .eval(), ops.params_changed())."""
        yp = self.shift_out.ypos
        gen = (self.shift_in._load_generation, self.shift_out._load_generation, ops.param_epoch(), yp.data_ptr(), yp._version)
        if getattr(self, "_out_win", None) is None or self._out_win[0] != gen:
            with torch.no_grad():
                fl = torch.floor(self.shift_out.ypos.detach().float())
                ok = bool((fl.max() - fl.min()).item() <= 2 and fl.min().item() >= -8 and fl.max().item() < 8)
            self._out_win = (gen, ok)
        return self._out_win[1]

    def _args(self):
        return (self.bn.weight, self.bn.bias, self.shift_in.xpos, self.shift_in.ypos, self.temporal_linear.weight,
                self.temporal_linear.bias, self.shift_out.xpos, self.shift_out.ypos, self.bn2.weight, self.bn2.bias)

    def forward_rows(self, h_rows, res_rows, relu):
        return FN.TemporalFn.apply(h_rows, res_rows, *self._args(), self, int(relu))

    def forward(self, x):
        _require_cuda(x, "Shift_tcn")
        if self.fused_supported(x):
            return from_rows(self.forward_rows(to_rows(x), None, 0))
        # general path (any channel count, bilinear xpos): stand-alone shift kernels between library ops
        x = self.bn(x)
        x = self.shift_in(x)
        x = self.temporal_linear(x)
        x = self.relu(x)
        x = self.shift_out(x)
        return self.bn2(x)


class Shift_gcn(_FrozenTables, nn.Module):
    """Spatial shift graph convolution (reference :77-142): joint-shift gather, tanh mask, C x D contraction,
    joint-shift gather, BatchNorm1d over (v, d), residual (identity or 1x1 conv + BN), ReLU."""

    def __init__(self, in_channels, out_channels, A, coff_embedding=4, num_subset=3, num_point=25):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_point = num_point
        if in_channels != out_channels:
            self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        else:
            self.down = lambda x: x
        dev = _param_device()
        self.Linear_weight = nn.Parameter(torch.empty(in_channels, out_channels, device=dev))
        nn.init.normal_(self.Linear_weight, 0, math.sqrt(1.0 / out_channels))
        self.Linear_bias = nn.Parameter(torch.zeros(1, 1, out_channels, device=dev))
        self.Feature_Mask = nn.Parameter(torch.zeros(1, num_point, in_channels, device=dev))
        self.bn = nn.BatchNorm1d(num_point * out_channels)
        self.relu = nn.ReLU()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                conv_init(m)
            elif isinstance(m, nn.BatchNorm2d):
                bn_init(m, 1)
        tab_in, tab_out = shift_tables(num_point, in_channels, out_channels)
        self.shift_in = nn.Parameter(torch.from_numpy(tab_in), requires_grad=False)
        self.shift_out = nn.Parameter(torch.from_numpy(tab_out), requires_grad=False)
        self._ws = FN.Workspace()

    def stem_supported(self, x0):
        """the 3-channel first layer (reference :178) has its own kernels (csrc/stem.cu)"""
        return (self.in_channels == 3 and self.out_channels == 64 and x0.shape[1] == 3 and x0.shape[3] == self.num_point
                and self.num_point <= 39 and isinstance(self.down, nn.Sequential))

    def forward_stem_rows(self, x_rows):
        conv, bn2 = self.down[0], self.down[1]
        return FN.StemSpatialFn.apply(x_rows, self.Linear_weight, self.Linear_bias, self.Feature_Mask, self.bn.weight,
                                      self.bn.bias, conv.weight, conv.bias, bn2.weight, bn2.bias, self)

    def fused_supported(self, x0):
        return (self.in_channels in FUSED_CHANNELS and self.out_channels in FUSED_CHANNELS
                and x0.shape[1] == self.in_channels and x0.shape[3] == self.num_point and self.num_point in (25, 33))

    def _args(self):
        return (self.Linear_weight, self.Linear_bias, self.Feature_Mask, self.bn.weight, self.bn.bias)

    def forward_rows(self, x_rows, x0):
        if self.in_channels == self.out_channels:
            res = None
        elif side_supported(self.down[0], self.num_point):
            conv, bn = self.down[0], self.down[1]
            res = FN.SideBranchFn.apply(x_rows, conv.weight, conv.bias, bn.weight, bn.bias, bn, self, "_down_sg")
        else:
            res = to_rows(self.down(x0))
        FN.set_grad_mode(torch.is_grad_enabled())
        return FN.SpatialFn.apply(x_rows, res, *self._args(), self)

    def forward(self, x0):
        _require_cuda(x0, "Shift_gcn")
        if self.stem_supported(x0):
            return from_rows(self.forward_stem_rows(to_rows(x0)))
        if not self.fused_supported(x0):
            # no second (library-op) implementation: the sm_100a kernels are the only path of this package
            raise RuntimeError(
                f"Shift_gcn({self.in_channels}, {self.out_channels}, num_point={self.num_point}) on input "
                f"{tuple(x0.shape)}: shiftgcn_b200 has kernels for 64 / 128 / 256 channels (and the 3 -> 64 first layer) "
                "on the 25-joint (NTU) and 33-landmark (MediaPipe) skeletons only")
        return from_rows(self.forward_rows(to_rows(x0), x0))


class TCN_GCN_unit(_FrozenTables, nn.Module):
    """relu(tcn1(gcn1(x)) + residual(x))  (reference :145-162)."""

    def __init__(self, in_channels, out_channels, A, stride=1, residual=True, num_point=25):
        super().__init__()
        self.gcn1 = Shift_gcn(in_channels, out_channels, A, num_point=num_point)
        self.tcn1 = Shift_tcn(out_channels, out_channels, stride=stride)
        self.relu = nn.ReLU()
        if not residual:
            self.residual = lambda x: 0
            self._res_mode = "none"
        elif (in_channels == out_channels) and (stride == 1):
            self.residual = lambda x: x
            self._res_mode = "identity"
        else:
            self.residual = tcn(in_channels, out_channels, kernel_size=1, stride=stride)
            self._res_mode = "conv"

    def _tcn_fused_for(self, x):
        """would tcn1 take the fused path for the gcn output of this input?"""
        probe = torch.empty((x.shape[0], self.gcn1.out_channels, x.shape[2], x.shape[3]), device="meta")
        return self.tcn1.fused_supported(probe)

    def forward(self, x):
        _require_cuda(x, "TCN_GCN_unit")
        gcn, tcn1 = self.gcn1, self.tcn1
        if self._res_mode == "identity" and gcn.fused_supported(x) and tcn1.fused_supported(x):
            FN.set_grad_mode(torch.is_grad_enabled())
            y = FN.UnitFn.apply(to_rows(x), *gcn._args(), *tcn1._args(), self)
            return y if y.dim() == 1 else from_rows(y)         # 1-D: the pooled-only handle of Model._trunk (inference)
        if (self._res_mode == "conv" and isinstance(gcn.down, nn.Sequential) and gcn.fused_supported(x)
                and side_supported(gcn.down[0], x.shape[3]) and side_supported(self.residual.conv, x.shape[3])
                and x.shape[2] % self.residual.conv.stride[0] == 0 and self._tcn_fused_for(x)):
            d, r = gcn.down, self.residual
            FN.set_grad_mode(torch.is_grad_enabled())
            y = FN.ConvUnitFn.apply(to_rows(x), *gcn._args(), d[0].weight, d[0].bias, d[1].weight, d[1].bias,
                                    *tcn1._args(), r.conv.weight, r.conv.bias, r.bn.weight, r.bn.bias, self)
            return from_rows(y)
        # first layer: the kernel that writes h also sums it for tcn1's first BatchNorm (FN.StemSpatialFn)
        hand_over = gcn.stem_supported(x) and self._tcn_fused_for(x) and tcn1.training == self.training
        gcn._h_stats_for = tcn1 if hand_over else None
        tcn1._h_stats_ready = False
        try:
            h = gcn(x)
        finally:
            gcn._h_stats_for = None
        if tcn1.fused_supported(h):
            tcn1._out_link = getattr(self, "_out_link", None)             # see functional._links
            if self._res_mode == "none":
                res = None
            elif self._res_mode == "conv" and side_supported(self.residual.conv, x.shape[3]) \
                    and x.shape[2] % self.residual.conv.stride[0] == 0:
                conv, bn = self.residual.conv, self.residual.bn
                xs = to_rows(x)
                if conv.stride[0] != 1:
                    xs = xs[:, ::conv.stride[0]].contiguous()           # frames the strided 1x1 conv reads
                res = FN.SideBranchFn.apply(xs, conv.weight, conv.bias, bn.weight, bn.bias, bn, tcn1, "_res_sg")
            else:
                res = to_rows(self.residual(x))
            try:
                return from_rows(tcn1.forward_rows(to_rows(h), res, 1))
            finally:
                tcn1._out_link = None
                tcn1._h_stats_ready = False
        if getattr(tcn1, "_h_stats_ready", False):                   # nobody consumed the sums: leave the buffer clean
            tcn1._ws.get("bn_a", 2 * gcn.out_channels, h.device).zero_()
            tcn1._h_stats_ready = False
        return self.relu(tcn1(h) + self.residual(x))


class Model(_FrozenTables, nn.Module):
    """data_bn -> 10 TCN_GCN_units (3-64-64-64-64-128-128-128-256-256-256, stride 2 at l5 and l8) -> global mean
    over (T, V) and persons -> fc  (reference :165-216)."""

    def __init__(self, num_class=60, num_point=25, num_person=2, graph=None, graph_args=dict(), in_channels=3):
        super().__init__()
        if graph is None:
            raise ValueError()
        Graph = import_class(graph)
        self.graph = Graph(**graph_args)
        A = self.graph.A
        self.data_bn = nn.BatchNorm1d(num_person * in_channels * num_point)
        plan = [(3, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 64, 1, True), (64, 128, 2, True),
                (128, 128, 1, True), (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True), (256, 256, 1, True)]
        for i, (cin, cout, stride, res) in enumerate(plan, start=1):
            setattr(self, f"l{i}", TCN_GCN_unit(cin, cout, A, stride=stride, residual=res, num_point=num_point))
        self.fc = nn.Linear(256, num_class)
        nn.init.normal_(self.fc.weight, 0, math.sqrt(2. / num_class))
        bn_init(self.data_bn, 1)

    def forward(self, x):
        N, C, T, V, M = x.size()
        # data_bn (reference :196-198) normalises feature (m, v, c) over (N, T).  Same arithmetic on the row-major
        # matrix [(N, T), (M, V, C)]: the 2-D batch-norm kernels are ~20x faster than the (N, F, T) ones here, and
        # the result is already channels-last, i.e. the logical (N*M, C, T, V) tensor the units consume without a copy.
        bn = self.data_bn
        if (x.is_cuda and x.dtype == torch.float32 and not x.requires_grad and FN.bn_training(bn) and bn.affine
                and V * M <= 128 and V * C <= 128 and C == 3 and N > 0):
            # native training-mode path: statistics, running buffers, normalisation and the row layout in three kernels
            if not hasattr(self, "_ws"):
                self.__dict__["_ws"] = FN.Workspace()
            rows = FN.DataBnFn.apply(x.contiguous(), bn.weight, bn.bias, bn, self.__dict__["_ws"])
            return self._trunk(rows.permute(0, 3, 1, 2), N, M)
        x = x.permute(0, 2, 4, 3, 1).reshape(N * T, M * V * C)
        track = bn.track_running_stats and bn.running_mean is not None
        if bn.training and track:
            bn.num_batches_tracked += 1
        if bn.momentum is not None:
            factor = bn.momentum
        else:                        # cumulative moving average (nn.BatchNorm semantics for momentum=None)
            factor = 1.0 / float(bn.num_batches_tracked) if (bn.training and track) else 0.0
        x = F.batch_norm(x, bn.running_mean if track else None, bn.running_var if track else None, bn.weight, bn.bias,
                         FN.bn_training(bn), factor, bn.eps)
        x = x.view(N, T, M, V, C).permute(0, 2, 1, 3, 4).reshape(N * M, T, V, C).permute(0, 3, 1, 2)
        return self._trunk(x, N, M)

    def forward_stream(self, joint, modality="joint", parents=None):
        """Logits of this model's ensemble stream straight from the JOINT batch (N, C, T, V, M): the bone / motion
        derivation (inference_pipeline.py:284-309, data_gen/gen_bone_data.py, gen_motion_data.py) runs on the device.
        In inference it is one kernel together with data_bn and the layout change (sgcn_input_stream); in training
        (batch statistics) the derived stream goes through ``forward``."""
        from . import ensemble, ops
        _require_cuda(joint, "Model.forward_stream")
        bone, motion = ensemble.stream_flags(modality)
        bn = self.data_bn
        if self.training or torch.is_grad_enabled() or not bn.track_running_stats or bn.running_mean is None:
            return self.forward(ensemble.derive_modality(joint, modality, parents))
        N, C, T, V, M = joint.shape
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        par = None
        if bone:
            par = torch.as_tensor(parents if parents is not None else ensemble.bone_parents(V), dtype=torch.int32,
                                  device=joint.device)
        rows = ops.input_stream(joint.contiguous().float(), parent=par, motion=motion, rows=True,
                                scale=scale.float().contiguous(), shift=shift.float().contiguous())
        return self._trunk(rows.permute(0, 3, 1, 2), N, M)

    def _head_fusable(self, unit, x):
        """l10 takes the fused identity-unit path and the classifier is a plain fp32 Linear on the same device"""
        ok = (unit._res_mode == "identity" and unit.gcn1.fused_supported(x) and unit.tcn1.fused_supported(x)
              and unit.tcn1.shift_out.stride == 1 and isinstance(self.fc, nn.Linear) and self.fc.weight.is_cuda
              and self.fc.weight.dtype == torch.float32 and self.fc.in_features == unit.gcn1.out_channels)
        if ok:
            self._last_tv = (x.shape[2], x.shape[3])
        return ok

    def _pool_buffer(self, n, C, device):
        key = (n, C, str(device))
        buf = getattr(self, "_pool_bufs", {}).get(key)
        if buf is None:
            buf = torch.zeros((n, C), device=device, dtype=torch.float64)   # sgcn_head_fwd hands it back zeroed
            self.__dict__.setdefault("_pool_bufs", {})[key] = buf
        return buf

    def _trunk(self, x, N, M):
        """l1..l10, pooling over (T, V) and persons, fc (reference :200-216); x: logical (N*M, C, T, V), channels-last"""
        # consecutive units exchange ReLU-masked gradients (functional._links): one dict per unit boundary, attached
        # only for the duration of the unit's forward call so that a unit used on its own never sees a stale link
        ops.restart_traversal()
        links = [dict(masked=False) for _ in range(10)] if torch.is_grad_enabled() else None   # links[9]: l10 -> head
        pool = None
        for i in range(1, 11):
            unit = getattr(self, f"l{i}")
            if links is not None:
                unit._in_link = links[i - 2] if i >= 2 else None
                unit._out_link = links[i - 1] if (i <= 9 or self._head_fusable(unit, x)) else None
            if i == 10 and self._head_fusable(unit, x):
                # the kernel that writes l10's output also accumulates the pooled sums of the head (and, in inference,
                # does not write the output at all)
                pool = self._pool_buffer(x.shape[0], unit.gcn1.out_channels, x.device)
                unit._pool_sums = pool
            try:
                x = unit(x)
            except BaseException:
                if pool is not None:
                    pool.zero_()                                   # never leave partial sums behind
                raise
            finally:
                unit._in_link = unit._out_link = None
                unit._pool_sums = None
        if pool is not None:
            rows = x if x.dim() == 1 else x.permute(0, 2, 3, 1)
            T_last, V_last = getattr(self, "_last_tv")
            return FN.HeadFn.apply(rows, self.fc.weight, self.fc.bias, pool, N, M, T_last * V_last,
                                   links[9] if links is not None else None)
        c_new = x.size(1)
        if x.is_cuda and x.dtype == torch.float32 and c_new % 4 == 0 and x.permute(0, 2, 3, 1).is_contiguous():
            x = FN.PoolRowsFn.apply(x.permute(0, 2, 3, 1))     # same mean; the gradient comes back in the row layout
        else:
            x = x.mean(dim=(2, 3))
        x = x.view(N, M, c_new).mean(1)                       # == view(N, M, C, T*V).mean(3).mean(1), stride-agnostic
        return self.fc(x)
