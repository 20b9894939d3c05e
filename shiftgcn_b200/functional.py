"""Launch sequences and autograd glue of the fused Shift-GCN units.

Everything here works on channels-last "row" tensors ``(n, T, V, C)`` (contiguous), i.e. the reference's own
``x.permute(0,2,3,1).contiguous()`` layout (model/shift_gcn.py:123) kept for the whole network.  The nn.Modules in
``modules.py`` convert at the boundary (zero-copy when the caller already uses ``torch.channels_last``).

Kernel schedule of one identity unit in training (a = one activation tensor pass over HBM):

  forward   spatial GEMM (x -> z, BN1d batch sums)            2a      sgcn_rowgemm  SPATIAL / ROT_RAW
            BN1d + residual + ReLU (-> h, BN2d sums of h)     3a      sgcn_bn_res_relu_fwd
            BN + shift + 1x1 conv + ReLU (h -> q)             2a      sgcn_rowgemm  LERP / LINEAR
            output shift, BN2d sums of s                      1a      sgcn_tshift_fwd  mode 0
            shift + BN + residual + ReLU (-> y)               3a      sgcn_tshift_fwd  mode 1
  backward  sums for bn2 / ypos_out                           2a      sgcn_tshift_bwd  mode 0   (g_y arrives with its ReLU mask
            dpre = [q>0] * Shift^T(BN-bwd)                    3a      sgcn_tshift_bwd  mode 1    applied, see _links: no y)
            dp = dpre * W_t                                   2a      sgcn_rowgemm  PLAIN / LINEAR
            dW_t                                              2a      sgcn_wgrad  TEMPORAL
            sums for bn (from dW_t, dbt: no tensor pass)      ~0      sgcn_tshift_in_bwd_sums
            gh = [h>0] * BN-bwd(Shift^T dp), BN1d + ypos_in   4a      sgcn_tshift_in_bwd  mode 1
            g_x = spatial backward-data (+ both residuals)    5a      sgcn_rowgemm  DY / SPATIAL_BWD
            dW                                                3a      sgcn_wgrad  SPATIAL
"""
import threading

import torch

from . import ops

BN_EPS = 1e-5
PREMASK = True      # chained units hand each other ReLU-masked gradients (see _links); False = every unit masks for itself

# Function.forward always runs with grad mode off, and needs_input_grad ignores torch.no_grad(): the modules record the
# caller's grad mode right before .apply so that inference takes the fully fused (nothing saved) kernels.  Per host
# thread: nn.DataParallel (the reference's multi-GPU mode, main.py:294-299) runs one replica per thread.
_state = threading.local()


def set_grad_mode(enabled):
    _state.grad_mode = bool(enabled)


def grad_mode():
    return getattr(_state, "grad_mode", True)


def bn_training(bn):
    """does this BatchNorm normalise with batch statistics?  Each BatchNorm decides for itself, exactly like
    nn.BatchNorm.forward: a frozen bn (``bn.eval()`` inside a training model) keeps its running statistics."""
    return bool(bn.training or not bn.track_running_stats or bn.running_mean is None)


def _bn_args(bn):
    """(gamma, beta, running_mean, running_var, nbt, momentum) of a torch BatchNorm module.  momentum=None (cumulative
    moving average) is passed as -1 with num_batches_tracked incremented HERE: the finalize kernels then use the factor
    1 / num_batches_tracked (include/shiftgcn_b200.h:sgcn_bn_fwd_finalize)."""
    track = bn.track_running_stats and bn.running_mean is not None
    rmean, rvar, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if track else (None, None, None)
    mom = bn.momentum
    if mom is None:
        mom = -1.0
        if track and bn.training:
            nbt.add_(1)
    return bn.weight, bn.bias, rmean, rvar, nbt, mom


class Workspace:
    """Zero-initialised fp64 reduction buffers that the finalize kernels hand back zeroed."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, numel, device):
        key = (name, numel, str(device))
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.zeros(numel, device=device, dtype=torch.float64)
            self._bufs[key] = buf
        return buf

    def get_int(self, name, numel, device):
        key = (name, numel, str(device), "i32")
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.zeros(numel, device=device, dtype=torch.int32)
            self._bufs[key] = buf
        return buf

    def reset(self):
        for b in self._bufs.values():
            b.zero_()


# ================================================================================================ spatial unit
def spatial_forward(x, res, W, bias, mask, bn, ws, fuse_eval, h_stats=None):
    """x: (n,T,V,C) rows; res: None (identity, needs C == D) or (n,T,V,D) rows already normalised.

    Returns (h, saved) where saved holds what the backward needs (None in the fused eval path).
    """
    n, T, V, C = x.shape
    D = W.shape[1]
    R = n * T
    dev = x.device
    mm, mm_rot = ops.mask_prepare(mask.reshape(V, C))
    wimg = ops.weight_image(W, 1, D, D, C)                         # B[n=d][k=c] = W[c][d]
    training = bn_training(bn)
    gamma, beta, rmean, rvar, nbt, mom = _bn_args(bn)
    resid = x if res is None else res
    if fuse_eval:
        mean, invstd, scale, shift = ops.bn_fwd_finalize(None, gamma, beta, rmean, rvar, None, V * D, R, mom, bn.eps, False)
        h = torch.empty((n, T, V, D), device=dev, dtype=torch.float32)
        ops.rowgemm(ops.PRO_SPATIAL, ops.EPI_ROT_FUSED, in0=x, out=h, wimg=wimg, groups=R, V=V, K=C, N=D, pro_a=mm,
                    bias=bias.reshape(D), epi_a=scale, epi_b=shift, res=resid, relu=1)
        return h, None
    z = torch.empty((n, T, V, D), device=dev, dtype=torch.float32)
    stats = ws.get("bn1d", 2 * V * D, dev)
    ops.rowgemm(ops.PRO_SPATIAL, ops.EPI_ROT_RAW, in0=x, out=z, wimg=wimg, groups=R, V=V, K=C, N=D, pro_a=mm,
                bias=bias.reshape(D), stats=stats)
    mean, invstd, scale, shift = ops.bn_fwd_finalize(stats if training else None, gamma, beta, rmean, rvar, nbt, V * D,
                                                      R, mom, bn.eps, training)
    if not training:
        stats.zero_()
    h = torch.empty((n, T, V, D), device=dev, dtype=torch.float32)
    ops.bn_res_relu_fwd(z, resid, h, scale, shift, h_stats, R * V, V, D, relu=1)
    saved = dict(x=x, z=z, h=h, mm=mm, mm_rot=mm_rot, mean=mean, invstd=invstd, training=training)
    return h, saved


def spatial_backward(saved, gh, W, mask, gamma, vd_sums_ready, ws, unit_res=None, side_fn=None, premask=False):
    """gh: grad wrt the gcn output with the ReLU mask already applied when ``vd_sums_ready`` (fused unit),
    otherwise the raw incoming gradient.  unit_res = (g_y, y) adds the block's identity-residual gradient
    (y None: g_y already carries its ReLU mask).  premask: return gx * [x > 0] -- for the unit whose ReLU produced x,
    which would apply that mask itself and can now skip reading its own output (see `_links`).
    side_fn(gh, column sums of gh) -> input gradient of the conv side branches, added inside the kernel.

    Returns dict(gx, gres, dW, dbias, dmask, dgamma, dbeta).
    """
    x, z, h, mm = saved["x"], saved["z"], saved["h"], saved["mm"]
    n, T, V, C = x.shape
    D = W.shape[1]
    R = n * T
    dev = x.device
    vd = ws.get("bn1d_bwd", 2 * V * D, dev)
    if not vd_sums_ready:
        ghm = torch.empty_like(gh)
        ops.relu_bn1d_bwd_stats(gh, h, z, saved["mean"], saved["invstd"], ghm, vd, R, V, D)
        gh = ghm
    fin = ops.bn1d_bwd_finalize(vd, gamma, saved["mean"], saved["invstd"], V, D, R, saved["training"])
    wimg_t = ops.weight_image(W, D, 1, C, D)                       # B[n=c][k=d] = W[c][d]
    gx = torch.empty_like(x)
    dmask_raw = ws.get("dmask", V * C, dev)
    identity = C == D and saved.get("identity_res", True)
    side_dx = side_fn(gh, fin["dbeta"].reshape(V, D).sum(0)) if side_fn is not None else None
    ops.rowgemm(ops.PRO_DY, ops.EPI_SPATIAL_BWD, in0=gh, in1=z, out=gx, wimg=wimg_t, groups=R, V=V, K=D, N=C,
                pro_a=fin["alpha"], pro_b=fin["beta"], pro_c=fin["gamma"], epi_a=mm,
                res=gh if identity else side_dx,
                res2=unit_res[0] if unit_res is not None else None,
                res2m=unit_res[1] if unit_res is not None else None, xin=x, red0=dmask_raw, relu=1 if premask else 0)
    dW = torch.zeros((C, D), device=dev, dtype=torch.float32)
    ops.wgrad(ops.WG_SPATIAL, a_src=x, a_tab0=saved["mm_rot"], b_src=gh, b_src2=z, b_tab0=fin["alpha"], b_tab1=fin["beta"],
              b_tab2=fin["gamma"], dw=dW, groups=R, V=V, CA=C, CB=D)
    dmask = ops.mask_grad_finalize(dmask_raw, mask.reshape(V, C))
    return dict(gx=gx, gres=None if (identity or side_fn is not None) else gh, dW=dW, dbias=fin["dbias"].reshape(1, 1, D),
                dmask=dmask.reshape(1, V, C), dgamma=fin["dgamma"], dbeta=fin["dbeta"])


# ================================================================================================ first spatial unit
def stem_forward(x, W, bias, mask, bn, Wd, bd, bn2, ws, h_stats=None):
    """l1.gcn1 = Shift_gcn(3, 64) with its conv + BN `down` branch (model/shift_gcn.py:82-86,121-142); x: (n,T,V,3).
    h_stats: optional fp64 [64][2] buffer that receives the per-channel sums of h (the statistics the temporal unit's
    first BatchNorm would otherwise collect with a pass of its own)."""
    n, T, V, C = x.shape
    D = W.shape[1]
    R = n * T
    dev = x.device
    mm, _ = ops.mask_prepare(mask.reshape(V, C))
    Wd2 = Wd.reshape(D, C)
    common = dict(groups=R, V=V, D=D, x=x, maskmul=mm, W=W, bias=bias.reshape(D), Wd=Wd2, bd=bd)
    g1, b1, rm1, rv1, nbt1, mom1 = _bn_args(bn)
    g2, b2, rm2, rv2, nbt2, mom2 = _bn_args(bn2)
    tr1, tr2 = bn_training(bn), bn_training(bn2)
    s_vd, s_r = ws.get("stem_vd", 2 * V * D, dev), ws.get("stem_r", 2 * D, dev)
    if tr1 or tr2:
        ops.stem_fwd(0, stats_vd=s_vd, stats_r=s_r, **common)
    mean1, invstd1, sc1, sh1 = ops.bn_fwd_finalize(s_vd if tr1 else None, g1, b1, rm1, rv1, nbt1, V * D, R, mom1, bn.eps, tr1)
    mean2, invstd2, sc2, sh2 = ops.bn_fwd_finalize(s_r if tr2 else None, g2, b2, rm2, rv2, nbt2, D, R * V, mom2, bn2.eps, tr2)
    if tr1 != tr2:                 # the statistics kernel fills both buffers; the frozen BatchNorm does not consume its own
        (s_r if tr1 else s_vd).zero_()
    h = torch.empty((n, T, V, D), device=dev, dtype=torch.float32)
    ops.stem_fwd(1, sc1=sc1, sh1=sh1, sc2=sc2, sh2=sh2, h=h, stats_h=h_stats, **common)
    saved = dict(x=x, h=h, mm=mm, mean1=mean1, invstd1=invstd1, mean2=mean2, invstd2=invstd2, training1=tr1,
                 training2=tr2)
    return h, saved


def stem_backward(saved, g, W, bias, mask, gamma1, Wd, bd, gamma2, ws):
    x, h, mm = saved["x"], saved["h"], saved["mm"]
    n, T, V, C = x.shape
    D = W.shape[1]
    R = n * T
    dev = x.device
    common = dict(groups=R, V=V, D=D, x=x, maskmul=mm, W=W, bias=bias.reshape(D), Wd=Wd.reshape(D, C), bd=bd, g=g, h=h)
    vd, rs = ws.get("stem_vd_bwd", 2 * V * D, dev), ws.get("stem_r_bwd", 2 * D, dev)
    ops.stem_bwd(0, mean1=saved["mean1"], invstd1=saved["invstd1"], mean2=saved["mean2"], invstd2=saved["invstd2"],
                 vd_sums=vd, r_sums=rs, **common)
    f1 = ops.bn1d_bwd_finalize(vd, gamma1, saved["mean1"], saved["invstd1"], V, D, R, saved["training1"])
    f2 = ops.bn1d_bwd_finalize(rs, gamma2, saved["mean2"], saved["invstd2"], 1, D, R * V, saved["training2"])
    dw_raw, dm_raw = ws.get("stem_dw", 8 * D, dev), ws.get("stem_dmask", V * C, dev)
    dx = torch.empty_like(x)
    ops.stem_bwd(1, al=f1["alpha"], be=f1["beta"], ga=f1["gamma"], a2=f2["alpha"], b2=f2["beta"], c2=f2["gamma"],
                 dw_raw=dw_raw, dmask_raw=dm_raw, dx=dx, **common)
    dw = ops.reduce_export(dw_raw).reshape(D, 8)
    dmask = ops.mask_grad_finalize(dm_raw, mask.reshape(V, C))
    return dict(dx=dx, dW=dw[:, 0:3].t().contiguous(), dbias=f1["dbias"].reshape(1, 1, D), dmask=dmask.reshape(1, V, C),
                dgamma1=f1["dgamma"], dbeta1=f1["dbeta"], dWd=dw[:, 3:6].reshape(D, C, 1, 1).contiguous(),
                dbd=dw[:, 6].contiguous(), dgamma2=f2["dgamma"], dbeta2=f2["dbeta"])


# ================================================================================================ conv + BN side branches
def side_forward(x, Wd, bd, bn, ws, gs=1):
    """BatchNorm2d(Conv2d_1x1(x)) on rows (`down` of Shift_gcn, model/shift_gcn.py:82-86; `tcn` residual, :31-45).

    x: (rows, C) with rows a multiple of V*... (any row count that is a multiple of the tile group size is fine: the
    branch has no joint structure).  The batch statistics of the conv output follow from the first two moments of x:
        mean_r = Wd mu + bd,   var_r[d] = Wd[d] Cov(x) Wd[d]^T
    so they cost one Gram-matrix contraction of x on the tensor cores (rounding errors of ~1e6 products average out)
    plus one channel-sum pass, and the normalisation is FOLDED into the conv weights before the only full-size GEMM:
    r' = x (sc*Wd)^T + (beta - sc*(Wd mu)).  The raw conv output is never stored.
    gs > 1: the conv has frame stride gs -- every kernel reads frames 0, gs, 2gs, .. of x in place (no gathered copy).
    """
    n, Tx, V, C = x.shape
    T = Tx // gs
    rows = n * T * V
    D = Wd.shape[0]
    dev = x.device
    training = bn_training(bn)
    gamma, beta, rmean, rvar, nbt, mom = _bn_args(bn)
    Wd2 = Wd.detach().reshape(D, C)
    saved = dict(x=x, training=training, gs=gs)
    st = XX = counter = None
    if training:
        st = ws.get("side_sx", 2 * C, dev)
        if gs == 1:
            ops.channel_stats(x, st, rows, C)
        else:
            ops.channel_stats_groups(x, st, n * T, V, C, gs)
        XX = torch.zeros((C, C), device=dev, dtype=torch.float32)
        ops.wgrad(ops.WG_PLAIN, a_src=x, b_src=x, dw=XX, groups=n * T, V=V, CA=C, CB=C, a_gs=gs, b_gs=gs)
        counter = ws.get_int("side_counter", 1, dev)
    f = ops.side_fold(Wd2, None if bd is None else bd.detach(), gamma.detach(), beta.detach(), rmean, rvar,
                      nbt if training else None, rows, bn.eps, mom, training,
                      sx_sums=st, XX=XX, counter=counter)
    if training:
        saved.update(sx=f["sx"], XX=XX)
    wimg = ops.weight_image(f["Wf"], C, 1, D, C)                   # B[n=d][k=c] = Wf[d][c]
    out = torch.empty((n, T, V, D), device=dev, dtype=torch.float32)
    ops.rowgemm(ops.PRO_PLAIN, ops.EPI_LINEAR, in0=x, out=out, wimg=wimg, groups=n * T, V=V, K=C, N=D, bias=f["bf"], relu=0,
                in0_gs=gs if gs > 1 else 0)
    saved.update(mean_r=f["mean_r"], invstd=f["invstd"])
    return out, saved


def side_backward(saved, G, sg, Wd, bd, gamma, ws, accum_into=None):
    """Gradients of the conv + BN branch without touching the conv output: with P = x^T G (one plain contraction),
    XX = x^T x and sx from the forward, everything else is C x C x D arithmetic on the device (fp64):
        dgamma = invstd * (rowsum(Wd * P^T) + (bd - mean_r) * sg),  dbeta = sg,
        dr = al*G + be*r + ga  (BatchNorm backward as an affine map),   dWd = al*P^T + be*(Wd XX + bd sx^T) + ga sx^T,
        dx = G (al*Wd) + x (Wd^T be Wd) + Wd^T (be*bd + ga)           -- ONE GEMM over the concatenation [G | x].
    accum_into: an existing gradient wrt the whole x; dx is then ADDED to it in the GEMM epilogue (at the strided
    frames when the conv has a frame stride) and the returned dx is that tensor.
    """
    x = saved["x"]
    gs = saved.get("gs", 1)
    n, Tx, V, C = x.shape
    T = Tx // gs
    rows = n * T * V
    D = Wd.shape[0]
    dev = x.device
    training = saved["training"]
    P = torch.zeros((C, D), device=dev, dtype=torch.float32)
    ops.wgrad(ops.WG_PLAIN, a_src=x, b_src=G, dw=P, groups=n * T, V=V, CA=C, CB=D, a_gs=gs)
    r = ops.side_bwd(P, sg.float().contiguous(), Wd.detach().reshape(D, C), None if bd is None else bd.detach(),
                     gamma.detach(), saved["invstd"], saved["mean_r"], rows, training, sx=saved.get("sx"),
                     XX=saved.get("XX"))
    wimg = ops.weight_image(r["Wcat"], 1, C, C, D + C)             # B[n=c][k] = Wcat[k][c]: out[c] = sum_k in[k] Wcat[k][c]
    if accum_into is None:
        if gs != 1:
            raise RuntimeError("side_backward: a strided branch adds its input gradient into accum_into")
        dx = torch.empty((n, T, V, C), device=dev, dtype=torch.float32)
    else:
        dx = accum_into
    ops.rowgemm(ops.PRO_PLAIN, ops.EPI_LINEAR, in0=G, in1=x, out=dx, wimg=wimg, groups=n * T, V=V, K=D + C, N=C, k0=D,
                bias=r["kvec"], relu=0, in1_gs=gs if gs > 1 else 0, out_gs=gs if gs > 1 else 0,
                accum=0 if accum_into is None else 1)
    return dict(dx=dx, dWd=r["dWd"].reshape(Wd.shape), dbd=r["dbd"], dgamma=r["dgamma"], dbeta=r["dbeta"])


# ================================================================================================ temporal unit
def temporal_forward(h, res, relu, bn, ypos_in, Wt, bt, ypos_out, bn2, stride, ws, h_stats_ready, pool=None,
                     store_out=True, fuse_eval=False):
    """h: (n,T,V,C) rows -> y: (n,T/stride,V,C) rows = [relu](bn2(Shift_s(relu(conv(Shift_1(bn(h)))))) + res)"""
    n, T, V, C = h.shape
    To = T // stride
    dev = h.device
    if res is not None and tuple(res.shape) != (n, To, V, C):
        # e.g. a stride-2 conv residual of an odd-length sequence has ceil(T/2) frames: the reference fails on the add
        raise RuntimeError(f"temporal unit: residual of shape {tuple(res.shape)} does not match the output "
                           f"{(n, To, V, C)} (rows layout)")
    tr_a, tr_b = bn_training(bn), bn_training(bn2)
    ga, ba, rma, rva, nbta, moma = _bn_args(bn)
    stats_a = ws.get("bn_a", 2 * C, dev)
    if tr_a and not h_stats_ready:
        ops.channel_stats(h, stats_a, n * T * V, C)
    mean_a, invstd_a, scale_a, shift_a = ops.bn_fwd_finalize(stats_a if tr_a else None, ga, ba, rma, rva, nbta, C,
                                                              n * T * V, moma, bn.eps, tr_a)
    wimg = ops.weight_image(Wt, C, 1, C, C)                        # B[n=co][k=ci] = Wt[co][ci]
    ypos_in_eff = ypos_in.detach().contiguous()
    if fuse_eval and not tr_a and not tr_b and stride == 1 and V == 25 and pool is None:
        # inference: the whole unit in ONE kernel (sgcn_rowgemm LERP x TSHIFT) -- q never crosses HBM
        gb, bb, rmb, rvb, nbtb, momb = _bn_args(bn2)
        _, _, scale_b, shift_b = ops.bn_fwd_finalize(None, gb, bb, rmb, rvb, nbtb, C, n * To * V, momb, bn2.eps, False)
        y = torch.empty((n, To, V, C), device=dev, dtype=torch.float32)
        ops.rowgemm(ops.PRO_LERP, ops.EPI_TSHIFT, in0=h, out=y, wimg=wimg, groups=n * T, V=V, K=C, N=C, T=T,
                    pro_a=scale_a, pro_b=shift_a, pro_c=ypos_in_eff, bias=bt, res2=ypos_out.detach().contiguous(),
                    epi_a=scale_b, epi_b=shift_b, res=res, relu=relu)
        return y, None
    q = torch.empty((n, T, V, C), device=dev, dtype=torch.float32)
    ops.rowgemm(ops.PRO_LERP, ops.EPI_LINEAR, in0=h, out=q, wimg=wimg, groups=n * T, V=V, K=C, N=C, T=T,
                pro_a=scale_a, pro_b=shift_a, pro_c=ypos_in_eff, bias=bt, relu=1)
    ypos_out_eff = (ypos_out.detach() + 0.5) if stride != 1 else ypos_out.detach().contiguous()   # cuda/shift.py:14-19
    gb, bb, rmb, rvb, nbtb, momb = _bn_args(bn2)
    stats_b = ws.get("bn_b", 2 * C, dev)
    if tr_b:
        ops.tshift_fwd(0, q=q, ypos_eff=ypos_out_eff, n_samples=n, T_in=T, T_out=To, V=V, C=C, stride=stride,
                       stats=stats_b)
    mean_b, invstd_b, scale_b, shift_b = ops.bn_fwd_finalize(stats_b if tr_b else None, gb, bb, rmb, rvb, nbtb, C,
                                                              n * To * V, momb, bn2.eps, tr_b)
    # pool: fp64 [n, C] buffer that receives the sums over (t, v) of y (head of the model, model/shift_gcn.py:212-214);
    # store_out=False (inference, pooled output only): y is never written
    y = torch.empty((n, To, V, C), device=dev, dtype=torch.float32) if (store_out or pool is None) else None
    ops.tshift_fwd(1, q=q, ypos_eff=ypos_out_eff, n_samples=n, T_in=T, T_out=To, V=V, C=C, stride=stride, res=res,
                   out=y, scale=scale_b, shift=shift_b, relu=relu, stats=pool)
    saved = dict(h=h, q=q, y=y, relu=relu, stride=stride, training_a=tr_a, training_b=tr_b, ypos_in_eff=ypos_in_eff,
                 ypos_out_eff=ypos_out_eff, mean_a=mean_a, invstd_a=invstd_a, scale_a=scale_a, shift_a=shift_a,
                 mean_b=mean_b, invstd_b=invstd_b)
    return y, saved


def temporal_backward(saved, gy, gamma_a, Wt, gamma_b, ws, spatial_saved=None, spatial_ws=None, relu_h=False,
                      want_raw=False, gy_masked=False):
    """Returns dict(gh, dWt, dbt, dgamma_a, dbeta_a, dgamma_b, dbeta_b, gx_in, gy_in, gx_out, gy_out[, raw_in, raw_out]).

    With ``spatial_saved`` (fused unit) the last kernel also applies the gcn ReLU mask and accumulates the
    BN1d backward sums, so the spatial backward can start immediately.
    """
    h, q, y = saved["h"], saved["q"], saved["y"]
    n, T, V, C = h.shape
    stride = saved["stride"]
    To = T // stride
    dev = h.device
    relu = 1 if (saved["relu"] and not gy_masked) else 0   # a pre-masked g_y needs no look at y (1a less per pass)
    sums5 = ws.get("tshift_bwd", 5 * C, dev)
    common = dict(q=q, gy=gy, y=y if relu else None, relu=relu, ypos_eff=saved["ypos_out_eff"], mean=saved["mean_b"],
                  invstd=saved["invstd_b"], n_samples=n, T_in=T, T_out=To, V=V, C=C, stride=stride)
    ops.tshift_bwd(0, sums=sums5, **common)
    fb = ops.tshift_bwd_finalize(sums5, gamma_b, saved["invstd_b"], C, n * To * V, n, saved["training_b"], input_shift=False,
                                 want_raw=want_raw)
    dpre = torch.empty((n, T, V, C), device=dev, dtype=torch.float32)
    dbias_acc = ws.get("dbt", C, dev)
    ops.tshift_bwd(1, k1=fb["k1"], m1=fb["m1"], m2=fb["m2"], dpre=dpre, dbias=dbias_acc, **common)
    dbt = ops.reduce_export(dbias_acc)
    wimg_t = ops.weight_image(Wt, 1, C, C, C)                      # B[n=ci][k=co] = Wt[co][ci]
    dp = torch.empty((n, T, V, C), device=dev, dtype=torch.float32)
    ops.rowgemm(ops.PRO_PLAIN, ops.EPI_LINEAR, in0=dpre, out=dp, wimg=wimg_t, groups=n * T, V=V, K=C, N=C, relu=0)
    dWt = torch.zeros((C, C), device=dev, dtype=torch.float32)
    ops.wgrad(ops.WG_TEMPORAL, a_src=dpre, b_src=h, b_tab0=saved["scale_a"], b_tab1=saved["shift_a"],
              b_tab2=saved["ypos_in_eff"], dw=dWt, groups=n * T, V=V, CA=C, CB=C, T=T)
    # BN(h) backward sums without a pass over dp and h: sum du = W_t^T dbt - boundary frames, sum du*U = colsum(W_t * dW_t)
    # (include/shiftgcn_b200.h:SgcnTShiftInSums); the gated statistics pass only runs for a degenerate BatchNorm weight
    sums3 = ws.get("tshift_in_bwd", 3 * C, dev)
    gate = ws.get_int("tshift_in_gate", 1, dev)
    common_in = dict(dp=dp, h=h, ypos_eff=saved["ypos_in_eff"], mean=saved["mean_a"], invstd=saved["invstd_a"],
                     scale=saved["scale_a"], shift=saved["shift_a"], n_samples=n, T=T, V=V, C=C)
    ops.tshift_in_bwd_sums(dp=dp, ypos_eff=saved["ypos_in_eff"], Wt=Wt.contiguous(), dWt=dWt, dbt=dbt,
                           mean=saved["mean_a"], invstd=saved["invstd_a"], scale=saved["scale_a"], shift=saved["shift_a"],
                           sums=sums3, gate=gate, n_samples=n, T=T, V=V, C=C)
    ops.tshift_in_bwd(0, sums=sums3, gate=gate, **common_in)
    fa = ops.tshift_bwd_finalize(sums3, gamma_a, saved["invstd_a"], C, n * T * V, n, saved["training_a"], input_shift=True)
    gh = torch.empty((n, T, V, C), device=dev, dtype=torch.float32)
    pos = ws.get("tshift_in_pos", C, dev)
    if spatial_saved is not None:
        vd = spatial_ws.get("bn1d_bwd", 2 * V * C, dev)
        ops.tshift_in_bwd(1, k1=fa["k1"], m1=fa["m1"], m2=fa["m2"], gh=gh, relu_h=1, z=spatial_saved["z"],
                          zmean=spatial_saved["mean"], zinvstd=spatial_saved["invstd"], vd_sums=vd, pos_sums=pos,
                          **common_in)
    else:
        ops.tshift_in_bwd(1, k1=fa["k1"], m1=fa["m1"], m2=fa["m2"], gh=gh, relu_h=1 if relu_h else 0, pos_sums=pos,
                          **common_in)
    gx_in, gy_in, raw_in = ops.shift_pos_finalize(pos, C, n, want_raw=want_raw)    # K4 sums of the apply pass -> K5
    out = dict(gh=gh, dWt=dWt.reshape(C, C, 1, 1), dbt=dbt, dgamma_a=fa["dgamma"], dbeta_a=fa["dbeta"],
               dgamma_b=fb["dgamma"], dbeta_b=fb["dbeta"], gx_in=gx_in, gy_in=gy_in, gx_out=fb["gx"],
               gy_out=fb["gy"])
    if want_raw:
        out["raw_in"], out["raw_out"] = raw_in, fb["raw"]
    return out


# ================================================================================================ autograd
def _stash(ctx, **groups):
    """Register every tensor of the given dicts through ctx.save_for_backward (outputs included: keeping them as plain
    ctx attributes would create an output -> grad_fn -> ctx -> output reference cycle and leak a whole step of
    activations until the cyclic GC runs); non-tensor entries stay on ctx."""
    tensors, layout = [], {}
    for gname, d in groups.items():
        if d is None:
            layout[gname] = None
            continue
        entries = {}
        for k, v in d.items():
            if torch.is_tensor(v):
                entries[k] = ("t", len(tensors))
                tensors.append(v)
            else:
                entries[k] = ("v", v)
        layout[gname] = entries
    ctx._layout = layout
    ctx.save_for_backward(*tensors)


def _unstash(ctx):
    tensors = ctx.saved_tensors
    out = {}
    for gname, entries in ctx._layout.items():
        out[gname] = None if entries is None else {k: (tensors[v] if kind == "t" else v) for k, (kind, v) in entries.items()}
    return out


def _links(ctx, owner, produces_gx=True):
    """Pre-masked gradients between chained units.  `Model.forward` joins consecutive units with a small dict per
    boundary (`owner._in_link` / `owner._out_link`, live only during that unit's forward call).  A unit whose input x is
    the ReLU output y of the previous unit can return gx * [x > 0] at no cost (its spatial backward kernel reads x
    anyway); it says so by setting link["masked"] when its backward runs, and the previous unit -- whose backward runs
    next and would multiply the incoming gradient by [y > 0] itself -- then skips reading y in three kernels.  The
    mask is idempotent, so a producer that does not mask (any non-fused path) leaves the flag unset and nothing
    changes."""
    ctx.in_link = getattr(owner, "_in_link", None) if produces_gx else None
    ctx.out_link = getattr(owner, "_out_link", None)


def _premask_now(ctx):
    """called by the backward of the unit that produces gx: decide and publish whether gx carries the mask"""
    pm = bool(PREMASK and ctx.in_link is not None)
    if ctx.in_link is not None:
        ctx.in_link["masked"] = pm
    return pm


def _gy_masked(ctx):
    return bool(ctx.out_link is not None and ctx.out_link.get("masked"))


class SpatialFn(torch.autograd.Function):
    """Shift_gcn.forward (model/shift_gcn.py:121-142) on rows; ``res`` = down(x0) rows or None for identity."""

    @staticmethod
    def forward(ctx, x, res, W, bias, mask, gamma, beta, module):
        need_grad = grad_mode() and any(ctx.needs_input_grad)
        h, saved = spatial_forward(x, res, W, bias, mask, module.bn, module._ws,
                                   fuse_eval=(not bn_training(module.bn) and not need_grad))
        ctx.module = module
        if saved is not None:
            saved["identity_res"] = res is None
        _stash(ctx, s=saved, p=dict(W=W, mask=mask, gamma=gamma))
        return h

    @staticmethod
    def backward(ctx, g):
        st = _unstash(ctx)
        p = st["p"]
        r = spatial_backward(st["s"], g.contiguous(), p["W"], p["mask"], p["gamma"], False, ctx.module._ws)
        if r["gres"] is not None:        # column sums of the gradient handed to the `down` branch (SideBranchFn)
            V, D = st["s"]["x"].shape[2], p["W"].shape[1]
            ctx.module._down_sg = r["dbeta"].reshape(V, D).sum(0)
        return r["gx"], r["gres"], r["dW"], r["dbias"], r["dmask"], r["dgamma"], r["dbeta"], None


class StemSpatialFn(torch.autograd.Function):
    """Shift_gcn(3, 64).forward including its conv + BN `down` branch (model/shift_gcn.py:82-86,121-142) on rows."""

    @staticmethod
    def forward(ctx, x, W, bias, mask, g1, b1, Wd, bd, g2, b2, module):
        # TCN_GCN_unit.forward names the temporal unit that consumes h (module._h_stats_for, live for this call only):
        # the kernel that writes h then also sums it for that unit's first BatchNorm
        tcn = getattr(module, "_h_stats_for", None)
        h_stats = None
        if tcn is not None and bn_training(tcn.bn):
            h_stats = tcn._ws.get("bn_a", 2 * W.shape[1], x.device)
            tcn._h_stats_ready = True
        h, saved = stem_forward(x, W, bias, mask, module.bn, Wd, bd, module.down[1], module._ws, h_stats=h_stats)
        ctx.module = module
        _stash(ctx, s=saved, p=dict(W=W, bias=bias, mask=mask, g1=g1, Wd=Wd, bd=bd, g2=g2))
        return h

    @staticmethod
    def backward(ctx, g):
        st = _unstash(ctx)
        p = st["p"]
        r = stem_backward(st["s"], g.contiguous(), p["W"], p["bias"], p["mask"], p["g1"], p["Wd"], p["bd"], p["g2"],
                          ctx.module._ws)
        return (r["dx"], r["dW"], r["dbias"], r["dmask"], r["dgamma1"], r["dbeta1"], r["dWd"], r["dbd"], r["dgamma2"],
                r["dbeta2"], None)


class SideBranchFn(torch.autograd.Function):
    """BatchNorm2d(Conv2d 1x1 (x)) on rows: the `down` branch of Shift_gcn (model/shift_gcn.py:82-86) and the conv
    residual of the strided units (:31-45, 157-158; the caller gathers the strided frames)."""

    @staticmethod
    def forward(ctx, x, Wd, bd, gamma, beta, bn, owner, hint_attr):
        out, saved = side_forward(x, Wd, bd, bn, owner._ws)
        ctx.owner, ctx.hint_attr = owner, hint_attr
        _stash(ctx, s=saved, p=dict(Wd=Wd, bd=bd, gamma=gamma))
        return out

    @staticmethod
    def backward(ctx, G):
        st = _unstash(ctx)
        p = st["p"]
        G = G.contiguous()
        hint = getattr(ctx.owner, ctx.hint_attr, None)             # column sums of G left by the kernel that produced it
        setattr(ctx.owner, ctx.hint_attr, None)
        if hint is None:
            D = G.shape[-1]
            acc = ctx.owner._ws.get("side_sg", 2 * D, G.device)
            ops.channel_stats(G, acc, G.numel() // D, D)
            hint = ops.reduce_export(acc).reshape(D, 2)[:, 0]
        r = side_backward(st["s"], G, hint, p["Wd"], p["bd"], p["gamma"], ctx.owner._ws)
        return r["dx"], r["dWd"], r["dbd"], r["dgamma"], r["dbeta"], None, None, None


class TemporalFn(torch.autograd.Function):
    """Shift_tcn.forward (model/shift_gcn.py:65-74) on rows, optionally fused with the block residual + ReLU
    of TCN_GCN_unit.forward (:160-162)."""

    @staticmethod
    def forward(ctx, h, res, ga, ba, xpos_in, ypos_in, Wt, bt, xpos_out, ypos_out, gb, bb, module, relu):
        need_grad = grad_mode() and any(ctx.needs_input_grad)
        stats_ready = bool(getattr(module, "_h_stats_ready", False))    # left by the producer of h (StemSpatialFn)
        module._h_stats_ready = False
        y, saved = temporal_forward(h, res, relu, module.bn, ypos_in, Wt.reshape(Wt.shape[0], Wt.shape[1]), bt,
                                    ypos_out, module.bn2, module.shift_out.stride, module._ws, stats_ready,
                                    fuse_eval=not need_grad and module.out_window_ok())
        ctx.module = module
        ctx.has_res = res is not None
        _links(ctx, module, produces_gx=False)
        _stash(ctx, t=saved, p=dict(ga=ga, Wt=Wt, gb=gb))
        return y

    @staticmethod
    def backward(ctx, gy):
        st = _unstash(ctx)
        saved, p = st["t"], st["p"]
        Wt = p["Wt"]
        gy = gy.contiguous()
        want_raw = getattr(ctx.module.shift_in, "_export_raw", False)
        masked = _gy_masked(ctx)
        r = temporal_backward(saved, gy, p["ga"], Wt.reshape(Wt.shape[0], Wt.shape[1]), p["gb"], ctx.module._ws,
                              want_raw=want_raw, gy_masked=masked)
        if want_raw:        # raw (pre-K5) means for the post-all-reduce constraint of dp.FlatSGDTrainer
            ctx.module.shift_in._raw_ypos_grad, ctx.module.shift_out._raw_ypos_grad = r["raw_in"], r["raw_out"]
        gres = None
        if ctx.has_res:
            gres = ops.relu_mask_grad(gy, saved["y"]) if (saved["relu"] and not masked) else gy
            ctx.module._res_sg = r["dbeta_b"] if saved["relu"] else None   # column sums of gres (conv residual branch)
        return (r["gh"], gres, r["dgamma_a"], r["dbeta_a"], r["gx_in"], r["gy_in"], r["dWt"], r["dbt"], r["gx_out"],
                r["gy_out"], r["dgamma_b"], r["dbeta_b"], None, None)


class UnitFn(torch.autograd.Function):
    """A whole identity-residual TCN_GCN_unit (in == out channels, stride 1; model/shift_gcn.py:155-156,160-162):
    relu(tcn1(gcn1(x)) + x) with every cross-stage fusion enabled."""

    @staticmethod
    def forward(ctx, x, W, bias, mask, g1, b1, ga, ba, xpos_in, ypos_in, Wt, bt, xpos_out, ypos_out, gb, bb, unit):
        gcn, tcn = unit.gcn1, unit.tcn1
        n, T, V, C = x.shape
        need_grad = grad_mode() and any(ctx.needs_input_grad)
        fuse_eval = not bn_training(gcn.bn) and not need_grad
        # the BatchNorm2d statistics of h come out of the kernel that writes h (not in the fully fused eval kernel)
        h_stats = tcn._ws.get("bn_a", 2 * C, x.device) if (bn_training(tcn.bn) and not fuse_eval) else None
        h, s_saved = spatial_forward(x, None, W, bias, mask, gcn.bn, gcn._ws, fuse_eval=fuse_eval, h_stats=h_stats)
        pool = getattr(unit, "_pool_sums", None)                   # set by Model._trunk on the last unit only
        y, t_saved = temporal_forward(h, x, 1, tcn.bn, ypos_in, Wt.reshape(C, C), bt, ypos_out, tcn.bn2, 1,
                                      tcn._ws, h_stats_ready=h_stats is not None, pool=pool,
                                      store_out=need_grad or pool is None,
                                      fuse_eval=not need_grad and tcn.out_window_ok())
        if y is None:                                              # pooled-only inference: nothing to return but a handle
            return x.new_empty((0,))
        ctx.unit = unit
        _links(ctx, unit)
        if s_saved is not None:
            s_saved["identity_res"] = True
        _stash(ctx, s=s_saved, t=t_saved, p=dict(W=W, mask=mask, g1=g1, ga=ga, Wt=Wt, gb=gb))
        return y

    @staticmethod
    def backward(ctx, gy):
        st = _unstash(ctx)
        s_saved, t_saved, p = st["s"], st["t"], st["p"]
        unit = ctx.unit
        gy = gy.contiguous()
        C = p["W"].shape[0]
        want_raw = getattr(unit.tcn1.shift_in, "_export_raw", False)
        masked = _gy_masked(ctx)
        t = temporal_backward(t_saved, gy, p["ga"], p["Wt"].reshape(C, C), p["gb"], unit.tcn1._ws, spatial_saved=s_saved,
                              spatial_ws=unit.gcn1._ws, want_raw=want_raw, gy_masked=masked)
        if want_raw:
            unit.tcn1.shift_in._raw_ypos_grad, unit.tcn1.shift_out._raw_ypos_grad = t["raw_in"], t["raw_out"]
        s = spatial_backward(s_saved, t["gh"], p["W"], p["mask"], p["g1"], True, unit.gcn1._ws,
                             unit_res=(gy, None if masked else t_saved["y"]), premask=_premask_now(ctx))
        return (s["gx"], s["dW"], s["dbias"], s["dmask"], s["dgamma"], s["dbeta"], t["dgamma_a"], t["dbeta_a"],
                t["gx_in"], t["gy_in"], t["dWt"], t["dbt"], t["gx_out"], t["gy_out"], t["dgamma_b"], t["dbeta_b"], None)


class ConvUnitFn(torch.autograd.Function):
    """A whole TCN_GCN_unit whose side branches are 1x1 conv + BatchNorm (l5, l8: in != out channels, frame stride 2;
    model/shift_gcn.py:82-86, 152-162): relu(tcn1(gcn1(x)) + residual(x)), sequenced so that every gradient that flows
    into x is added inside the spatial backward kernel instead of by separate full-tensor additions."""

    @staticmethod
    def forward(ctx, x, W, bias, mask, g1, b1, Wd, bd, gd, bdn, ga, ba, xpos_in, ypos_in, Wt, bt, xpos_out, ypos_out, gb,
                bb, Wr, br, gr, brn, unit):
        gcn, tcn = unit.gcn1, unit.tcn1
        stride = tcn.shift_out.stride
        D = W.shape[1]
        need_grad = grad_mode() and any(ctx.needs_input_grad)
        fuse_eval = not bn_training(gcn.bn) and not need_grad
        res_d, sd_saved = side_forward(x, Wd, bd, gcn.down[1], gcn._ws)
        h_stats = tcn._ws.get("bn_a", 2 * D, x.device) if (bn_training(tcn.bn) and not fuse_eval) else None
        h, s_saved = spatial_forward(x, res_d, W, bias, mask, gcn.bn, gcn._ws, fuse_eval=fuse_eval, h_stats=h_stats)
        res_r, sr_saved = side_forward(x, Wr, br, unit.residual.bn, tcn._ws, gs=stride)   # strided in place
        y, t_saved = temporal_forward(h, res_r, 1, tcn.bn, ypos_in, Wt.reshape(D, D), bt, ypos_out, tcn.bn2, stride,
                                      tcn._ws, h_stats_ready=h_stats is not None)
        ctx.unit = unit
        _links(ctx, unit)
        if s_saved is not None:
            s_saved["identity_res"] = False
        _stash(ctx, s=s_saved, t=t_saved, sd=sd_saved, sr=sr_saved,
               p=dict(W=W, mask=mask, g1=g1, Wd=Wd, bd=bd, gd=gd, ga=ga, Wt=Wt, gb=gb, Wr=Wr, br=br, gr=gr))
        return y

    @staticmethod
    def backward(ctx, gy):
        st = _unstash(ctx)
        s_saved, t_saved, sd_saved, sr_saved, p = st["s"], st["t"], st["sd"], st["sr"], st["p"]
        unit = ctx.unit
        gcn, tcn = unit.gcn1, unit.tcn1
        gy = gy.contiguous()
        D = p["W"].shape[1]
        stride = t_saved["stride"]
        want_raw = getattr(tcn.shift_in, "_export_raw", False)
        masked = _gy_masked(ctx)
        t = temporal_backward(t_saved, gy, p["ga"], p["Wt"].reshape(D, D), p["gb"], tcn._ws, spatial_saved=s_saved,
                              spatial_ws=gcn._ws, want_raw=want_raw, gy_masked=masked)
        if want_raw:
            tcn.shift_in._raw_ypos_grad, tcn.shift_out._raw_ypos_grad = t["raw_in"], t["raw_out"]
        gres = gy if masked else ops.relu_mask_grad(gy, t_saved["y"])   # gradient into the conv residual branch
        rd, rr = {}, {}

        def side_fn(gh, sg):
            # input gradients of both conv branches in ONE tensor: the `down` branch writes it, the (strided) residual
            # branch adds into it inside its GEMM epilogue, the spatial backward kernel adds it to gx
            rd.update(side_backward(sd_saved, gh, sg, p["Wd"], p["bd"], p["gd"], gcn._ws))
            rr.update(side_backward(sr_saved, gres, t["dbeta_b"], p["Wr"], p["br"], p["gr"], tcn._ws, accum_into=rd["dx"]))
            return rd["dx"]

        s = spatial_backward(s_saved, t["gh"], p["W"], p["mask"], p["g1"], True, gcn._ws, side_fn=side_fn,
                             premask=_premask_now(ctx))
        return (s["gx"], s["dW"], s["dbias"], s["dmask"], s["dgamma"], s["dbeta"], rd["dWd"], rd["dbd"], rd["dgamma"],
                rd["dbeta"], t["dgamma_a"], t["dbeta_a"], t["gx_in"], t["gy_in"], t["dWt"], t["dbt"], t["gx_out"],
                t["gy_out"], t["dgamma_b"], t["dbeta_b"], rr["dWd"], rr["dbd"], rr["dgamma"], rr["dbeta"], None)


class DataBnFn(torch.autograd.Function):
    """The model's input BatchNorm in training mode + the change to the row layout (model/shift_gcn.py:196-198):
    x (N, C, T, V, M) -> rows (N*M, T, V, C).  Statistics in the input layout (sgcn_data_bn_stats), running buffers in
    sgcn_bn_fwd_finalize, normalisation fused with the layout change (sgcn_input_stream); the backward needs gamma.grad
    and beta.grad only (the input is data)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, ws):
        N, C, T, V, M = x.shape
        F_ = M * V * C
        stats = ws.get("data_bn", 2 * F_, x.device)
        ops.data_bn_stats(x, stats)
        g, b, rmean, rvar, nbt, mom = _bn_args(bn)
        mean, invstd, scale, shift = ops.bn_fwd_finalize(stats, g, b, rmean, rvar, nbt, F_, N * T, mom, bn.eps, True)
        rows = ops.input_stream(x, rows=True, scale=scale, shift=shift)
        ctx.save_for_backward(x, mean, invstd)
        ctx.ws = ws
        return rows

    @staticmethod
    def backward(ctx, g):
        x, mean, invstd = ctx.saved_tensors
        sums = ctx.ws.get("data_bn_bwd", 2 * mean.numel(), x.device)
        ops.data_bn_bwd(g.contiguous(), x, mean, invstd, sums)
        out = ops.reduce_export(sums).view(-1, 2)
        return None, out[:, 1].contiguous(), out[:, 0].contiguous(), None, None


class HeadFn(torch.autograd.Function):
    """Global pooling over (T, V) and persons + fc (model/shift_gcn.py:212-216) from the pooled sums that the last unit's
    output kernel left in ``pool_sums`` ([N*M, C] fp64).  ``y`` (the last unit's output rows, or an empty handle in
    inference) only ties the node into the autograd graph; its gradient comes back in the row layout, written once."""

    @staticmethod
    def forward(ctx, y, W, b, pool_sums, N, M, rows_per_person, link=None):
        pooled, logits = ops.head_fwd(pool_sums, W.detach().contiguous(), None if b is None else b.detach(), N, M,
                                      rows_per_person * M)
        ctx.dims = (N, M, rows_per_person, tuple(y.shape))
        ctx.has_bias = b is not None
        # link: the out-link of the last unit (functional._links).  y is that unit's ReLU output; masking the broadcast
        # gradient here costs one read of y, and the unit's backward then skips y in three kernels.
        ctx.link = link if (PREMASK and link is not None and y.dim() == 4 and y.is_contiguous()) else None
        if ctx.link is not None:
            ctx.save_for_backward(pooled, W, y)
        else:
            ctx.save_for_backward(pooled, W)
        return logits

    @staticmethod
    def backward(ctx, dl):
        pooled, W = ctx.saved_tensors[:2]
        N, M, rpp, yshape = ctx.dims
        dW, db, gpool = ops.head_bwd(dl.contiguous().float(), pooled, W.detach().contiguous(), N, M, rpp * M, ctx.has_bias)
        gy = None
        if ctx.needs_input_grad[0]:
            mask_y = ctx.saved_tensors[2] if ctx.link is not None else None
            gy = ops.bcast_rows(gpool, rpp, 1.0, mask_y=mask_y).view(yshape)
            if ctx.link is not None:
                ctx.link["masked"] = True
        return gy, dW, db, None, None, None, None, None


class PoolRowsFn(torch.autograd.Function):
    """Mean over the frames and joints of a row tensor (n, T, V, C) -> (n, C)  (model/shift_gcn.py:212-214).  The
    backward writes the broadcast gradient once, in the row layout the last unit's backward kernels read (autograd's
    own backward expands, scales and then copies: three full-tensor passes)."""

    @staticmethod
    def forward(ctx, rows):
        ctx.shape = rows.shape
        return rows.mean(dim=(1, 2))

    @staticmethod
    def backward(ctx, g):
        n, T, V, C = ctx.shape
        return ops.bcast_rows(g.contiguous().float(), T * V, 1.0 / (T * V)).view(n, T, V, C)
