"""fp32-accurate mode vs the fp64 oracle: per-tensor errors of one module (diagnostic)"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT), sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import model_ref
from shiftgcn_b200 import ops
from shiftgcn_b200.modules import Shift_gcn, Shift_tcn, TCN_GCN_unit
from util import fill_pair, rel_err

dev = torch.device("cuda:0")
PREC = os.environ.get("PREC", "fp32")
if os.environ.get("POISON") == "1":      # uninitialised reads become NaN
    _empty, _empty_like = torch.empty, torch.empty_like

    def empty(*a, **k):
        t = _empty(*a, **k)
        return t.fill_(float("nan")) if t.dtype.is_floating_point else t

    def empty_like(*a, **k):
        t = _empty_like(*a, **k)
        return t.fill_(float("nan")) if t.dtype.is_floating_point else t
    torch.empty, torch.empty_like = empty, empty_like


def one(kind, C, D, V, n, T, s, seed, cap, train=True):
    torch.manual_seed(1)
    if kind == "gcn":
        mod, ref = Shift_gcn(C, D, None, num_point=V), model_ref.RefShiftGcn(C, D, None, num_point=V)
    elif kind == "tcn":
        mod, ref = Shift_tcn(C, C, stride=s), model_ref.RefShiftTcn(C, C, stride=s)
    else:
        mod = TCN_GCN_unit(C, D, None, stride=s, residual=True, num_point=V)
        ref = model_ref.RefUnit(C, D, None, stride=s, residual=True, num_point=V)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, C, T, V, generator=g)
    go = torch.randn(n, D, T // s, V, generator=g)
    mod = mod.to(dev).train(train)
    ref = copy.deepcopy(ref).double().train(train)
    xc = x.to(dev).requires_grad_(True)
    ops.set_max_ctas(cap)
    with ops.precision(PREC):
        out = mod(xc)
        out.backward(go.to(dev))
    torch.cuda.synchronize()
    ops.set_max_ctas(0)
    xr = x.double().requires_grad_(True)
    out_r = ref(xr)
    out_r.backward(go.double())
    errs = [("out", rel_err(out, out_r)), ("gx", rel_err(xc.grad, xr.grad))]
    rp = dict(ref.named_parameters())
    for k, p in mod.named_parameters():
        if p.grad is not None and not k.endswith("pos") and rp[k].grad is not None and rp[k].grad.abs().max() > 1e-9:
            errs.append((k, rel_err(p.grad, rp[k].grad)))
    yp = {k: [round(v, 2) for v in p.detach().flatten().tolist() if abs(v) > 6] for k, p in mod.named_parameters() if k.endswith("ypos")}
    print(f"{kind} {C}->{D} V{V} n{n} T{T} s{s} seed{seed} cap{cap}: " + " ".join(f"{k}={e:.1e}" for k, e in errs), "| big ypos:", yp, flush=True)


CFGS = [("unit", 256, 256, 25, 1, 47, 1, 54), ("unit", 256, 256, 25, 1, 21, 1, 53), ("tcn", 256, 256, 25, 1, 47, 1, 54),
            ("gcn", 256, 256, 25, 1, 47, 1, 54), ("unit", 256, 256, 25, 1, 47, 1, 53), ("unit", 256, 256, 25, 2, 47, 1, 54),
            ("unit", 128, 128, 25, 1, 47, 1, 54), ("unit", 64, 64, 25, 1, 47, 1, 54)]
if os.environ.get("ONLY"):
    CFGS = [CFGS[int(i)] for i in os.environ["ONLY"].split(",")]
for cfg in CFGS:
    for cap in [int(c) for c in os.environ.get("CAPS", "0,3").split(",")]:
        one(*cfg, cap)
