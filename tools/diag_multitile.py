"""capped-grid vs full-grid runs of the same module on the same input (diagnostic for the multi-tile pipelines)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import model_ref
from shiftgcn_b200 import ops
from shiftgcn_b200.modules import Shift_gcn, Shift_tcn, TCN_GCN_unit

dev = torch.device("cuda:0")


FP32 = os.environ.get("FP32", "0") == "1"
if FP32:
    ops.set_precision("fp32")


def run(mod, x, go, cap, train):
    ops.set_max_ctas(cap)
    mod.train(train)
    mod.zero_grad(set_to_none=True)
    xc = x.clone().requires_grad_(True)
    out = mod(xc)
    out.backward(go)
    torch.cuda.synchronize()
    ops.set_max_ctas(0)
    res = {"out": out.detach().clone(), "gx": xc.grad.clone()}
    for k, p in mod.named_parameters():
        if p.grad is not None and not k.endswith("pos"):
            res["g:" + k] = p.grad.clone()
    return res


def cmp(tag, a, b):
    worst = []
    for k in a:
        d = (a[k].double() - b[k].double()).abs()
        scale = b[k].double().abs().max().item() + 1e-30
        nbad = (d > (2e-6 if FP32 else 1e-4) * scale).sum().item()
        worst.append((d.max().item() / scale, k, nbad, d.numel()))
    worst.sort(reverse=True)
    print(tag, " | ".join(f"{k}: {e:.2e} bad {nb}/{n}" for e, k, nb, n in worst[:4]), flush=True)


cases = [("gcn", 64, 128, 25, 2, 33, 1), ("gcn", 128, 256, 25, 1, 43, 1), ("gcn", 256, 256, 25, 1, 41, 1),
         ("unit", 64, 64, 25, 2, 43, 1), ("unit", 256, 256, 25, 1, 38, 1), ("unit", 128, 256, 25, 1, 46, 2),
         ("unit", 128, 128, 33, 1, 40, 1), ("tcn", 64, 64, 25, 2, 61, 1), ("tcn", 256, 256, 25, 1, 53, 1),
         ("gcn", 64, 64, 25, 8, 300, 1), ("unit", 64, 64, 25, 8, 300, 1), ("unit", 256, 256, 25, 16, 75, 1)]
if FP32:
    cases = [("gcn", 256, 256, 25, 1, 47, 1), ("tcn", 256, 256, 25, 1, 47, 1), ("unit", 256, 256, 25, 1, 47, 1),
             ("tcn", 128, 128, 25, 1, 47, 1), ("gcn", 128, 256, 25, 1, 47, 1)]
for kind, C, D, V, n, T, s in cases:
    torch.manual_seed(1)
    if kind == "gcn":
        mod = Shift_gcn(C, D, None, num_point=V)
    elif kind == "tcn":
        mod = Shift_tcn(C, C, stride=s)
    else:
        mod = TCN_GCN_unit(C, D, None, stride=s, residual=True, num_point=V)
    model_ref.fill_module_(mod)
    mod = mod.to(dev)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(n, C, T, V, generator=g).to(dev)
    go = torch.randn(n, D, T // s, V, generator=g).to(dev)
    for train in (True, False):
        base = run(mod, x, go, 0, train)
        again = run(mod, x, go, 0, train)
        cmp(f"{kind} {C}->{D} V{V} n{n} T{T} s{s} train={train} rerun  :", again, base)
        for cap in (3, 7, 40):
            r = run(mod, x, go, cap, train)
            cmp(f"{kind} {C}->{D} V{V} n{n} T{T} s{s} train={train} cap={cap:3d}:", r, base)
