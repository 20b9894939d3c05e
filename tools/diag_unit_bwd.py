"""Backward intermediates of one identity TCN_GCN_unit against the fp64 oracle (diagnostic)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT), sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from oracle import model_ref
from shiftgcn_b200 import ops, functional as FN
from shiftgcn_b200.modules import TCN_GCN_unit
from util import fill_pair, rel_err

dev = torch.device("cuda:0")
C, V, n, T, seed = [int(v) for v in os.environ.get("CFG", "256,25,2,47,54").split(",")]
PREC = os.environ.get("PREC", "fp32")
torch.manual_seed(1)
mod = TCN_GCN_unit(C, C, None, stride=1, residual=True, num_point=V)
ref = model_ref.RefUnit(C, C, None, stride=1, residual=True, num_point=V)
fill_pair(mod, ref)
g = torch.Generator().manual_seed(seed)
x = torch.randn(n, C, T, V, generator=g)
go = torch.randn(n, C, T, V, generator=g)
mod = mod.to(dev).train()
ref = copy.deepcopy(ref).double().train()

# ---- oracle with retained intermediates
xr = x.double().requires_grad_(True)
h = ref.gcn1(xr); h.retain_grad()
t = ref.tcn1
u = t.bn(h); p = t.shift_in(u); p.retain_grad()
lin = t.temporal_linear(p); lin.retain_grad()
q = F.relu(lin); s = t.shift_out(q); s.retain_grad()
y = F.relu(t.bn2(s) + xr)
y.backward(go.double())
rows = lambda a: a.permute(0, 2, 3, 1).contiguous()
want = dict(dpre=rows(lin.grad), dp=rows(p.grad), gh=rows(h.grad * (h > 0)), gx=rows(xr.grad), q=rows(q.detach()), h=rows(h.detach()),
            y=rows(y.detach()))

# ---- product, recording the outputs of the backward kernels
rec = {}
orig = dict(tshift_bwd=ops.tshift_bwd, rowgemm=ops.rowgemm, tshift_in_bwd=ops.tshift_in_bwd, tshift_fwd=ops.tshift_fwd,
            bn_res_relu_fwd=ops.bn_res_relu_fwd)


def wrap(name):
    def f(*a, **k):
        r = orig[name](*a, **k)
        torch.cuda.synchronize()
        if name == "tshift_bwd" and a[0] == 1:
            rec["dpre"] = k["dpre"].clone()
        if name == "tshift_in_bwd" and a[0] == 1:
            rec["gh"] = k["gh"].clone()
        if name == "rowgemm":
            if a[0] == ops.PRO_PLAIN and "dp" not in rec and "dpre" in rec:
                rec["dp"] = k["out"].clone()
            if a[0] == ops.PRO_DY:
                rec["gx"] = k["out"].clone()
            if a[0] == ops.PRO_LERP:
                rec["q"] = k["out"].clone()
        if name == "bn_res_relu_fwd":
            rec["h"] = a[2].clone()
        if name == "tshift_fwd" and a[0] == 1:
            rec["y"] = k["out"].clone()
        return r
    return f


for k_ in orig:
    setattr(ops, k_, wrap(k_))
xc = x.to(dev).requires_grad_(True)
with ops.precision(PREC):
    out = mod(xc)
    out.backward(go.to(dev))
torch.cuda.synchronize()
for k_ in ("h", "q", "y", "dpre", "dp", "gh", "gx"):
    a, b = rec[k_].double().cpu(), want[k_]
    d = (a - b).abs()
    scale = b.abs().max().item()
    bad = (d > 1e-3 * scale)
    idx = bad.nonzero()
    print(f"{k_:5s} rel {d.max().item() / scale:.2e}  bad {bad.sum().item()}/{d.numel()}", flush=True)
    if idx.numel():
        print("   first bad (n,t,v,c):", idx[:6].tolist(), " frames:", sorted(set(idx[:, 1].tolist()))[:20],
              " channels:", sorted(set(idx[:, 3].tolist()))[:20], " samples:", sorted(set(idx[:, 0].tolist())))
print("ypos_in big:", [(i, round(v, 2)) for i, v in enumerate(mod.tcn1.shift_in.ypos.tolist()) if abs(v) > 4])
print("ypos_out big:", [(i, round(v, 2)) for i, v in enumerate(mod.tcn1.shift_out.ypos.tolist()) if abs(v) > 4])
