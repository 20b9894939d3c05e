import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False
from shiftgcn_b200 import modules as M
from oracle import model_ref
from util import fill_pair
dev = torch.device("cuda:0")
for train in (True, False):
    torch.manual_seed(1)
    mod = M.Shift_gcn(64, 128, None, num_point=25); ref = model_ref.RefShiftGcn(64, 128, None, num_point=25)
    fill_pair(mod, ref)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 7, 25, generator=g); go = torch.randn(2, 128, 7, 25, generator=g)
    res = {}
    for side in (True, False):
        M_side = M.side_supported
        if not side: M.side_supported = lambda *a, **k: False
        m = mod.to(dev).train(train)
        for p in m.parameters(): p.grad = None
        xc = x.to(dev).requires_grad_(True)
        out = m(xc); out.backward(go.to(dev)); torch.cuda.synchronize()
        res[side] = (out.detach().clone(), xc.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
        M.side_supported = M_side
    o1, g1, p1 = res[True]; o0, g0, p0 = res[False]
    print(f"train={train}: out diff {(o1-o0).abs().max().item():.3e} (max {o0.abs().max().item():.2f}); dx diff {(g1-g0).abs().max().item():.3e} (max {g0.abs().max().item():.2f})")
    flips = ((o1 > 0) != (o0 > 0)).sum().item()
    dd = (o1 - o0).abs()
    print(f"  relu flips {flips}; out abs diff mean {dd.mean().item():.2e} p99 {dd.flatten().kthvalue(int(dd.numel()*0.99)).values.item():.2e}; near-zero count(|o0|<5e-3 & o0>0) {((o0>0)&(o0<5e-3)).sum().item()}")
    d = (g1 - g0).abs()
    idx = torch.nonzero(d > 0.02 * g0.abs().max())
    print("  bad dx entries:", idx.shape[0], idx[:8].tolist())
    for k in p1:
        print(f"   {k:28s} {(p1[k]-p0[k]).abs().max().item():.3e} / {p0[k].abs().max().item():.3e}")
