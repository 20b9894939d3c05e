import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from shiftgcn_b200 import functional as FN
dev = torch.device("cuda:0")
torch.manual_seed(0)
class Owner:  # workspace holder
    def __init__(self): self._ws = FN.Workspace(); self._h = None
for train in (True, False):
    for (n, T, V, C, D) in ((2, 7, 25, 64, 128), (2, 12, 25, 64, 128), (1, 9, 33, 128, 256)):
        conv = nn.Conv2d(C, D, 1).to(dev); bn = nn.BatchNorm2d(D).to(dev)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        bn.train(train); 
        x = torch.randn(n, T, V, C, device=dev)
        G = torch.randn(n, T, V, D, device=dev)
        # reference in fp64
        conv64 = nn.Conv2d(C, D, 1).to(dev).double(); bn64 = nn.BatchNorm2d(D).to(dev).double()
        conv64.load_state_dict({k: v.double() for k, v in conv.state_dict().items()}); bn64.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in bn.state_dict().items()})
        bn64.train(train)
        x64 = x.double().requires_grad_(True)
        r64 = bn64(conv64(x64.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
        r64.backward(G.double())
        own = Owner()
        xr = x.clone().requires_grad_(True)
        out = FN.SideBranchFn.apply(xr, conv.weight, conv.bias, bn.weight, bn.bias, bn, own, "_h")
        out.backward(G)
        def e(a, b): return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
        print(f"train={train} {(n,T,V,C,D)}: out {e(out, r64):.2e} dx {e(xr.grad, x64.grad):.2e} dW {e(conv.weight.grad, conv64.weight.grad):.2e} "
              f"db {(conv.bias.grad.double()-conv64.bias.grad).abs().max().item():.2e} dgamma {e(bn.weight.grad, bn64.weight.grad):.2e} dbeta {e(bn.bias.grad, bn64.bias.grad):.2e} "
              f"rm {e(bn.running_mean, bn64.running_mean):.2e} rv {e(bn.running_var, bn64.running_var):.2e}", flush=True)
