"""Which library formulation of the conv1x1 + BN side branches (down / strided residual) is fastest on rows?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
dev = torch.device("cuda:0")
n, T, V, C, D = 128, 300, 25, 64, 128
x_rows = torch.randn(n, T, V, C, device=dev, requires_grad=True)
conv = nn.Conv2d(C, D, 1).to(dev); bn = nn.BatchNorm2d(D).to(dev)
go = torch.randn(n, T, V, D, device=dev)
def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1)/reps*1e3:9.1f} us", flush=True)
def v_module():      # what modules.py does today
    x0 = x_rows.permute(0, 3, 1, 2)
    y = bn(conv(x0)).permute(0, 2, 3, 1).contiguous()
    y.backward(go)
def v_rows_conv():   # rows as the H axis of a (1, C, rows, 1) channels_last image
    x4 = x_rows.view(1, n * T * V, 1, C).permute(0, 3, 1, 2)
    y4 = F.conv2d(x4, conv.weight, conv.bias)
    y = y4.permute(0, 2, 3, 1).reshape(n * T * V, D)
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, 0.1, 1e-5)
    y.view(n, T, V, D).backward(go)
def v_rows_mm():
    y = torch.addmm(conv.bias, x_rows.view(-1, C), conv.weight.view(D, C).t())
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, 0.1, 1e-5)
    y.view(n, T, V, D).backward(go)
for name, fn in (("module (conv2d+BN2d on NCHW view, .contiguous())", v_module), ("rows conv2d (1,C,rows,1) + batch_norm 2D", v_rows_conv), ("rows addmm fp32 + batch_norm 2D", v_rows_mm)):
    timeit(name, fn)
torch.backends.cuda.matmul.allow_tf32 = True
timeit("rows addmm tf32 + batch_norm 2D", v_rows_mm)
# strided residual
conv2 = nn.Conv2d(C, D, 1, stride=(2, 1)).to(dev)
go2 = torch.randn(n, T // 2, V, D, device=dev)
def r_module():
    x0 = x_rows.permute(0, 3, 1, 2)
    y = bn(conv2(x0)).permute(0, 2, 3, 1).contiguous()
    y.backward(go2)
def r_rows():
    xs = x_rows[:, ::2].contiguous()
    y = torch.addmm(conv2.bias, xs.view(-1, C), conv2.weight.view(D, C).t())
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, 0.1, 1e-5)
    y.view(n, T // 2, V, D).backward(go2)
timeit("residual module (strided conv2d + BN2d)", r_module)
timeit("residual rows (gather even frames, addmm tf32, batch_norm 2D)", r_rows)
