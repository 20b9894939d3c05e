"""Where does the end-to-end step lose time against the device-resident step?  (bench.py's `e2e` vs `value`)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shiftgcn_b200.dp import FlatSGDTrainer
from shiftgcn_b200.modules import Model
dev = torch.device("cuda:0")
torch.manual_seed(1)
model = Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph", graph_args=dict(labeling_mode="spatial")).to(dev).train()
hx = torch.randn(64, 3, 300, 25, 2).pin_memory(); hy = torch.randint(0, 60, (64,)).pin_memory()
dx, dy = hx.to(dev), hy.to(dev)
tr = FlatSGDTrainer(model)
for _ in range(3): tr.train_step(dx, dy)
tr.capture(dx, dy, warmup=1)
def run(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{name:50s} device {e0.elapsed_time(e1)/n:8.2f} ms/step   wall {(t1-t0)*1e3/n:8.2f} ms/step", flush=True)
run("replay, device inputs, no sync", lambda: tr.replay(dx, dy))
run("replay, device inputs, item() each step", lambda: tr.replay(dx, dy).item())
run("replay, pinned host inputs, no sync", lambda: tr.replay(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)))
run("replay, pinned host inputs, item() each step", lambda: tr.replay(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)).item())
run("replay, host -> static buffers directly, item()", lambda: (tr._static_x.copy_(hx, non_blocking=True), tr._static_y.copy_(hy, non_blocking=True), tr._graph.replay(), tr._static_loss.item()))
run("H2D only (11.5 MB)", lambda: hx.to(dev, non_blocking=True))
