"""How far do the temporal shift positions drift in the benchmark?  floor(ypos) histograms per layer after N training steps
of the bench configuration (batch 16 to keep it short):  python tools/ypos_probe.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shiftgcn_b200.modules import Model
from shiftgcn_b200.dp import FlatSGDTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
torch.manual_seed(1)
model = Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph",
              graph_args=dict(labeling_mode="spatial")).to(dev).train()
x = torch.randn(16, 3, 300, 25, 2, device=dev)
y = torch.randint(0, 60, (16,), device=dev)
trainer = FlatSGDTrainer(model, lr=0.1, momentum=0.9, nesterov=True)


def report(tag):
    print(tag)
    for i in range(1, 11):
        t = getattr(model, f"l{i}").tcn1
        for nm in ("shift_in", "shift_out"):
            yp = getattr(t, nm).ypos.detach().float().cpu()
            fl = torch.floor(yp).long()
            vals, cnt = torch.unique(fl, return_counts=True)
            print(f"  l{i}.{nm:9s} ypos [{yp.min():+.3f}, {yp.max():+.3f}]  floor histogram {dict(zip(vals.tolist(), cnt.tolist()))}")


report("at initialisation")
for s in range(steps):
    trainer.train_step(x, y)
torch.cuda.synchronize()
report(f"after {steps} steps")
