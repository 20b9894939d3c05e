"""Inference throughput of the NTU-60 model at batch 64 when the batch is walked in micro-batches whose activation
tensors fit the 126 MB L2 (samples are independent in eval mode): one CUDA graph per variant, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shiftgcn_b200.modules import Model
from shiftgcn_b200.dp import GraphedInference

dev = torch.device("cuda:0")
torch.manual_seed(1)
model = Model(num_class=60, num_point=25, num_person=2, graph="graph.ntu_rgb_d.Graph", graph_args=dict(labeling_mode="spatial")).to(dev).eval()
x = torch.randn(64, 3, 300, 25, 2, device=dev)
ref = None
for mb in [int(v) for v in os.environ.get("MBS", "64,32,16,8,4").split(",")]:
    def fn(inp, mb=mb):
        return torch.cat([model(inp[i:i + mb]) for i in range(0, inp.shape[0], mb)])
    with torch.no_grad():
        g = GraphedInference(model, x, fn=fn)
        for _ in range(3):
            out = g.replay(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = g.replay(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None:
        ref = out.clone()
    print(f"micro-batch {mb:3d}: {ms:7.3f} ms/step  {64 / ms * 1e3:9.1f} samples/s   max |diff| vs first {float((out - ref).abs().max()):.2e}", flush=True)
