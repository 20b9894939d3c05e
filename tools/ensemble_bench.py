"""4-stream ensemble (BASELINE.json config 5) sharded over the ranks of one node, timed and checked.

    python tools/ensemble_bench.py                       # one GPU: all four streams
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ensemble_bench.py

Every rank receives the same joint batch (pinned host -> device inside the timed region), derives the stream(s) it owns on
the device (sgcn_input_stream fused with data_bn), runs its model(s) and ONE all-reduce of the alpha-weighted logits is the
ensemble.  Rank 0 also evaluates all four streams by itself once and reports the difference to the sharded result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from shiftgcn_b200 import ensemble as E, ops
from shiftgcn_b200.modules import Model

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local); ops.device_check()
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
N, T, V, M, ncls = 64, 300, 25, 2, 60


def build(k):
    torch.manual_seed(100 + k)                      # same weights for stream k on whichever rank owns it
    return Model(num_class=ncls, num_point=V, num_person=M, graph="graph.ntu_rgb_d.Graph",
                 graph_args=dict(labeling_mode="spatial")).to(dev).eval()


mine = {E.MODALITIES[k]: build(k) for k, _, _ in E.placement(world, rank)}
fns = {name: (lambda jb, m=m, name=name: m.forward_stream(jb, name)) for name, m in mine.items()}
ens = E.StreamEnsemble(fns, num_class=ncls, world_size=world, rank=rank, stream_fn=lambda jb, name: jb)
torch.manual_seed(1)
host = torch.randn(N, 3, T, V, M).pin_memory()
for _ in range(3):
    out = ens.logits(host.to(dev, non_blocking=True))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
steps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = ens.logits(host.to(dev, non_blocking=True))
    top = out.argmax(1).cpu()                       # D2H read of the result
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    full = {E.MODALITIES[k]: build(k) for k in range(4)} if world > 1 else mine
    fns_all = {name: (lambda jb, m=m, name=name: m.forward_stream(jb, name)) for name, m in full.items()}
    ref = E.StreamEnsemble(fns_all, num_class=ncls, stream_fn=lambda jb, name: jb).logits(host.to(dev))
    diff = (ref - out).abs().max().item() / ref.abs().max().item()
    print(json.dumps({"metric": "4-stream ensemble samples/s (NTU 3x300x25x2, batch 64, joint batch from host)",
                      "value": N / (ms.item() * 1e-3), "ms_per_batch": ms.item(), "n_gpus": world, "scaling": "strong",
                      "placement": [E.placement(world, r) for r in range(world)],
                      "rel_diff_vs_single_rank": diff, "top1_equal": bool((ref.argmax(1).cpu() == top).all())}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
