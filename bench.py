#!/usr/bin/env python
"""bench.py -- Shift-GCN fwd+bwd samples/s on synthetic NTU tensors (BASELINE.json's metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload ntu60-train|ntu120-train|ntu60-infer|mediapipe-train|ensemble]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of 64 synthetic samples per GPU: forward, cross-entropy, backward,
(N > 1: one flat-buffer NCCL gradient all-reduce) and the SGD-Nesterov update -- nothing is skipped.  Prints ONE JSON line
from rank 0.  Keys beyond the base contract:
  roofline      dominant kernel of the step (largest total device time in a CUDA-event profiling pass after the timed
                region): algorithmic bytes per launch / mean launch duration, against MEASURED_PEAKS.json
  roofline_step whole-step algorithmic bytes (930.0 MB/sample NTU training) / step time
  cpu_baseline  the oracle port (oracle/model_ref.py, torch CPU, all host threads) on a bounded sample, N = 1 only, timed
                BEFORE any GPU work; its `c1` entry is SURVEY.md section 8d's C1 (one Shift_gcn(64,64) eval forward on
                (32,64,300,25)) on the host cores with this library's time for the same shape beside it
Default workload: ntu60-train on one GPU, ntu120-train (BASELINE.json config 4: same tensors, 120 classes) on several;
``--workload ensemble`` is config 5 (four streams placed on the ranks, one all-reduce of the weighted logits).
  e2e           same metric through the public nn.Module API with HOST inputs: pinned H2D copy of every batch and a
                D2H read of the loss inside the timed region
``--impl reference`` times the reference's CPU implementation of the path (the oracle port; the reference itself is
Python and cannot travel to the GPU box) on the same workload, bounded per step.  ``--impl reference-gpu`` is the
informational second baseline BASELINE.json asks for: the reference's PyTorch path (oracle/model_ref.py) with the
reference's own compiled ``shift_cuda`` kernels (oracle/_ref) on one B200.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (num_class, V, M, T, per-GPU batch, training, algorithmic MB/sample, fwd GFLOP/sample)
    "ntu60-train": (60, 25, 2, 300, 64, True, 930.0, 7.139),
    "ntu120-train": (120, 25, 2, 300, 64, True, 930.0, 7.139),
    "ntu60-infer": (60, 25, 2, 300, 64, False, 196.0, 7.139),
    "mediapipe-train": (2, 33, 1, 300, 64, True, 613.8, 4.711),
}


def _ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture
    of the same step (profiles/r*_traffic_*.json, newest round first); None when no capture exists for this kernel /
    workload.  A file holds either one record or {"workload": .., "kernels": {name: bytes per launch}}."""
    import glob
    family = "train" if workload in ("ntu60-train", "ntu120-train") else workload
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_*.json")), reverse=True):
        try:
            with open(path) as f:
                rec = json.load(f)
        except (OSError, ValueError):
            continue
        rec_family = "train" if rec.get("workload", "ntu60-train") in ("ntu60-train", "ntu120-train") else rec.get("workload")
        if rec_family != family:
            continue
        if rec.get("kernel") == kernel:
            return float(rec["dram_bytes_per_launch"]), os.path.relpath(path, ROOT)
        if kernel in rec.get("kernels", {}):
            return float(rec["kernels"][kernel]), os.path.relpath(path, ROOT)
    return None, None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def _cpu_reference_rate(workload, batch, steps, warmup, train):
    """fwd(+bwd) samples/s of the oracle port on the host cores (bounded sample)."""
    import torch
    from oracle import model_ref
    num_class, V, M, T = WORKLOADS[workload][:4]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1)
    model = model_ref.RefModel(num_class=num_class, num_point=V, num_person=M).train(train)
    x = torch.randn(batch, 3, T, V, M)
    y = torch.randint(0, num_class, (batch,))
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True) if train else None

    def step():
        if train:
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(x)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, cores


def _c1_cpu(reps=5):
    """Shift_gcn(64, 64, num_point=25) eval forward on x0 = randn(32, 64, 300, 25), CPU, all
    host threads, 2 warm-up + `reps` timed calls -> (best seconds, median seconds)."""
    import torch
    from oracle import model_ref
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1)
    m = model_ref.RefShiftGcn(64, 64, None, num_point=25).eval()
    x = torch.randn(32, 64, 300, 25)
    ts = []
    with torch.no_grad():
        for i in range(2 + reps):
            t0 = time.perf_counter()
            m(x)
            if i >= 2:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def _c1_gpu(dev, reps=20):
    """the same C1 call through this package's Shift_gcn on the GPU (eval, no_grad; the input is 61 MB: a 256 MB buffer
    is rewritten between calls so every call starts with a cold L2) -> median milliseconds"""
    import torch
    from shiftgcn_b200.modules import Shift_gcn
    torch.manual_seed(1)
    m = Shift_gcn(64, 64, None, num_point=25).to(dev).eval()
    x = torch.randn(32, 64, 300, 25, device=dev).contiguous(memory_format=torch.channels_last)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    ts = []
    with torch.no_grad():
        for i in range(3 + reps):
            flush.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m(x)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def default_workload(args):
    if args.workload:
        return args.workload
    return "ntu60-train" if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) <= 1 else "ntu120-train"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    if workload == "ensemble":
        emit({"impl": "reference", "unavailable": "the ensemble workload has no CPU reference arm (four models; see tests)"})
        return
    train = WORKLOADS[workload][5]
    batch = args.ref_batch
    rate, dt, cores = _cpu_reference_rate(workload, batch, args.steps, max(args.warmup, 1), train)
    line = {
        "impl": "reference", "metric": "Shift-GCN fwd+bwd samples/sec (NTU 3x300x25x2)" if train else "Shift-GCN inference samples/sec",
        "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload, "batch_per_step": batch,
                                        "note": "reference path restated in oracle/model_ref.py (torch CPU); bounded sample"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {batch} ({'fwd+bwd+SGD' if train else 'fwd'})"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_reference_gpu(args):
    """--impl reference-gpu (informational second baseline of BASELINE.json's north_star): the reference's PyTorch
    path -- restated module by module in oracle/model_ref.py (index_select gathers, einsum, BatchNorm, cuDNN 1x1
    convs, torch.optim.SGD) -- on ONE B200 with the reference's OWN compiled shift_cuda kernels (oracle/_ref, built
    from /root/reference's sources).  None of this repository's kernels run in this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import build_ref_ext, model_ref, shift_torch
    if build_ref_ext.built_library() is None:
        emit({"impl": "reference-gpu", "unavailable": "oracle/_ref/shift_cuda_ref*.so was not built (needs /root/reference)"})
        return
    shift_torch.REF_EXT = build_ref_ext.load()
    num_class, V, M, T, batch, train, algo_mb, _ = WORKLOADS[args.workload]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    model = model_ref.RefModel(num_class=num_class, num_point=V, num_person=M).to(dev).train(train)
    x = torch.randn(batch, 3, T, V, M, device=dev)
    y = torch.randint(0, num_class, (batch,), device=dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True) if train else None

    def step():
        if train:
            opt.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
        else:
            with torch.no_grad():
                model(x)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    peak, peak_src = _peaks()
    emit({"impl": "reference-gpu", "metric": "Shift-GCN fwd+bwd samples/sec (NTU 3x300x25x2)" if train else f"Shift-GCN samples/sec ({args.workload})",
          "value": batch / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
          "ms_per_step": ms, "higher_is_better": True, "dtype": "f32 (cuDNN convolutions may use TF32, torch default)",
          "data": "synthetic", "config": {"workload": args.workload, "per_gpu_batch": batch,
                                          "note": "reference modules restated in oracle/model_ref.py + the reference's compiled shift_cuda"},
          "roofline_step": {"achieved": algo_mb * 1e6 * batch / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": algo_mb * 1e6 * batch / (ms * 1e-3) / 1e9 / peak},
          "clocks": clocks, "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9, "gpu_launches": 0})


def run_b200(args):
    import torch
    import torch.distributed as dist

    from shiftgcn_b200 import ops
    from shiftgcn_b200.dp import FlatSGDTrainer, GraphedInference
    from shiftgcn_b200.modules import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def note(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    num_class, V, M, T, batch, train, algo_mb, fwd_gf = WORKLOADS[args.workload]
    # CPU baseline: N = 1 only, and BEFORE any GPU work or process group exists -- no GPU spins in a collective while
    # the host cores are being timed, and the host cores are not shared with launch threads
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores = _cpu_reference_rate(args.workload, args.ref_batch, 2, 1, train)
        c1_best, c1_med = _c1_cpu()
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"2 steps of batch {args.ref_batch} of the same workload ({'fwd+bwd+SGD' if train else 'fwd'}), "
                         "oracle/model_ref.py on torch CPU",
               "c1": {"what": "SURVEY 8d C1: Shift_gcn(64,64,num_point=25) eval forward on (32,64,300,25), 16 samples",
                      "cpu_s_best": c1_best, "cpu_s_median": c1_med, "cpu_samples_per_s": 16 / c1_med}}
        note("CPU baseline done")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.device_check()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    graph = "graph.ntu_rgb_d.Graph" if V == 25 else "graph.mediapipe_pose.Graph"
    torch.manual_seed(1)                                     # identical init on every rank (main.py:24-28 seeds 1)
    model = Model(num_class=num_class, num_point=V, num_person=M, graph=graph,
                  graph_args=dict(labeling_mode="spatial")).to(dev).train(train)
    torch.manual_seed(1 + rank)                              # rank-seeded synthetic shard
    host_x = torch.randn(batch, 3, T, V, M).pin_memory()
    host_y = torch.randint(0, num_class, (batch,)).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    trainer = FlatSGDTrainer(model, lr=0.1, momentum=0.9, nesterov=True) if train else None
    note("model and trainer ready")

    graphed = False
    launches_per_step = None

    infer_graph = None

    def step(x, y):
        if train:
            return trainer.replay(x, y) if graphed else trainer.train_step(x, y)
        if infer_graph is not None:
            return infer_graph.replay(x)
        with torch.no_grad():
            return model(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host):
        """max-over-ranks device time of n_steps (CUDA events, barrier + synchronize on both sides)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        if from_host:
            # every step copies ITS batch from pinned host memory on the launching stream and its result is read back on
            # the host before the next step starts (the plain user-level loop; dp.HostPrefetcher can overlap the copy
            # with the previous step, but its benefit was not stable across boxes, so the headline uses the plain loop)
            for _ in range(n_steps):
                x = host_x.to(dev, non_blocking=True)
                y = host_y.to(dev, non_blocking=True)
                out = step(x, y)
                last = float(out.item()) if train else float(out[0, 0].item())     # D2H read of the step's result
        else:
            for _ in range(n_steps):
                step(dev_x, dev_y)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n_steps, last

    for _ in range(max(args.warmup, 3)):
        step(dev_x, dev_y)
    barrier()
    note("eager warm-up done")
    if train and not args.no_graph:
        # the whole step (fwd, loss, bwd, all-reduce, K5, SGD) as ONE CUDA graph: removes ~1000 host launches
        l0 = ops.LAUNCHES
        trainer.capture(dev_x, dev_y, warmup=1)
        launches_per_step = (ops.LAUNCHES - l0) // 2         # one eager warm-up step + the captured one
        graphed = True
        for _ in range(2):
            step(dev_x, dev_y)
        barrier()
        note("graph captured and replayed")
    if not train and not args.no_graph:
        l0 = ops.LAUNCHES
        infer_graph = GraphedInference(model, dev_x, warmup=1)
        launches_per_step = (ops.LAUNCHES - l0) // 2         # one eager warm-up call + the captured one
        graphed = True
        for _ in range(2):
            step(dev_x, dev_y)
        barrier()
        note("inference graph captured and replayed")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    ms_step, _ = timed(args.steps, from_host=False)
    launches = launches_per_step if graphed else (ops.LAUNCHES - launches0) // args.steps
    timed(2, from_host=True)                                 # untimed: first pinned-copy / allocator use of the host path
    ms_e2e, last = timed(args.steps, from_host=True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- profiling pass (rank 0): per-call CUDA events -> dominant kernel of the step
    roofline = None
    peak, peak_src = _peaks()
    # every rank runs the profiled step (it contains the gradient all-reduce); rank 0 keeps the events
    torch.cuda.synchronize()
    ops.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The eager step is host-bound (~0.3 s of Python for ~25 ms of kernels): without a head start the GPU idles between
    # launches and every event pair would also time the host gap in front of its kernel.  A spin kernel keeps the stream
    # busy while the host enqueues the whole step, so the events then measure back-to-back device execution.
    torch.cuda._sleep(int(0.6 * getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1.9e6) * 1e3))
    e0.record()
    if train:
        trainer.train_step(dev_x, dev_y)                     # eager (not the graph): per-call events need real launches
    else:
        with torch.no_grad():
            model(dev_x)                                     # eager, like the training branch
    e1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
        agg = {}
        for name, a, b, nbytes in prof:
            t, cnt, by = agg.get(name, (0.0, 0, 0))
            agg[name] = (t + a.elapsed_time(b), cnt + 1, by + nbytes)
        total_kernel_ms = sum(v[0] for v in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][0])
        name, (t_ms, cnt, by) = top
        achieved = by / cnt / (t_ms / cnt * 1e-3) / 1e9
        traffic, traffic_src = _ncu_traffic(name, args.workload)
        roofline = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "launches_per_step": cnt,
                    "ms_per_launch": t_ms / cnt, "share_of_step_kernel_time": t_ms / max(total_kernel_ms, 1e-9),
                    "algorithmic_bytes_per_launch": by / cnt,
                    "breakdown_ms": {k: round(v[0], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
                    # the same classes as fractions of the HBM peak on their own algorithmic bytes (0 = no tensor pass)
                    "breakdown_frac": {k: round(v[2] / max(v[0] * 1e-3, 1e-12) / 1e9 / peak, 3)
                                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
                    "profiled_step_ms": e0.elapsed_time(e1)}
    note("profiling pass done")
    if world > 1:
        dist.barrier()

    if cpu is not None:
        gpu_ms = _c1_gpu(dev)
        cpu["c1"].update(gpu_ms_median=gpu_ms, gpu_samples_per_s=16 / (gpu_ms * 1e-3),
                         gpu_hbm_frac=2 * 32 * 64 * 300 * 25 * 4 / (gpu_ms * 1e-3) / 1e9 / peak)
    if rank == 0:
        value = batch * world / (ms_step * 1e-3)
        e2e = batch * world / (ms_e2e * 1e-3)
        step_gbs = algo_mb * 1e6 * batch / (ms_step * 1e-3) / 1e9
        line = {
            "metric": ("Shift-GCN fwd+bwd samples/sec (NTU 3x300x25x2)" if args.workload.startswith("ntu") and train
                       else f"Shift-GCN samples/sec ({args.workload})"),
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic",
            "config": {"workload": args.workload, "per_gpu_batch": batch, "global_batch": batch * world,
                       "step": "fwd + CE loss + bwd + flat-buffer NCCL all-reduce (N>1) + SGD-Nesterov" if train else "fwd",
                       "l2": "activation tensors are 246 MB each (> 126 MB L2); no explicit flush needed",
                       "precision": "fp32 storage, TF32 tensor-core contractions (tcgen05 kind::tf32), fp32 accumulate",
                       "parallelism": f"dp{world}", "cuda_graph": bool(graphed)},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(host_x.numel() * 4 + host_y.numel() * 8),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "last_result": last},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_mb_per_sample": algo_mb, "peak_source": peak_src},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        # the captured step holds NCCL work: drop it before the communicator, and do not wait on teardown
        trainer = None
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_ensemble(args):
    """--workload ensemble (BASELINE.json config 5): the 4-stream joint / bone / joint-motion / bone-motion ensemble
    (weights 0.6 / 0.6 / 0.4 / 0.4, ensemble.py:18-27) as batched inference, streams placed on the ranks by
    shiftgcn_b200.ensemble.placement (1 GPU: all four models; 2: two each; 4: one each; 8: two ranks per stream, batch
    halved inside the pair).  Every rank receives the joint batch from pinned host memory, derives its stream on the
    device, and ONE all-reduce of the weighted logits is the ensemble.  A step = one batch of 64 samples through all
    four streams: total work is fixed, so scaling is "strong".  Rank 0 also evaluates all four streams alone and
    reports the difference to the sharded result."""
    import torch
    import torch.distributed as dist

    from shiftgcn_b200 import ensemble as E, ops
    from shiftgcn_b200.modules import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ops.device_check()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N, T, V, M, ncls = 64, 300, 25, 2, 60

    def build(k):
        torch.manual_seed(100 + k)                      # same weights for stream k on whichever rank owns it
        return Model(num_class=ncls, num_point=V, num_person=M, graph="graph.ntu_rgb_d.Graph",
                     graph_args=dict(labeling_mode="spatial")).to(dev).eval()

    def make(models):
        fns = {name: (lambda jb, m=m, name=name: m.forward_stream(jb, name)) for name, m in models.items()}
        return fns

    mine = {E.MODALITIES[k]: build(k) for k, _, _ in E.placement(world, rank)}
    ens = E.StreamEnsemble(make(mine), num_class=ncls, world_size=world, rank=rank, stream_fn=lambda jb, name: jb)
    torch.manual_seed(1)
    host = torch.randn(N, 3, T, V, M).pin_memory()
    dev_x = host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        top = None
        for _ in range(n_steps):
            out = ens.logits(host.to(dev, non_blocking=True) if from_host else dev_x)
            if from_host:
                top = out.argmax(1).cpu()               # D2H read of the step's result
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n_steps, out, top

    for _ in range(max(args.warmup, 3)):
        ens.logits(dev_x)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    ms_dev, _, _ = timed(args.steps, False)
    launches = (ops.LAUNCHES - l0) // args.steps
    timed(2, True)
    ms_e2e, out, top = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        full = {E.MODALITIES[k]: build(k) for k in range(4)} if world > 1 else mine
        ref = E.StreamEnsemble(make(full), num_class=ncls, stream_fn=lambda jb, name: jb).logits(dev_x)
        diff = (ref - out).abs().max().item() / ref.abs().max().item()
        peak, peak_src = _peaks()
        gbs = 4 * 196.0e6 * N / (ms_dev * 1e-3) / 1e9 / world    # per GPU: four streams' inference bytes over `world` GPUs
        emit({"metric": "4-stream ensemble samples/sec (NTU 3x300x25x2, batch 64)", "value": N / (ms_dev * 1e-3),
              "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
              "config": {"workload": "ensemble", "global_batch": N, "streams": list(E.MODALITIES),
                         "weights": list(E.ENSEMBLE_WEIGHTS_DEFAULT),
                         "placement": [E.placement(world, r) for r in range(world)],
                         "l2": "activation tensors are 246 MB each on one GPU (> 126 MB L2)",
                         "collective": "one all-reduce(sum) of the (64, 60) weighted logits"},
              "e2e": {"value": N / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(host.numel() * 4),
                      "d2h_bytes_per_step": N * 8, "ms_per_step": ms_e2e},
              "gpu_launches": int(launches), "clocks": clocks,
              "roofline_step": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "algorithmic_mb_per_sample": 4 * 196.0, "peak_source": peak_src,
                                "note": "per-GPU: 4 streams x 196 MB/sample inference bytes spread over n_gpus"},
              "parity": {"rel_diff_vs_single_rank": diff, "top1_equal": bool((ref.argmax(1).cpu() == top).all())},
              "cpu_baseline": None})
    if world > 1:
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, cuDNN notes) write to fd 1: point it at stderr for the whole run and keep the
    real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["ensemble"],
                    help="default: ntu60-train on one GPU, ntu120-train (config 4) on several")
    ap.add_argument("--ref-batch", type=int, default=4, help="bounded CPU sample (samples per CPU step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="progress notes on stderr")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph")
    args = ap.parse_args()
    args.workload = default_workload(args)
    if args.workload == "ensemble" and args.impl == "b200":
        run_ensemble(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
