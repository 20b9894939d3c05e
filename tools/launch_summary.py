"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) for ONE step of bench.py.

    python tools/launch_summary.py gpurun_out/launches.csv [step_marker_substring]

The step is delimited by two consecutive occurrences of the marker kernel (default: nll_loss_forward, the CE loss)."""
import collections, csv, re, sys
path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "nll_loss_forward"
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
names = [re.sub(r"\(.*", "", r["Kernel Name"])[:110] for r in rows]
t = [float(r["Metric Value"]) / 1e3 for r in rows]
idx = [i for i, n in enumerate(names) if marker in n]
a, b = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(rows))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, x in zip(names[a:b], t[a:b]):
    agg[n][0] += 1
    agg[n][1] += x
tot = sum(v[1] for v in agg.values())
OURS = ("sgcn::", "sb::", "fg::", "stem::", "side::")   # ncu prints the innermost namespace of libshiftgcn_b200.so kernels
ours = sum(v[1] for k, v in agg.items() if any(t in k for t in OURS))
n_ours = sum(v[0] for k, v in agg.items() if any(t in k for t in OURS))
print(f"# {b - a} launches in one step, {tot / 1e3:.2f} ms serialised; libshiftgcn_b200 kernels: {n_ours} launches, {ours / 1e3:.2f} ms ({ours / tot * 100:.1f}%)")
for n, (c, x) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 70]:
    print(f"{x / tot * 100:5.1f}% {x / 1e3:8.2f} ms n={c:3d} avg={x / c:8.1f} us  {n}")
