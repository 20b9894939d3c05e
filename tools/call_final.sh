# end-of-round evidence: parity tests, smoke, the three single-GPU bench lines, per-kernel bench, launch lists
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
timeout 400 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --workload ntu60-infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 300 python bench.py --workload mediapipe-train --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1
if [ -z "$NO_LISTS" ]; then
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b_nograph.json 2>gpurun_out/b_nograph.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_step.log 2>&1
fi
if [ -n "$INFER_LIST" ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_infer.csv \
  python bench.py --workload ntu60-infer --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_infer.log 2>&1
fi
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench", "bench_infer", "bench_mp", "bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("config", {}).get("workload"), round(d["value"], 1), d["unit"], round(d.get("ms_per_step", 0), 3), "ms  e2e", round(d["e2e"]["value"], 1),
              " step frac", round(d.get("roofline_step", {}).get("frac", 0), 4), " kernel", d.get("roofline", {}).get("kernel"), round(d.get("roofline", {}).get("frac", 0), 3))
    except Exception as e:
        print(f, "ERR", e)
PY
