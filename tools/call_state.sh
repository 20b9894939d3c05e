# state of the tree: GPU parity tests, smoke, the three single-GPU bench lines, per-kernel bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
timeout 400 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --workload ntu60-infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 300 python bench.py --workload mediapipe-train --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err
timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench", "bench_infer", "bench_mp"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["config"]["workload"], round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1),
              " step frac", round(d["roofline_step"]["frac"], 4), " kernel", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
        print("   ", {k: v for k, v in list(d["roofline"]["breakdown_ms"].items())[:16]})
    except Exception as e:
        print(f, "ERR", e)
PY
