# trimmed end-of-round evidence (last visit of round 2): parity tests, smoke, three single-GPU bench lines, launch list of one
# training step, ncu --set full of the two first-layer backward kernels
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
timeout 300 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 120 python bench.py --workload ntu60-infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 120 python bench.py --workload mediapipe-train --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_step.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:stem_bwd" --launch-skip 4 -c 2 -f -o gpurun_out/ncu_stem_final \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_stem_final.log 2>&1
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench", "bench_infer", "bench_mp"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("config", {}).get("workload"), round(d["value"], 1), d["unit"], round(d.get("ms_per_step", 0), 3), "ms  e2e", round(d["e2e"]["value"], 1),
              " step frac", round(d.get("roofline_step", {}).get("frac", 0), 4), " kernel", d.get("roofline", {}).get("kernel"), round(d.get("roofline", {}).get("frac", 0), 3))
    except Exception as e:
        print(f, "ERR", e)
PY
