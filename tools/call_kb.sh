# quick GPU visit: parity tests (-x), selected kernel benches, optional training bench.  tools/call_kb.sh "kernel filters" [layers] [bench:0/1]
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20
: > gpurun_out/kb.txt
for k in $1; do timeout 120 python tools/kernel_bench.py --only "$k" --layers ${2:-64,128,256} --reps 5 >> gpurun_out/kb.txt 2>&1; done
cut -c1-120 gpurun_out/kb.txt
if [ "${3:-1}" = 1 ]; then
  timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  timeout 300 python bench.py --workload ntu60-infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
  python - <<'PY'
import json
for f in ("bench", "bench_infer"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["config"]["workload"], round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1),
              " step frac", round(d["roofline_step"]["frac"], 4), " kernel", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
        bf = d["roofline"].get("breakdown_frac", {})
        print("   ", {k: (v, bf.get(k)) for k, v in list(d["roofline"]["breakdown_ms"].items())[:18]})
    except Exception as e:
        print(f, "ERR", e)
PY
fi
