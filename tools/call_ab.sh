# A/B of an environment switch on the three single-GPU bench lines:  tools/call_ab.sh "VAR=a" "VAR=b" [workloads]
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20
i=0
for setting in "$1" "$2"; do
  i=$((i+1))
  for w in ${3:-ntu60-train ntu60-infer}; do
    env $setting timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/ab_${i}_$w.json 2> gpurun_out/ab_${i}_$w.err || tail -5 gpurun_out/ab_${i}_$w.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["config"]["workload"], round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1),
              " step frac", round(d["roofline_step"]["frac"], 4), " kernel", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
