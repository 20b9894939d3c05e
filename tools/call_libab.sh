# A/B of library variants on the training bench line (same box):  tools/call_libab.sh "base NAME .."  [workload]
mkdir -p gpurun_out
for v in $1; do
  if [ $v = base ]; then lib=""; else lib="$PWD/shiftgcn_b200/lib/variants/lib$v.so"; fi
  SGCN_LIB=$lib timeout 300 python bench.py --workload ${2:-ntu60-train} --no-cpu-baseline --steps 5 > gpurun_out/libab_$v.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/libab_$v.json").read().strip().splitlines()[-1])
b=d["roofline"]["breakdown_ms"]
print("$v", round(d["ms_per_step"],3), {k:v for k,v in list(b.items())[:8]})
PY
done
