# quick GPU visit: parity tests + selected kernel benches.  tools/call_quick.sh "kernel filters" [layers] [pytest args]
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x ${3:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-200 | head -20
: > gpurun_out/kb.txt
for k in $1; do timeout 120 python tools/kernel_bench.py --only "$k" --layers ${2:-64,128,256} --reps 5 >> gpurun_out/kb.txt 2>&1; done
cat gpurun_out/kb.txt
