#!/bin/bash
# one GPU visit: parity tests, smoke, bench, per-kernel bench, launch list, full ncu captures of named kernels
#   tools/gpu_round.sh [kernel-bench-name:ncu-kernel-regex ...]
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
timeout 300 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 200 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1
i=0
for spec in "$@"; do
  name="${spec%%:*}"; rx="${spec##*:}"; i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip 1 -c 1 -f \
    -o gpurun_out/ncu_$i python tools/kernel_bench.py --only "$name" --layers ${NCU_LAYER:-64} --reps 1 > gpurun_out/ncu_$i.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json
