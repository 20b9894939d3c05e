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
# optional extras, selected by environment variables
if [ -n "$EXTRA_WORKLOADS" ]; then
  for w in $EXTRA_WORKLOADS; do
    timeout 300 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  done
fi
if [ -n "$LAUNCH_LIST" ]; then
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b_nograph.json 2>gpurun_out/b_nograph.err && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
fi
if [ -n "$TRAFFIC" ]; then   # ncu --set full of the 9 spatial-backward launches of one training step
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spatial_bwd_kernel -s 27 -c 9 -f \
    -o gpurun_out/ncu_spatial_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_traffic.log 2>&1
fi
