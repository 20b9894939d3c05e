mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
timeout 400 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 200 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1
timeout 300 python bench.py --workload ntu60-infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 300 python bench.py --workload mediapipe-train --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -40; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_infer.json gpurun_out/bench_mp.json
