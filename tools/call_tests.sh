mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -60
