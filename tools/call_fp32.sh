mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fp32.py -q > gpurun_out/pytest_fp32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fp32.log)
grep -E "^E  +(Assertion|assert)|FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_fp32.log | cut -c1-200 | head -80
