mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_nccl.py -q > gpurun_out/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl.log)
tail -3 gpurun_out/pytest_nccl.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 700 gpurun_out/bench_n2.json | head -c 400; echo
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), d["e2e"]["value"])
PY
