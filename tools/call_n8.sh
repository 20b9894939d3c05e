# 8-GPU visit (charged 8x): NCCL parity test on two ranks, data-parallel training and the 4-stream ensemble on 8 ranks
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_nccl.py -q > gpurun_out/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl.log)
tail -4 gpurun_out/pytest_nccl.log
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --workload ensemble --steps 10 --warmup 3 > gpurun_out/bench_ensemble_n$n.json 2> gpurun_out/bench_ensemble_n$n.err
  cat gpurun_out/bench_ensemble_n$n.json | cut -c1-600
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_train_n8.json 2> gpurun_out/bench_train_n8.err
cat gpurun_out/bench_train_n8.json | cut -c1-700
