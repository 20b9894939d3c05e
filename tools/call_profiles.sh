# round-2 evidence: launch lists of one training / inference step, DRAM traffic of the dominant kernels, kernel bench.
# ncu reports are converted to CSV on the box and removed (gpurun_out/ must stay under 64 MiB).
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_infer.csv \
  python bench.py --workload ntu60-infer --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_infer.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:spatial_bwd_kernel -s 27 -c 9 -f \
  -o gpurun_out/ncu_spatial_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_traffic.log 2>&1
ncu -i gpurun_out/ncu_spatial_bwd.ncu-rep --page raw --csv > gpurun_out/raw_spatial_bwd.csv 2>/dev/null; rm -f gpurun_out/ncu_spatial_bwd.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:fused_gemm_kernel -s 60 -c 20 -f \
  -o gpurun_out/ncu_infer_gemm python bench.py --workload ntu60-infer --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_traffic_infer.log 2>&1
ncu -i gpurun_out/ncu_infer_gemm.ncu-rep --page raw --csv > gpurun_out/raw_infer_gemm.csv 2>/dev/null; rm -f gpurun_out/ncu_infer_gemm.ncu-rep
timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out/*.csv
