# ncu --set full of the 9 spatial-backward launches of one training step (DRAM traffic per launch of the dominant kernel)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:spatial_bwd_kernel -s 27 -c 9 -f \
  -o gpurun_out/ncu_spatial_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_traffic.log 2>&1
ncu -i gpurun_out/ncu_spatial_bwd.ncu-rep --page raw --csv > gpurun_out/raw_spatial_bwd.csv 2>/dev/null; rm -f gpurun_out/ncu_spatial_bwd.ncu-rep
tail -3 gpurun_out/ncu_traffic.log; ls -la gpurun_out/raw_spatial_bwd.csv
