# tools/call_ncu.sh "name:regex[:layer] ..."  -- one full ncu capture per spec of a tools/kernel_bench.py kernel
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  IFS=: read name rx layer <<< "$spec"; i=$((i+1))
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip 1 -c 1 -f \
    -o gpurun_out/ncu_$i python tools/kernel_bench.py --only "$name" --layers ${layer:-64} --reps 1 > gpurun_out/ncu_$i.log 2>&1
  tail -2 gpurun_out/ncu_$i.log
done
ls -la gpurun_out/*.ncu-rep
