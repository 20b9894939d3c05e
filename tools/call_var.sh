# A/B of variant libraries: tools/call_var.sh "var1 var2 ..." "kernel-filter" [layers]
mkdir -p gpurun_out; : > gpurun_out/variants.txt
for v in base $1; do
  if [ $v = base ]; then lib=""; else lib="$PWD/shiftgcn_b200/lib/variants/lib$v.so"; fi
  for k in $2; do
    SGCN_LIB=$lib timeout 120 python tools/kernel_bench.py --only "$k" --layers ${3:-64,128,256} --reps 5 2>&1 | sed "s/^/$v | /" >> gpurun_out/variants.txt
  done
done
cat gpurun_out/variants.txt
