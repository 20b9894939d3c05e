# per-kernel A/B of library variants (tools/build_variant.sh):  tools/call_variants.sh "variants" "kernel filters" [layers]
mkdir -p gpurun_out; : > gpurun_out/variants.txt
for v in $1; do
  if [ $v = base ]; then lib=""; else lib="$PWD/shiftgcn_b200/lib/variants/lib$v.so"; fi
  for k in $2; do
    SGCN_LIB=$lib timeout 120 python tools/kernel_bench.py --only "$k" --layers ${3:-64,128,256} --reps 5 2>&1 | sed "s/^/$v | /" >> gpurun_out/variants.txt
  done
done
sort -t'|' -k2,2 -s gpurun_out/variants.txt | cut -c1-120
