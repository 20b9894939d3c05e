for v in base mb3 un8 un2 g16 g4 w16; do
  if [ $v = base ]; then lib=""; else lib="$PWD/shiftgcn_b200/lib/variants/lib$v.so"; fi
  for k in tshift bn_res; do
    SGCN_LIB=$lib timeout 120 python tools/kernel_bench.py --only $k --layers 64,256 --reps 5 2>&1 | sed "s/^/$v | /" >> gpurun_out/variants.txt
  done
done
