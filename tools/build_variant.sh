#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..": every csrc/*.cu rebuilt with extra flags into
# shiftgcn_b200/lib/variants/libNAME.so (select it with SGCN_LIB=... for A/B runs of tools/kernel_bench.py / bench.py)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/.."
d=shiftgcn_b200/lib/variants/obj_$name
mkdir -p $d
for f in shiftgcn_b200/csrc/*.cu; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I include $flags \
    -c $f -o $d/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o shiftgcn_b200/lib/variants/lib$name.so $d/*.o -lcudart
rm -rf $d
echo built shiftgcn_b200/lib/variants/lib$name.so
