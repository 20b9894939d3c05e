#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..": pointwise.cu rebuilt with extra flags, linked with the stock objects into
# shiftgcn_b200/lib/variants/libNAME.so (select it with SGCN_LIB=... for A/B runs of tools/kernel_bench.py)
set -e
name=$1; flags=$2
cd "$(dirname "$0")/.."
mkdir -p shiftgcn_b200/lib/variants
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I include $flags \
  -c shiftgcn_b200/csrc/pointwise.cu -o shiftgcn_b200/lib/variants/pointwise_$name.o
objs=$(ls shiftgcn_b200/lib/obj/*.o | grep -v pointwise.o)
nvcc -shared -o shiftgcn_b200/lib/variants/lib$name.so $objs shiftgcn_b200/lib/variants/pointwise_$name.o -lcudart
echo built shiftgcn_b200/lib/variants/lib$name.so
