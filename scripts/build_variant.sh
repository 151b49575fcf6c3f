#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : liboriana_b200 with extra nvcc flags -> oriana_b200/lib/variants/NAME.so
# (kernel A/B experiments on the GPU box: ORIANA_B200_LIB=oriana_b200/lib/variants/NAME.so python ...)
set -e
cd "$(dirname "$0")/../oriana_b200"
mkdir -p lib/variants/obj_$1
for f in api kernels_simt synth kernels_tc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c csrc/$f.cu -o lib/variants/obj_$1/$f.o &
done
wait
nvcc -shared -o lib/variants/$1.so lib/variants/obj_$1/*.o -lcudart
rm -rf lib/variants/obj_$1
echo lib/variants/$1.so
