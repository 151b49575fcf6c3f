#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : liboriana_b200 with extra nvcc flags for kernels_tc.cu (the other objects are
# the product build's, oriana_b200/lib/*.o) -> oriana_b200/lib/variants/NAME.so
# (kernel A/B experiments on the GPU box: ORIANA_B200_LIB=oriana_b200/lib/variants/NAME.so python ...)
set -e
cd "$(dirname "$0")/../oriana_b200"
mkdir -p lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $2 \
     -c csrc/kernels_tc.cu -o lib/variants/$1.kernels_tc.o
nvcc -shared -o lib/variants/$1.so lib/api.o lib/kernels_simt.o lib/synth.o lib/variants/$1.kernels_tc.o -lcudart
rm -f lib/variants/$1.kernels_tc.o
echo lib/variants/$1.so
