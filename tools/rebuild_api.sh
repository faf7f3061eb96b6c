#!/bin/bash
# recompile only cs_api.o (bin / mix / peer kernels, C ABI) and relink: the other units are untouched
set -e
cd "$(dirname "$0")/../cosinesampler_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -c csrc/cs_api.cu -o build/cs_api.o 2> build/cs_api.ptxas.log
nvcc -shared -o libcosine_sampler_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a
touch libcosine_sampler_b200.so
