#!/bin/bash
# recompile the one-pass kernel objects (cs_fused_inst.cu, 10 variants) and cs_api.o in parallel, relink
set -e
cd "$(dirname "$0")/../cosinesampler_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $EXTRA"
for d in 2 3; do for l in 0 1 2 3 4; do
  nvcc $FLAGS -DCS_DIM=$d -DCS_LSHIFT=$l -c csrc/cs_fused_inst.cu -o build/cs_fused_d${d}_l${l}.o 2> build/cs_fused_d${d}_l${l}.log &
done; done
nvcc $FLAGS -c csrc/cs_api.cu -o build/cs_api.o 2> build/cs_api.ptxas.log &
wait
nvcc -shared -o libcosine_sampler_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a
grep -h "spill" build/cs_fused_d2_l2.log | sort | uniq -c
