#!/bin/bash
# link a variant of the library that differs only in the fused-step kernels: tools/build_fused_variant.sh NAME -DFOO=1 ...
NAME=$1; shift
B=cosinesampler_b200/build; V=cosinesampler_b200/build_var_$NAME; mkdir -p $V
for d in 2 3; do for l in 0 1 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -DCS_DIM=$d -DCS_LSHIFT=$l -c cosinesampler_b200/csrc/cs_fused_inst.cu -o $V/cs_fused_d${d}_l${l}.o &
done; done; wait
OBJS=$(ls $B/*.o | grep -v cs_fused_d); nvcc -shared -o cosinesampler_b200/libcosine_sampler_b200_$NAME.so $OBJS $V/*.o -gencode arch=compute_100a,code=sm_100a && echo built $NAME
