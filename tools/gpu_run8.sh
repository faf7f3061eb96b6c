#!/bin/bash
mkdir -p gpurun_out
for v in "" _mb2 _pg2 _pg2mb3 _mb4; do
  COSINE_SAMPLER_LIB=$PWD/cosinesampler_b200/libcosine_sampler_b200$v.so CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py > gpurun_out/stage_bench$v.jsonl 2> gpurun_out/stage_bench$v.err; echo "variant '$v' exit $?"
done
