#!/bin/bash
# time fused-kernel build variants: tools/gpu_variants.sh name1 name2 ...
for v in "" "$@"; do
  lib=$PWD/cosinesampler_b200/libcosine_sampler_b200${v:+_$v}.so
  echo "== variant '${v:-default}'"
  COSINE_SAMPLER_LIB=$lib timeout 200 python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass 2>&1 | grep "^{" | python -c "import sys,json; [print(d['points'], d['ms_per_step'], d['stages_ms']) for d in map(json.loads, sys.stdin)]"
  COSINE_SAMPLER_LIB=$lib timeout 200 python tools/onepass_bench.py cfg4 --points 4194304 --variants onepass 2>&1 | grep "^{" | python -c "import sys,json; [print(d['points'], d['ms_per_step'], d['stages_ms']) for d in map(json.loads, sys.stdin)]"
done
