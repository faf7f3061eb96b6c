#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
for v in "" _mb1pg1 _mb1pg2 _mb2pg1 _mb2pg2 _mb3pg1; do
  echo "== stage bench variant '$v'"
  COSINE_SAMPLER_LIB=$PWD/cosinesampler_b200/libcosine_sampler_b200$v.so CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py > gpurun_out/stage_bench$v.jsonl 2> gpurun_out/stage_bench$v.err; echo "exit $?"
done
