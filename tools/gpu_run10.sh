#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
for mb in 100000 80 48 8; do
  CS_CELL_MAJOR_MB=$mb CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py cfg4 > gpurun_out/stage_bench_cm$mb.jsonl 2> gpurun_out/stage_bench_cm$mb.err; echo "cm $mb exit $?"
done
CS_CELL_MAJOR_MB=8 CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py cfg3 > gpurun_out/stage_bench_cfg3_cm8.jsonl 2>&1
