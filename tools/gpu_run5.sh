#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py > gpurun_out/stage_bench.jsonl 2> gpurun_out/stage_bench.err; echo "stage bench exit $?"
