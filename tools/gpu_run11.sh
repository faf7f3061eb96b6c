#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail 25 --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for m in never auto; do
COSINE_SAMPLER_SMALL_CELL=$m CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py pixel2d pixel3d > gpurun_out/stage_bench_pixel_$m.jsonl 2> gpurun_out/stage_bench_pixel_$m.err; echo "pixel $m exit $?"
done
