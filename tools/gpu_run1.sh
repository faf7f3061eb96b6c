#!/bin/bash
# First GPU pass: parity tests, smoke, micro-benchmarks, stage timings, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
echo "== microbench"
timeout 300 ./tools/microbench > gpurun_out/microbench.jsonl 2>&1; echo "microbench exit $?"
echo "== stage bench"
timeout 600 python tools/stage_bench.py > gpurun_out/stage_bench.jsonl 2> gpurun_out/stage_bench.err; echo "stage_bench exit $?"
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-sample 262144 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench.json
