#!/bin/bash
# full GPU round: whole -m gpu suite, default bench, kernel variant timing, launch list of one one-pass step
TAG=${1:-a}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; tail -3 gpurun_out/r2_bench_$TAG.err; head -c 3000 gpurun_out/r2_bench_$TAG.json; echo
timeout 200 python tools/onepass_bench.py cfg3 --points 4194304,33554432 --variants onepass > gpurun_out/r2_onepass_cfg3_$TAG.jsonl 2>&1; cat gpurun_out/r2_onepass_cfg3_$TAG.jsonl
python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_onepass_$TAG.csv python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/ncu_l.log 2>&1
grep -o '"void[^"]*\|"cs_[^"]*\|,"[0-9.]*"$' gpurun_out/r2_launches_onepass_$TAG.csv | paste - - | tail -12
