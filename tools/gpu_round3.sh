#!/bin/bash
TAG=${1:-b}
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_reference_chain.py tests/test_gpu_peer_single.py -x -q > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_$TAG.log
timeout 200 python tools/onepass_bench.py cfg3 --points 4194304,33554432 --variants onepass > gpurun_out/r2_onepass_cfg3_$TAG.jsonl 2>&1; cat gpurun_out/r2_onepass_cfg3_$TAG.jsonl
timeout 200 python tools/onepass_bench.py cfg4 --points 4194304 --variants onepass > gpurun_out/r2_onepass_cfg4_$TAG.jsonl 2>&1; cat gpurun_out/r2_onepass_cfg4_$TAG.jsonl
python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_onepass_$TAG.csv python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/ncu_l.log 2>&1
grep -o '"void[^"]*\|"cs_[^"]*\|,"[0-9.]*"$' gpurun_out/r2_launches_onepass_$TAG.csv | paste - - | tail -16
