#!/bin/bash
# one GPU round for the one-pass step: parity tests, timings, one ncu capture.  usage: tools/gpu_onepass_round.sh TAG
TAG=${1:-vX}
timeout 600 python -m pytest tests/test_gpu_onepass.py -q -x > gpurun_out/r2_pytest_onepass.log 2>&1; tail -5 gpurun_out/r2_pytest_onepass.log
timeout 300 python tools/onepass_bench.py cfg3 --points 1048576,4194304,33554432 --variants onepass,onepass_noagg,onepass_unbinned > gpurun_out/r2_onepass_cfg3_$TAG.jsonl 2> gpurun_out/r2_onepass_cfg3_$TAG.err
cat gpurun_out/r2_onepass_cfg3_$TAG.jsonl; tail -5 gpurun_out/r2_onepass_cfg3_$TAG.err
timeout 300 python tools/onepass_bench.py cfg4 --points 4194304 --variants onepass,onepass_noagg,onepass_unbinned > gpurun_out/r2_onepass_cfg4_$TAG.jsonl 2> gpurun_out/r2_onepass_cfg4_$TAG.err
cat gpurun_out/r2_onepass_cfg4_$TAG.jsonl
python tools/onepass_bench.py cfg3 --points 4194304 --variants onepass --once > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_pde_fused -c 1 -o gpurun_out/r2_onepass_$TAG python tools/onepass_bench.py cfg3 --points 4194304 --variants onepass --once > gpurun_out/ncu_onepass_$TAG.log 2>&1
tail -3 gpurun_out/ncu_onepass_$TAG.log
