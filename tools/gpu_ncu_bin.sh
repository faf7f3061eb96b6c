#!/bin/bash
python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_bin_scatter -s 1 -c 1 -o gpurun_out/r2_bin_v1 python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > gpurun_out/ncu_bin.log 2>&1
tail -3 gpurun_out/ncu_bin.log
