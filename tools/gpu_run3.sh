#!/bin/bash
mkdir -p gpurun_out
python tools/profile_stages.py cfg3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cs_stage_kernel -c 12 -f -o gpurun_out/prof_stages_cfg3 python tools/profile_stages.py cfg3 > gpurun_out/ncu_stages.log 2>&1
echo "ncu stages exit $?"
tail -3 gpurun_out/ncu_stages.log
