#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
CS_SKIP_REF=1 timeout 600 python tools/stage_bench.py > gpurun_out/stage_bench.jsonl 2> gpurun_out/stage_bench.err; echo "stage bench exit $?"
python tools/profile_stages.py cfg3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:cs_stage_kernel -c 12 --csv --log-file gpurun_out/ncu_quick.csv python tools/profile_stages.py cfg3 > gpurun_out/ncu_quick.log 2>&1
echo "ncu exit $?"
