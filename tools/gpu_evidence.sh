#!/bin/bash
# The GPU evidence pass of a round (run under gpurun): parity tests, smoke, stage timings vs the
# reference CUDA op, bench (device-resident + e2e), ncu launch list and full captures.  ncu reports
# are exported to CSV on the box and deleted (gpurun_out/ may not exceed 64 MiB).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python tools/stage_bench.py > gpurun_out/stage_bench.jsonl 2> gpurun_out/stage_bench.err; echo "stage bench exit $?"
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/profile_stages.py cfg3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cs_stage_kernel -c 12 -f -o gpurun_out/prof_stages_cfg3 python tools/profile_stages.py cfg3 > gpurun_out/ncu_stages.log 2>&1
echo "ncu stages exit $?"
ncu -i gpurun_out/prof_stages_cfg3.ncu-rep --page raw --csv > gpurun_out/ncu_stages_cfg3_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_stages_cfg3.ncu-rep --page source --print-source cuda,sass --csv --kernel-id ::regex:cs_stage_kernel:4 > gpurun_out/ncu_source_B2dG.csv 2>/dev/null
rm -f gpurun_out/prof_stages_cfg3.ncu-rep
python tools/profile_stages.py cfg4 > gpurun_out/profile_plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:cs_stage_kernel -c 12 -f -o gpurun_out/prof_stages_cfg4 python tools/profile_stages.py cfg4 > gpurun_out/ncu_stages4.log 2>&1
echo "ncu stages cfg4 exit $?"
ncu -i gpurun_out/prof_stages_cfg4.ncu-rep --page raw --csv > gpurun_out/ncu_stages_cfg4_raw.csv 2>/dev/null
rm -f gpurun_out/prof_stages_cfg4.ncu-rep
du -sh gpurun_out
