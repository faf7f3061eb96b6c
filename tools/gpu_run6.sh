#!/bin/bash
# bench: ours + reference arm at N=1, then ncu launch list and full capture of the same commands
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref exit $?"
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/profile_stages.py cfg3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cs_stage_kernel -c 12 -f -o gpurun_out/prof_stages_cfg3 python tools/profile_stages.py cfg3 > gpurun_out/ncu_stages.log 2>&1
echo "ncu stages exit $?"
