#!/bin/bash
# Short GPU pass on the final tree (run under gpurun): parity tests, smoke, both bench workloads and the
# ncu --set full captures of the three fused kernels (exported to CSV; the reports are deleted).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 25 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
for cfg in cfg3 cfg4; do
python tools/profile_fused.py $cfg > gpurun_out/profile_fused_plain_$cfg.log 2>&1 &&
ncu --set full --clock-control none -k regex:'cs_jet|cs_pde_head' -c 6 -f -o gpurun_out/prof_fused_$cfg python tools/profile_fused.py $cfg > gpurun_out/ncu_fused_$cfg.log 2>&1
echo "ncu fused $cfg exit $?"
ncu -i gpurun_out/prof_fused_$cfg.ncu-rep --page raw --csv > gpurun_out/ncu_fused_${cfg}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_fused_$cfg.ncu-rep
done
