#!/bin/bash
mkdir -p gpurun_out
for cfg in cfg3 cfg4; do
python tools/profile_fused.py $cfg > gpurun_out/profile_fused_plain_$cfg.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'cs_jet|cs_pde_head' -c 6 -f -o gpurun_out/prof_fused_$cfg python tools/profile_fused.py $cfg > gpurun_out/ncu_fused_$cfg.log 2>&1
echo "ncu fused $cfg exit $?"
ncu -i gpurun_out/prof_fused_$cfg.ncu-rep --page raw --csv > gpurun_out/ncu_fused_${cfg}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_fused_cfg3.ncu-rep --page source --print-source sass --csv --kernel-id ::regex:cs_pde_head:2 > gpurun_out/ncu_source_head2d.csv 2>/dev/null
ncu -i gpurun_out/prof_fused_cfg3.ncu-rep --page source --print-source sass --csv --kernel-id ::regex:cs_jet_fwd:2 > gpurun_out/ncu_source_jetfwd2d.csv 2>/dev/null
rm -f gpurun_out/prof_fused_cfg3.ncu-rep gpurun_out/prof_fused_cfg4.ncu-rep
python bench.py --steps 3 --warmup 3 --points 8388608 --no-cpu-baseline > gpurun_out/bench_fused_try.json 2> gpurun_out/bench_fused_try.err; echo "bench exit $?"
du -sh gpurun_out
