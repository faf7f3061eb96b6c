#!/bin/bash
# the ncu evidence of round 2 (profiles/r2/): one-pass kernel at both configs, the stage kernels with the chain's
# expanded gOut, the post-mix kernel, the launch list of the default bench.  Reports are exported to CSV on the
# box and deleted (gpurun_out/ is limited to 64 MiB).
O=gpurun_out
for cfg in cfg3 cfg4; do
  python tools/profile_onepass.py $cfg > $O/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_pde_fused -c 2 -o /tmp/op_$cfg python tools/profile_onepass.py $cfg > $O/ncu_onepass_$cfg.log 2>&1
  ncu -i /tmp/op_$cfg.ncu-rep --page raw --csv > $O/r2_ncu_onepass_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/op_$cfg.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$cfg.csv 2>/dev/null; python tools/ncu_lines.py /tmp/src_$cfg.csv 70 > $O/r2_ncu_onepass_${cfg}_lines.txt 2>&1
  python tools/profile_stages.py $cfg > $O/plain_st_$cfg.log 2>&1 && ncu --set full --clock-control none -k regex:cs_stage_kernel -c 12 -o /tmp/st_$cfg python tools/profile_stages.py $cfg > $O/ncu_stages_$cfg.log 2>&1
  ncu -i /tmp/st_$cfg.ncu-rep --page raw --csv > $O/r2_ncu_stages_${cfg}_raw.csv 2>/dev/null
done
python tools/profile_onepass.py cfg3 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"cs_head_postmix|cs_head_premix|cs_bin_count" -c 6 -o /tmp/mix python tools/profile_onepass.py cfg3 > $O/ncu_postmix.log 2>&1
ncu -i /tmp/mix.ncu-rep --page raw --csv > $O/r2_ncu_mix_raw.csv 2>/dev/null
python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline --no-extras > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_ncu_launches_bench.csv python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
timeout 600 python -m pytest tests/test_gpu_f64.py tests/test_gpu_onepass.py -x -q > $O/r2_pytest_f64.log 2>&1; tail -5 $O/r2_pytest_f64.log
du -sh $O; ls -la $O | head -30
