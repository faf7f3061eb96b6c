#!/bin/bash
O=gpurun_out
for cfg in cfg3 cfg4; do
  python tools/profile_onepass.py $cfg > $O/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_pde_fused -c 2 -o /tmp/op_$cfg python tools/profile_onepass.py $cfg > $O/ncu_onepass_$cfg.log 2>&1
  ncu -i /tmp/op_$cfg.ncu-rep --page raw --csv > $O/r2_ncu_onepass_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/op_$cfg.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$cfg.csv 2>/dev/null; python tools/ncu_lines.py /tmp/src_$cfg.csv 70 > $O/r2_ncu_onepass_${cfg}_lines.txt 2>&1
done
# red segments at the bench's density: one launch over 2^25 binned points
python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > /dev/null 2>&1 && ncu --metrics lts__t_sectors_srcunit_tex_op_red.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cs_pde_fused -c 1 --csv --log-file $O/r2_ncu_onepass_2p25.csv python tools/onepass_bench.py cfg3 --points 33554432 --variants onepass --once > /dev/null 2>&1
cat $O/r2_ncu_onepass_2p25.csv | tail -8
python tools/profile_onepass.py cfg3 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"cs_head_postmix|cs_head_premix" -c 2 -o /tmp/mix python tools/profile_onepass.py cfg3 > $O/ncu_postmix.log 2>&1
ncu -i /tmp/mix.ncu-rep --page raw --csv > $O/r2_ncu_mix_raw.csv 2>/dev/null
du -sh $O
