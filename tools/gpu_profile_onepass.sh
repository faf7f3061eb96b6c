#!/bin/bash
# ncu --set full of the one-pass kernel (2^22 binned points): counters + per-line shares; usage: [POINTS=n] gpu_profile_onepass.sh TAG [cfg3 cfg4]
O=gpurun_out; TAG=$1; shift
for cfg in "${@:-cfg3}"; do
  python tools/profile_onepass.py $cfg $POINTS > $O/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_pde_fused -c 2 -o /tmp/op_$cfg python tools/profile_onepass.py $cfg $POINTS > $O/ncu_onepass_$cfg.log 2>&1
  ncu -i /tmp/op_$cfg.ncu-rep --page raw --csv > $O/${TAG}_ncu_onepass_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/op_$cfg.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$cfg.csv 2>/dev/null; python tools/ncu_lines.py /tmp/src_$cfg.csv 90 > $O/${TAG}_ncu_onepass_${cfg}_lines.txt 2>&1
  python tools/ncu_keys.py $O/${TAG}_ncu_onepass_${cfg}_raw.csv | tail -46
done
