#!/bin/bash
# ncu --set full of one launch of the PDE head kernel (tools/head_bench.py): details page, raw page and the
# SASS source page (stall samples / executed instructions per instruction) as CSV under gpurun_out/.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cs_pde_head -s 6 -c 1 -f -o gpurun_out/prof_head python tools/head_bench.py > gpurun_out/ncu_head.log 2>&1
ncu -i gpurun_out/prof_head.ncu-rep --page raw --csv > gpurun_out/ncu_head_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_head.ncu-rep --page source --print-source sass --csv > gpurun_out/ncu_head_source.csv 2>/dev/null
ncu -i gpurun_out/prof_head.ncu-rep --page details > gpurun_out/ncu_head_details.txt 2>/dev/null
rm -f gpurun_out/prof_head.ncu-rep
