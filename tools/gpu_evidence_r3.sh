#!/bin/bash
# The GPU evidence pass of round 2's second session (files land under gpurun_out/r3_*; copy the ones to keep into
# profiles/r2_session2/): ncu --set full of the one-pass kernel at the bench's launch sizes (summaries are written on the
# box too, so that the bench run below reads the instruction counts of THIS build), parity suite, smoke, the default
# bench line, its launch list, and the measured errors of the one-pass step.  ncu reports are exported and deleted.
O=gpurun_out; mkdir -p $O profiles/r2_session2
for spec in "cfg3 33554432" "cfg4 4194304"; do
  set -- $spec; cfg=$1; pts=$2
  python tools/profile_onepass.py $cfg $pts > $O/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cs_pde_fused -c 2 -o /tmp/op_$cfg python tools/profile_onepass.py $cfg $pts > $O/ncu_onepass_$cfg.log 2>&1
  echo "ncu onepass $cfg exit $?"
  ncu -i /tmp/op_$cfg.ncu-rep --page raw --csv > $O/r3_ncu_onepass_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/op_$cfg.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$cfg.csv 2>/dev/null; python tools/ncu_lines.py /tmp/src_$cfg.csv 70 > $O/r3_ncu_onepass_${cfg}_lines.txt 2>&1
  python tools/summarize_ncu.py $O/r3_ncu_onepass_${cfg}_raw.csv $O/r3_ncu_onepass_${cfg}_summary.json onepass_$cfg $pts
  cp $O/r3_ncu_onepass_${cfg}_summary.json profiles/r2_session2/ncu_onepass_${cfg}_summary.json
  python tools/ncu_keys.py $O/r3_ncu_onepass_${cfg}_raw.csv | tail -46 > $O/r3_ncu_onepass_${cfg}_keys.txt
done
timeout 1500 python -m pytest tests -m gpu -q --maxfail 10 --timeout 600 > $O/r3_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $O/r3_bench_n1.json 2> $O/r3_bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r3_bench_reference.json 2> $O/r3_bench_reference.err; echo "reference arm exit $?"
python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline --no-extras > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r3_ncu_launches_bench.csv python bench.py --steps 1 --warmup 1 --points 2097152 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1; echo "launch list exit $?"
timeout 300 python tools/onepass_errors.py > $O/r3_onepass_errors.jsonl 2>&1; echo "errors exit $?"
timeout 300 python tools/onepass_bench.py cfg3 --points 1048576,4194304,33554432 > $O/r3_onepass_bench.jsonl 2>/dev/null; timeout 300 python tools/onepass_bench.py cfg4 --points 1048576,4194304 >> $O/r3_onepass_bench.jsonl 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/r3_bench_n1.json"))
print("dropin", d["value"], d["e2e"]["value"], "fused", d["fused_value"], d["fused_e2e_value"], "cfg4", d["cfg4"]["value"], d["cfg4"]["fused_value"], d["cfg4"]["fused_e2e_value"])
print(d["fused"]["roofline"]["issue_roofline"]); print(d["cfg4"]["fused"]["roofline"]["issue_roofline"])
PY
du -sh $O
