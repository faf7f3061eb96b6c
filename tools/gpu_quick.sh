#!/bin/bash
timeout 300 python tools/onepass_errors.py > gpurun_out/r2_onepass_errors.jsonl 2>&1; cat gpurun_out/r2_onepass_errors.jsonl | cut -c1-600
timeout 900 python bench.py > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; tail -2 gpurun_out/r2_bench_h.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_h.json"))
print(d["value"], d["fused_value"], d["fused_e2e_value"], d["fused"]["roofline"]["issue_roofline"], d["cfg4"]["fused"]["roofline"]["issue_roofline"], d["cfg4"]["fused_value"])
PY
