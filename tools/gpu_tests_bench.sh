#!/bin/bash
# full parity suite + the default bench line (run under gpurun); usage: gpu_tests_bench.sh TAG
TAG=${1:-x}
timeout 1500 python -m pytest tests -m gpu -q --maxfail 10 --timeout 600 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_n1.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n1.json"))
print("dropin", d["value"], d["e2e"], "fused", d["fused_value"], d["fused"]["e2e"], "cfg4 fused", d["cfg4"]["fused_value"], d["cfg4"]["fused"]["e2e"]["value"])
print(d["fused"]["stages"]); print(d["fused"]["roofline"]["issue_roofline"])
PY
