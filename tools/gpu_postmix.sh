#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_onepass.py -x -q 2>&1 | tail -2
for m in 2 4 8; do echo "== CS_POSTMIX_BLOCKS=$m"; for cfg in cfg3 cfg4; do CS_POSTMIX_BLOCKS=$m timeout 100 python tools/onepass_bench.py $cfg --points 4194304 --variants onepass 2>&1 | grep "^{" | python -c "import sys,json; [print(d['workload'], d['stages_ms']['POSTMIX'], d['stages_ms']['PREMIX']) for d in map(json.loads, sys.stdin)]"; done; done
