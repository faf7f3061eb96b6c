#!/bin/bash
# the bench at N GPUs, launched as the driver launches it; the multi-GPU peer-reduce test on real peers.  usage: tools/gpu_multi_r3.sh N
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_dp_gloo.py -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r3_bench_n${N}.json 2> gpurun_out/r3_bench_n${N}.err; tail -3 gpurun_out/r3_bench_n${N}.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r3_bench_n${N}.json") if l.startswith("{")][-1])
    f=d["fused"]
    print("dropin", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("h2d_GBps_alone"))
    print("fused", f["value"], "e2e", f["e2e"]["value"], f["gradient_reduce"], f["ms_per_step"], f["stages"])
except Exception as e:
    print("no bench line", e)
PY
