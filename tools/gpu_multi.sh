#!/bin/bash
# multi-GPU round: peer reduce variants vs NCCL, then the bench at N GPUs.  usage: tools/gpu_multi.sh N TAG
N=${1:-2}; TAG=${2:-a}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/peer_test.py > gpurun_out/r2_peer_test_n${N}_$TAG.log 2>&1; grep "^{" gpurun_out/r2_peer_test_n${N}_$TAG.log | tail -1; tail -5 gpurun_out/r2_peer_test_n${N}_$TAG.log | grep -v "^{" | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n${N}_$TAG.json 2> gpurun_out/r2_bench_n${N}_$TAG.err; tail -3 gpurun_out/r2_bench_n${N}_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n${N}_$TAG.json") if l.startswith("{")][-1])
    f=d["fused"]
    print("dropin", d["value"], "e2e", d["e2e"]["value"])
    print("fused", f["value"], "e2e", f["e2e"]["value"], f["gradient_reduce"], f["ms_per_step"], f["stages"])
except Exception as e:
    print("no bench line", e)
PY
