"""The fused step's gradient reduce over peer memory (peer.PeerReducer) against
the NCCL path: needs at least two GPUs on the box (skipped on the single-GPU test box; run by
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_peer.py`, results in profiles/README.md)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_reduce_equals_nccl_reduce(cuda):
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 1 << (n.bit_length() - 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tools", "peer_test.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["match"] and out["same_bits_on_all_ranks"], out
