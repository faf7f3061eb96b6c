"""Time cs_pde_head_step alone (2D and 3D, C = 16, 2^20 points); COSINE_SAMPLER_LIB selects a build variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosinesampler_b200 import jet  # noqa: E402
from cosinesampler_b200.chain import make_head  # noqa: E402

dev = torch.device("cuda:0")
P = 1 << 20
for dim, residual in ((2, "helmholtz"), (3, "laplace")):
    J = 1 + 2 * dim
    jets = torch.randn(J, 16, P, device=dev)
    head = make_head(16, seed=0, device=dev)
    out = torch.empty_like(jets)
    for _ in range(5):
        jet.pde_head_step(jets, head, dim, residual, scale=1.0 / P)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for i in range(20):
        jet.pde_head_step(jets, head, dim, residual, scale=1.0 / P)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    nbytes = 8 * jets.numel()
    print(json.dumps({"lib": os.environ.get("COSINE_SAMPLER_LIB", "default"), "kernel": "HEAD%dd" % dim, "ms": ts[10],
                      "GBps": nbytes / ts[10] / 1e6}), flush=True)
