#!/usr/bin/env python
"""Run the one-pass fused step twice on 2^22 points (or argv[2] points) of config 3 (2D) or config 4 (3D) so that
`ncu -k regex:cs_pde_fused` captures one warm-up launch and one launch to keep (tools/summarize_ncu.py
onepass_cfg3 / onepass_cfg4)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosinesampler_b200 import chain, fused  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dim, shape, kernel, residual = {"cfg3": (2, (4, 16, 256, 256), "cosine", "helmholtz"),
                                "cfg4": (3, (4, 16, 64, 64, 64), "smooth-step", "laplace")}[cfg]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cells = torch.nn.Parameter(torch.rand(shape, generator=g).to(dev))
head = chain.make_head(shape[1], seed=0, device=dev)
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 22
coords = (torch.rand(P, dim, generator=g) * 2 - 1).to(dev)
for _ in range(2):
    cells.grad = None
    loss = fused.one_pass_pde_step(cells, coords, head, residual, kernel=kernel)
    torch.cuda.synchronize()
print(cfg, "loss", float(loss), flush=True)
