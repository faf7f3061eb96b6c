#!/usr/bin/env python
"""Launch the three kernels of the fused jet step (cs_jet_fwd_kernel, cs_pde_head_kernel,
cs_jet_bwd_kernel) exactly twice each at config-3 (2D) or config-4 (3D) sizes, so that
`ncu -k regex:"cs_jet|cs_pde_head"` captures a short, ordered list (second launch = the one kept)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosinesampler_b200 import jet, ops  # noqa: E402
from cosinesampler_b200.autograd import cell_offsets  # noqa: E402
from cosinesampler_b200.chain import make_head  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dim, shape, P, kernel, residual = {"cfg3": (2, (4, 16, 256, 256), 2 ** 20, 0, "helmholtz"),
                                   "cfg4": (3, (4, 16, 64, 64, 64), 2 ** 22, 2, "laplace")}[cfg]
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, C = shape[:2]
cells = torch.rand(shape, device=dev)
coords = torch.rand(P, dim, device=dev) * 2 - 1
off = cell_offsets(N, True, dev)
staged = ops.stage(cells)
head = make_head(C, seed=0, device=dev)
acc = jet.new_accumulator(cells)
for _ in range(2):
    jets = jet.jet_forward(cells, coords, off, 0, True, kernel, True, 2, staged=staged)
torch.cuda.synchronize()
print("fwd done", flush=True)
for _ in range(2):
    _, gJets, _, _ = jet.pde_head_step(jets, head, dim, residual, scale=1.0 / P)
torch.cuda.synchronize()
print("head done", flush=True)
for _ in range(2):
    jet.jet_backward_into(acc, gJets, cells, coords, off, 0, True, kernel, True, 2)
torch.cuda.synchronize()
print("bwd done", flush=True)
