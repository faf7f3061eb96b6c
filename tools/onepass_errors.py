#!/usr/bin/env python
"""Measured error of the one-pass step against the fp64 oracle chain (loss, d loss / d cells, head gradients), as a
fraction of each tensor's scale: the numbers behind the tolerances of tests/test_gpu_onepass.py."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_onepass import CASES, _head, _oracle_step  # noqa: E402
from util import safe_coords  # noqa: E402
from cosinesampler_b200 import fused, jet  # noqa: E402

dev = torch.device("cuda:0")
for case in CASES:
    dim, shape, kernel, residual, K = case
    gen = torch.Generator().manual_seed(3 * dim + K)
    N, C = shape[:2]
    P = 6000
    cells0 = torch.rand(shape, generator=gen)
    coords0 = safe_coords(P, dim, shape[2:][::-1], N, True, gen).float()
    head64 = _head(C, K, seed=5, dtype=torch.float64)
    rl, rg, rh = _oracle_step(cells0, coords0, head64, dim, kernel, residual)
    row = {"case": "%dD %s %s K=%d C=%d" % (dim, kernel, residual, K, C)}
    modes = ["onepass"] + (["jets"] if K == 16 and C in (4, 8, 16, 32) else [])
    for mode in modes:
        cells = torch.nn.Parameter(cells0.clone().to(dev))
        head = _head(C, K, seed=5, device=dev)
        loss = jet.fused_pde_step(cells, coords0.to(dev).contiguous(), head, residual, kernel=kernel, mode=mode)
        rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
        relw = lambda a, b: float(((a.double().cpu() - b).abs() / (b.abs() + 1e-300)).max())
        row[mode] = {"loss_rel": abs(float(loss) - float(rl)) / abs(float(rl)),
                     "cells_grad_maxerr_over_scale": rel(cells.grad, rg),
                     "head_grads_maxerr_over_scale": max(rel(p.grad, g) for p, g in zip(head.parameters(), rh))}
    print(json.dumps(row), flush=True)
