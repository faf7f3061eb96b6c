#!/usr/bin/env python
"""Launch every stage variant of the config-3 (2D) or config-4 (3D) step exactly twice (one
warm-up, one to profile) so that `ncu -k regex:cs_stage_kernel` captures a short, ordered list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosinesampler_b200 import ops  # noqa: E402
from cosinesampler_b200.autograd import cell_offsets  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dim, shape, P, kernel = {"cfg3": (2, (4, 16, 256, 256), 2 ** 20, 0),
                         "cfg4": (3, (4, 16, 64, 64, 64), 2 ** 22, 2)}[cfg]
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, C = shape[:2]
gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
inp = torch.rand(shape, device=dev)
grid = (torch.rand((1,) + gshape[1:], device=dev) * 2 - 1).repeat((N,) + (1,) * (len(gshape) - 1))
# the chain hands the operator a gOut that is EXPANDED over the cells (PIXEL's `val.sum(0)`, stride 0): profile
# the launches the bench really issues (round 1 profiled a dense gOut and over-counted the moved bytes)
gOut = torch.randn((1, C) + gshape[1:-1], device=dev).expand((N, C) + gshape[1:-1])
gOut2 = torch.randn((1, C) + gshape[1:-1], device=dev).expand((N, C) + gshape[1:-1])
gOG = torch.randn(gshape, device=dev)
gOgG = torch.randn(gshape, device=dev)
off = cell_offsets(N, True, dev)
staged = ops.stage(inp)
cases = [
    ("F", lambda: ops.forward(inp, grid, off, 0, True, kernel, True, staged=staged)),
    ("B[G]", lambda: ops.backward(gOut, inp, grid, off, 0, True, False, kernel, True, staged=staged)),
    ("B[I]", lambda: ops.backward(gOut, inp, grid, off, 0, True, True, kernel, True, staged=staged, want_grid=False)),
    ("BB[GO]", lambda: ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel, True,
                                             staged=staged, want=(False, True, True))),
    ("BB[IO]", lambda: ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel, True,
                                             staged=staged, want=(True, False, True))),
    ("BBB[IO+X2]", lambda: ops.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False, kernel,
                                                          True, staged=staged, gOutggOut=gOut2)),
]
for name, fn in cases:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    print(name, "done", flush=True)
