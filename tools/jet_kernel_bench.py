"""Time cs_jet_forward / cs_jet_backward alone at config-3 sizes; COSINE_SAMPLER_LIB selects a build variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosinesampler_b200 import jet, ops  # noqa: E402
from cosinesampler_b200.autograd import cell_offsets  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
P = 1 << 20
cells = torch.rand(4, 16, 256, 256, device=dev)
coords = torch.rand(P, 2, device=dev) * 2 - 1
off = cell_offsets(4, True, dev)
staged = ops.stage(cells)
G = torch.randn(5, 16, P, device=dev)
acc = jet.new_accumulator(cells)


def timeit(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for i in range(20):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))[10]


ref = jet.jet_forward(cells, coords, off, 0, True, 0, True, 2, staged=staged)
tf = timeit(lambda: jet.jet_forward(cells, coords, off, 0, True, 0, True, 2, staged=staged))
tb = timeit(lambda: jet.jet_backward_into(acc, G, cells, coords, off, 0, True, 0, True, 2))
print(json.dumps({"lib": os.path.basename(os.environ.get("COSINE_SAMPLER_LIB", "default")), "fwd_ms": tf, "bwd_ms": tb,
                  "checksum": float(ref.double().sum())}), flush=True)
