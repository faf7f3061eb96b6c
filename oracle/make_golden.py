"""ORACLE tooling (test infrastructure): mint golden vectors from the REAL
reference sampler.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/ref_sampler_*.npz

It imports `/root/reference/test/grid_sampler.py` unmodified and evaluates
`grid_sample_2d` / `grid_sample_3d` plus the derivative chain of
`test/test_2d.py` / `test/test_3d.py` on small seeded inputs, on the CPU, in
fp64 and fp32.  Two shims are needed to run that file without a GPU and
without the compiled extension, neither of which changes its arithmetic:

  * its first line imports `cosine_sampler_2d.modules_2d` (the CUDA op); a
    placeholder module of that name is registered first;
  * `offset=True` moves the offsets with `.to("cuda")` (`grid_sampler.py:34,121`);
    while the reference function runs, `torch.Tensor.to` ignores a "cuda"
    device argument.

The reference has no golden vectors of its own, so these
files are what pins `oracle/grid_sampler_oracle.py` to the reference.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.grid_sampler_oracle import derivative_chain, make_head  # noqa: E402


def import_reference():
    stub_pkg = types.ModuleType("cosine_sampler_2d")
    stub_mod = types.ModuleType("cosine_sampler_2d.modules_2d")
    stub_mod.CosineSampler2d = None
    stub_pkg.modules_2d = stub_mod
    saved = {k: sys.modules.get(k) for k in ("cosine_sampler_2d", "cosine_sampler_2d.modules_2d")}
    sys.modules["cosine_sampler_2d"] = stub_pkg
    sys.modules["cosine_sampler_2d.modules_2d"] = stub_mod
    sys.path.insert(0, "/root/reference/test")
    try:
        import grid_sampler as ref  # noqa
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


@contextlib.contextmanager
def cuda_to_is_noop():
    real_to = torch.Tensor.to

    def to(self, *a, **kw):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            a = a[1:]
            if not a and not kw:
                return self
        return real_to(self, *a, **kw)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = real_to


CASES = [
    # name, dim, cells shape, points, residual
    ("2d", 2, (3, 4, 9, 7), 96, "t2d"),
    ("3d", 3, (3, 2, 6, 6, 6), 80, "laplace"),
]
STEPS = {2: ["cosine", "smoothstep", "bilinear"], 3: ["cosine", "smoothstep", "trilinear"]}


def main():
    ref = import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, nd, shape, P, residual in CASES:
        gen = torch.Generator().manual_seed(1234 + nd)
        cells64 = torch.rand(shape, generator=gen, dtype=torch.float64)
        # open interval so that no coordinate sits exactly on a cell edge
        coords64 = torch.rand(P, nd, generator=gen, dtype=torch.float64) * 1.96 - 0.98
        for dtype in (torch.float64, torch.float32):
            tag = "f64" if dtype == torch.float64 else "f32"
            blob = {"cells": cells64.to(dtype).numpy(), "coords": coords64.to(dtype).numpy()}
            for step in STEPS[nd]:
                for offset in (True, False):
                    cells = cells64.to(dtype).clone().requires_grad_(True)
                    coords = [coords64[:, a:a + 1].to(dtype).clone().requires_grad_(True)
                              for a in range(nd)]
                    head = make_head(shape[1], seed=7, dtype=dtype)
                    fn = ref.grid_sample_2d if nd == 2 else ref.grid_sample_3d

                    def sampler(c, g, fn=fn, step=step, offset=offset):
                        with cuda_to_is_noop():
                            return fn(c, g, step=step, offset=offset)

                    q = derivative_chain(sampler, cells, coords, head, residual=residual)
                    key = "%s|%d" % (step, int(offset))
                    for k, v in q.items():
                        blob["%s|%s" % (key, k)] = v.detach().numpy()
            path = os.path.join(outdir, "ref_sampler_%s_%s.npz" % (name, tag))
            np.savez_compressed(path, **blob)
            print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
