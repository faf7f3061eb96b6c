#!/usr/bin/env python
"""Stage-level timing on the GPU box: every C-ABI stage of this repo at BASELINE.json's
config-3 (2D) and config-4 (3D) sizes, next to the reference's own CUDA op recompiled for
sm_100 (oracle/_ref, when present) -- the kernel-for-kernel bar of BASELINE.md section 4.
CUDA events, 3 warm-up + median of 10.  One JSON object per line on stdout."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cosinesampler_b200 import ops  # noqa: E402
from cosinesampler_b200.autograd import cell_offsets  # noqa: E402


def timeit(fn, warm=3, iters=10, reps=5):
    """Median over `reps` of (time of `iters` back-to-back calls) / iters: the queue stays full,
    so host-side launch overhead does not leak into the device time of short kernels."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / iters)
    return statistics.median(ts)


def main():
    dev = torch.device("cuda:0")
    peak = 6550.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    ref_mods = {}
    try:
        if os.environ.get("CS_SKIP_REF") == "1":
            raise RuntimeError("skipped (CS_SKIP_REF=1)")
        from oracle import build_ref
        for d in (2, 3):
            ref_mods[d] = build_ref.load("_cosine_%dd" % d)
    except Exception as e:
        print(json.dumps({"note": "reference op unavailable: %s" % e}))
    configs = [("cfg3", 2, (4, 16, 256, 256), 2 ** 20, 0), ("cfg4", 3, (4, 16, 64, 64, 64), 2 ** 22, 2),
               # the shapes of the reference's own scripts (test_2d.py:26-38, test_3d.py:19-32)
               ("pixel2d", 2, (96, 4, 16, 16), 100000, 0), ("pixel3d", 3, (50, 4, 16, 16, 16), 100000, 0)]
    if len(sys.argv) > 1:
        configs = [c for c in configs if c[0] in sys.argv[1:]]
    for name, dim, shape, P, kernel in configs:
        torch.manual_seed(0)
        N, C = shape[:2]
        T = 1
        for s in shape[2:]:
            T *= s
        gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
        inp = torch.rand(shape, device=dev)
        grid = (torch.rand((1,) + gshape[1:], device=dev) * 2 - 1).repeat((N,) + (1,) * (len(gshape) - 1))
        gOut = torch.randn((N, C) + gshape[1:-1], device=dev)
        gOut2 = torch.randn((N, C) + gshape[1:-1], device=dev)
        gOG = torch.randn(gshape, device=dev)
        gOgG = torch.randn(gshape, device=dev)
        off = cell_offsets(N, True, dev)
        staged = ops.stage(inp)
        pairs = N * P
        G = 4 * N * C * T

        def rec(stage, ms, nbytes, impl):
            print(json.dumps({"config": name, "stage": stage, "impl": impl, "ms": round(ms, 4),
                              "alg_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1),
                              "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4),
                              "Gpairs_per_s": round(pairs / ms / 1e6, 2)}), flush=True)

        d = dim
        cases = [
            ("F", lambda: ops.forward(inp, grid, off, 0, True, kernel, True, staged=staged),
             pairs * (4 * d + 4 * C) + G),
            ("B[G]", lambda: ops.backward(gOut, inp, grid, off, 0, True, False, kernel, True, staged=staged),
             pairs * (8 * d + 4 * C) + G),
            ("B[I]", lambda: ops.backward(gOut, inp, grid, off, 0, True, True, kernel, True, staged=staged,
                                          want_grid=False), pairs * (4 * d + 4 * C) + G),
            ("B[IG]", lambda: ops.backward(gOut, inp, grid, off, 0, True, True, kernel, True, staged=staged),
             pairs * (8 * d + 4 * C) + 2 * G),
            ("BB[GO]", lambda: ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel,
                                                     True, staged=staged, want=(False, True, True)),
             pairs * (12 * d + 8 * C) + G),
            ("BB[I]", lambda: ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel,
                                                    True, staged=staged, want=(True, False, False)),
             pairs * (8 * d + 4 * C) + G),
            ("BB[IGO]", lambda: ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel,
                                                      True, staged=staged), pairs * (12 * d + 8 * C) + 2 * G),
            ("BBB[IO]", lambda: ops.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False,
                                                               kernel, True, staged=staged),
             pairs * (12 * d + 8 * C) + 2 * G),
            ("BBB[IO+X2]", lambda: ops.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True,
                                                                  False, kernel, True, staged=staged,
                                                                  gOutggOut=gOut2),
             pairs * (12 * d + 12 * C) + 2 * G),
            ("stage_to_channel_last", lambda: ops.to_channel_last(inp), 2 * G),
        ]
        for stage, fn, nbytes in cases:
            rec(stage, timeit(fn), nbytes, "ours")
        # expanded gOut (PIXEL: val.sum(0)) read in place
        gexp = gOut[:1].expand_as(gOut)
        rec("B[G] expanded gOut", timeit(lambda: ops.backward(gexp, inp, grid, off, 0, True, False, kernel, True,
                                                             staged=staged)), pairs * (8 * d + 4 * C) + G, "ours")
        ref = ref_mods.get(dim)
        if ref is not None:
            z1 = torch.zeros(1, device=dev)
            rcases = [
                ("F", lambda: ref.forward(inp, grid, off, 0, True, kernel, True), pairs * (4 * d + 4 * C) + G),
                ("B[G]", lambda: ref.backward(gOut, inp, grid, off, 0, True, False, kernel, True),
                 pairs * (8 * d + 4 * C) + G),
                ("B[IG]", lambda: ref.backward(gOut, inp, grid, off, 0, True, True, kernel, True),
                 pairs * (8 * d + 4 * C) + 2 * G),
                ("BB[IGO]", lambda: ref.backward_backward(z1, gOG, inp, grid, gOut, off, 0, True, False, kernel, True),
                 pairs * (12 * d + 8 * C) + 2 * G),
                ("BBB[IO]", lambda: ref.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False,
                                                                   kernel, True), pairs * (12 * d + 8 * C) + 2 * G),
            ]
            for stage, fn, nbytes in rcases:
                rec(stage, timeit(fn, warm=2, iters=5, reps=3), nbytes, "reference_cuda_op_sm100")
        del inp, grid, gOut, gOut2, gOG, gOgG, staged
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
