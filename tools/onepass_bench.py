#!/usr/bin/env python
"""Time the pieces of the one-pass fused step (cosinesampler_b200/fused.py) and the round-1 jets path
on the BASELINE shapes: one JSON line per (workload, points, variant).

    python tools/onepass_bench.py [cfg3|cfg4] [--points 1048576,4194304] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cosinesampler_b200 import chain, fused, jet, ops  # noqa: E402

SHAPES = {"cfg3": (2, (4, 16, 256, 256), "cosine", "helmholtz"),
          "cfg4": (3, (4, 16, 64, 64, 64), "smooth-step", "laplace")}


class Prof:
    def __init__(self):
        self.records = []

    def record(self, label, nbytes, s, e):
        self.records.append((label, s, e))

    def summary(self):
        agg = {}
        for label, s, e in self.records:
            a = agg.setdefault(label, [0, 0.0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
        return {k: round(v[1] / v[0], 4) for k, v in agg.items()}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    prof = Prof()
    ops.profiler = prof
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ops.profiler = None
    return ms, prof.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg3")
    ap.add_argument("--points", default="1048576,4194304")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="onepass,onepass_noagg,onepass_unbinned,jets")
    ap.add_argument("--once", action="store_true", help="one step per variant, no timing (for ncu)")
    ap.add_argument("--chunk", type=int, default=0, help="points per chunk of the one-pass step (0: one chunk)")
    args = ap.parse_args()
    dim, shape, kernel, residual = SHAPES[args.workload]
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    cells = torch.nn.Parameter(torch.rand(shape, generator=g).to(dev))
    head = chain.make_head(shape[1], seed=0, device=dev)
    for P in [int(x) for x in args.points.split(",")]:
        coords = (torch.rand(P, dim, generator=g) * 2 - 1).to(dev)
        ref = None
        for variant in args.variants.split(","):
            def step():
                cells.grad = None
                for p in head.parameters():
                    p.grad = None
                if variant == "jets":
                    return jet.fused_pde_step(cells, coords, head, residual, kernel=kernel, mode="jets")
                kw = {"onepass": dict(bin=True, aggregate="auto"), "onepass_noagg": dict(bin=True, aggregate="off"),
                      "onepass_unbinned": dict(bin=False, aggregate="off")}[variant]
                return fused.one_pass_pde_step(cells, coords, head, residual, kernel=kernel, chunk=args.chunk or None, **kw)
            if args.once:
                step()
                torch.cuda.synchronize()
                continue
            ms, stages = timed(step, args.reps)
            loss = float(step())
            gsum = float(cells.grad.double().abs().sum())
            if ref is None:
                ref = (loss, gsum)
            print(json.dumps({"workload": args.workload, "points": P, "variant": variant, "chunk": args.chunk, "ms_per_step": round(ms, 4),
                              "points_per_s": P / (ms * 1e-3), "ms_per_2^20": round(ms * 2 ** 20 / P, 4),
                              "stages_ms": stages, "loss": loss, "grad_abs_sum": gsum}), flush=True)


if __name__ == "__main__":
    main()
