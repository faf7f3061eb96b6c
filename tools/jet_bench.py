"""Time the fused jet operator (cosinesampler_b200/jet.py) at config-3 / config-4 sizes:
the two kernels alone, and a whole PDE step (jets + head + backward) next to the drop-in step.
CUDA events on the current stream, warm-up first; prints one JSON line per measurement."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosinesampler_b200 import jet, ops  # noqa: E402
from cosinesampler_b200.autograd import cell_offsets  # noqa: E402
from cosinesampler_b200.chain import make_head, training_step  # noqa: E402
from cosine_sampler_2d import CosineSampler2d  # noqa: E402
from cosine_sampler_3d import CosineSampler3d  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6550.7


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    for name, dim, shape, P, kernel, residual in (
            ("cfg3", 2, (4, 16, 256, 256), 1 << 20, "cosine", "helmholtz"),
            ("cfg4", 3, (4, 16, 64, 64, 64), 1 << 22, "smooth-step", "laplace")):
        N, C = shape[:2]
        T = 1
        for s in shape[2:]:
            T *= s
        cells = torch.rand(shape, device=dev)
        coords = torch.rand(P, dim, device=dev) * 2 - 1
        off = cell_offsets(N, True, dev)
        kn = {"cosine": 0, "smooth-step": 2}[kernel]
        staged = ops.stage(cells)
        G = torch.randn(1 + 2 * dim, C, P, device=dev)
        nbytes = jet.jet_bytes(dim, N, C, P, T, 2)
        t = timeit(lambda: jet.jet_forward(cells, coords, off, 0, True, kn, True, 2, staged=staged))
        print(json.dumps({"config": name, "kernel": "JET%dd[fwd]" % dim, "ms": t, "GBps": nbytes / t / 1e6,
                          "frac": nbytes / t / 1e6 / PEAK, "bytes": nbytes}), flush=True)
        t = timeit(lambda: jet.jet_backward(G, cells, coords, off, 0, True, kn, True, 2))
        print(json.dumps({"config": name, "kernel": "JET%dd[bwd] (+memset, +layout)" % dim, "ms": t,
                          "GBps": nbytes / t / 1e6, "frac": nbytes / t / 1e6 / PEAK, "bytes": nbytes}), flush=True)
        # whole step
        cols = [coords[:, a:a + 1].contiguous() for a in range(dim)]
        Pstep = min(P, 1 << 20)
        cols = [c[:Pstep] for c in cols]
        head = make_head(C, seed=1, device=dev)
        param = torch.nn.Parameter(cells.clone())
        J = jet.SamplerJet2d if dim == 2 else jet.SamplerJet3d
        D = CosineSampler2d if dim == 2 else CosineSampler3d

        def step_jet():
            param.grad = None
            return training_step(lambda c, x: J.apply(c, x, "zeros", True, kernel, True), param, cols, head,
                                 residual, jet=True)

        def step_dropin():
            param.grad = None
            return training_step(lambda c, g: D.apply(c, g, "zeros", True, kernel, True), param, cols, head,
                                 residual)
        xy = torch.cat(cols, -1).contiguous()

        def step_fused():
            param.grad = None
            return jet.fused_pde_step(param, xy, head, residual, kernel=kernel)
        jets_t = jet.jet_forward(cells, xy, off, 0, True, kn, True, 2, staged=staged)
        th = timeit(lambda: jet.pde_head_step(jets_t, head, dim, residual, scale=1.0 / Pstep))
        hb = 4 * 2 * jets_t.numel()
        print(json.dumps({"config": name, "kernel": "HEAD%dd (+memset)" % dim, "points": Pstep, "ms": th,
                          "GBps": hb / th / 1e6, "frac": hb / th / 1e6 / PEAK, "bytes": hb}), flush=True)
        lf = float(step_fused())
        gf = param.grad.clone()
        tf = timeit(step_fused, iters=20, warm=5)
        lj = float(step_jet())
        gj = param.grad.clone()
        ld = float(step_dropin())
        gd = param.grad.clone()
        tj = timeit(step_jet, iters=10, warm=3)
        td = timeit(step_dropin, iters=10, warm=3)
        print(json.dumps({"config": name, "points": Pstep, "ms_step_jet": tj, "ms_step_dropin": td,
                          "ms_step_fused": tf, "points_per_s_fused": Pstep / tf * 1e3,
                          "loss_fused": lf, "speedup_fused_vs_dropin": td / tf, "speedup": td / tj, "points_per_s_jet": Pstep / tj * 1e3,
                          "points_per_s_dropin": Pstep / td * 1e3, "loss_jet": lj, "loss_dropin": ld,
                          "max_abs_diff_cells_grad_over_max": float((gj - gd).abs().max() / gd.abs().max()),
                          "max_abs_diff_cells_grad_fused_over_max": float((gf - gd).abs().max() / gd.abs().max())}),
              flush=True)


if __name__ == "__main__":
    main()
