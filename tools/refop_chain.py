#!/usr/bin/env python
"""Drive the reference's ORIGINAL CUDA op (oracle/_ref/_cosine_{2,3}d.so, built from the sources
under /root/reference by oracle/build_ref.py) through the same PIXEL step as bench.py, next to
this repo's op: the kernel-for-kernel bar of BASELINE.md section 4 at chain level.

The three autograd Functions below issue the reference's call pattern (modules_2d.py:20-111:
every first backward scatters gInput, the triple backward is a BBB call plus a second BB call
with a `ones` gOutInput) on top of a backend object with the four pybind entry points.  They are
written for this tool; the reference's Python layer cannot be shipped.  Its host syncs
(`.any().item()`) are not reproduced -- the flags are decided by `None`-ness -- which only helps
the reference's timing.  Test infrastructure, not product code."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_refop_sampler(backend, kernel_code):
    def offsets(n, device):
        return torch.linspace(0, 1 - (1 / n), n).to(device)

    class BB(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, grid, gOut, gOutInput, gOutGrid, off):
            flag = gOutInput is not None
            goi = gOutInput.contiguous() if flag else torch.zeros(1, device=inp.device)
            gI, gG, ggO = backend.backward_backward(goi, gOutGrid.contiguous(), inp, grid, gOut.contiguous(), off,
                                                    0, True, flag, kernel_code, True)
            ctx.save_for_backward(inp, grid, gOut, gOutGrid)
            ctx.off = off
            return gI, gG, ggO

        @staticmethod
        def backward(ctx, gOutgInput, gOutgGrid, gOutggOut):
            inp, grid, gOut, gOutGrid = ctx.saved_tensors
            gI, ggO = backend.backward_backward_backward(inp, grid, gOut.contiguous(), gOutGrid.contiguous(),
                                                         gOutgGrid.contiguous(), ctx.off, 0, True, False,
                                                         kernel_code, True)
            b_input, _, _ = BB.apply(inp, grid, gOutggOut.contiguous(), torch.ones_like(inp), gOutGrid, ctx.off)
            return gI + b_input, None, ggO, None, None, None

    class B(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, grid, gOut, off):
            gI, gG = backend.backward(gOut, inp, grid, off, 0, True, bool(inp.requires_grad), kernel_code, True)
            ctx.save_for_backward(inp, grid, gOut)
            ctx.off = off
            return gI, gG

        @staticmethod
        def backward(ctx, gOutInput, gOutGrid):
            inp, grid, gOut = ctx.saved_tensors
            # the reference tests the VALUE of gOutInput (mod2d:87); a materialised zero means "no"
            gI, gG, ggO = BB.apply(inp, grid, gOut, None, gOutGrid, ctx.off)
            return gI, gG, ggO, None

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, grid):
            off = offsets(inp.shape[0], inp.device)
            out = backend.forward(inp, grid, off, 0, True, kernel_code, True)
            ctx.save_for_backward(inp, grid)
            ctx.off = off
            return out

        @staticmethod
        def backward(ctx, gOut):
            inp, grid = ctx.saved_tensors
            return B.apply(inp, grid, gOut.contiguous(), ctx.off)

    return F.apply


def time_step(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    from oracle import build_ref
    from cosinesampler_b200 import chain
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    dev = torch.device("cuda:0")
    cfgs = [("cfg3", 2, (4, 16, 256, 256), 2 ** 20, 0, "cosine", "helmholtz"),
            ("cfg4", 3, (4, 16, 64, 64, 64), 2 ** 20, 2, "smooth-step", "laplace")]
    for name, dim, shape, P, kcode, kname, residual in cfgs:
        ref = build_ref.load("_cosine_%dd" % dim)
        if ref is None:
            print(json.dumps({"config": name, "note": "oracle/_ref not built"}))
            continue
        torch.manual_seed(0)
        cells = torch.nn.Parameter(torch.rand(shape, device=dev))
        coords = torch.rand(P, dim, device=dev) * 2 - 1
        head = chain.make_head(shape[1], seed=0, device=dev)
        S = CosineSampler2d if dim == 2 else CosineSampler3d
        ours = lambda c, g: S.apply(c, g, "zeros", True, kname, True)
        theirs = make_refop_sampler(ref, kcode)

        def step(sampler):
            cells.grad = None
            for p in head.parameters():
                p.grad = None
            return chain.training_step(sampler, cells, [coords[:, a:a + 1] for a in range(dim)], head,
                                       residual=residual)
        from cosinesampler_b200 import jet

        def step_fused():
            cells.grad = None
            for p in head.parameters():
                p.grad = None
            return jet.fused_pde_step(cells, coords, head, residual, kernel=kname)
        l_fused = float(step_fused()); g_fused = cells.grad.clone()
        t_fused = time_step(step_fused, warm=3, iters=10)
        l_ours = float(step(ours)); g_ours = cells.grad.clone()
        l_ref = float(step(theirs)); g_ref = cells.grad.clone()
        rel = float((g_ours - g_ref).abs().max() / g_ref.abs().max())
        t_ours = time_step(lambda: step(ours))
        t_ref = time_step(lambda: step(theirs))
        print(json.dumps({"config": name, "points": P, "ms_step_ours": round(t_ours, 3),
                          "ms_step_reference_cuda_op": round(t_ref, 3), "speedup": round(t_ref / t_ours, 2),
                          "ms_step_fused": round(t_fused, 3), "speedup_fused_vs_reference_cuda_op": round(t_ref / t_fused, 1),
                          "loss_ours": l_ours, "loss_reference_op": l_ref, "loss_fused": l_fused,
                          "max_abs_diff_cells_grad_over_max": rel,
                          "max_abs_diff_cells_grad_fused_vs_reference_op_over_max":
                              float((g_fused - g_ref).abs().max() / g_ref.abs().max())}), flush=True)


if __name__ == "__main__":
    main()
