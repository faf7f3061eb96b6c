"""Half-precision path (csrc/cs_scalar.cuh `ScalarParamsT<__half>`, ops_f64.py): the third type of the reference's
AT_DISPATCH_FLOATING_TYPES_AND_HALF (cu2d:905), which the reference cannot run either (float offset through
TensorInfo<at::Half>, cu2d:914).  Contract: tensors are half (coordinates included), arithmetic is fp32, results
are rounded to half once, gInput accumulates in fp32.

  * every stage against the stage oracle evaluated in float64 on the SAME half-rounded inputs: the only error
    allowed is the final rounding to half (2^-11 relative) -- rtol 1e-3 + 1e-3 of the tensor's scale;
  * CosineSampler2d / 3d.apply on half tensors: value, d/d cells and d/d coordinates through torch autograd."""
import numpy as np
import pytest
import torch

from oracle import stage_oracle as so
from util import assert_close_scaled

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-3, atol_scale=1e-3)


def _safe_half_coords(P, dim, sizes, n_cells, multicell, gen, align=True, margin=0.02):
    """half-representable coordinates in (-1, 1) whose index stays `margin` away from a texel edge for every cell:
    the fp32 kernel and the fp64 oracle then pick the same texel."""
    offs = np.linspace(0, 1 - 1 / n_cells, n_cells) if multicell else np.zeros(1)
    out = torch.empty(P, dim, dtype=torch.float16)
    for a in range(dim):
        got = 0
        while got < P:
            g = (torch.rand(8 * P, generator=gen) * 1.96 - 0.98).half()
            gd = g.double().numpy()
            if align:
                i = (gd[:, None] + 1) / 2 * (sizes[a] - 1 - (1 if multicell else 0)) + offs[None, :]
            else:
                i = ((gd[:, None] + 1) * sizes[a] - 1) / 2 + offs[None, :]
            fr = i - np.floor(i)
            ok = torch.from_numpy(((fr > margin) & (fr < 1 - margin)).all(1))
            g = g[ok][:P - got]
            out[got:got + len(g), a] = g
            got += len(g)
    return out


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_f16_stages_match_the_stage_oracle_on_half_rounded_inputs(cuda, dim, kernel):
    from cosinesampler_b200 import ops
    from cosinesampler_b200.autograd import cell_offsets
    gen = torch.Generator().manual_seed(23 * dim + kernel)
    N, C, P = 3, 5, 777
    sizes = (9, 12) if dim == 2 else (6, 7, 8)                  # (H, W) / (D, H, W)
    H = torch.float16
    inp = torch.rand((N, C) + sizes, generator=gen).to(H)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    gOut = torch.randn((N, C) + gshape[1:-1], generator=gen).to(H)
    gOut2 = torch.randn((N, C) + gshape[1:-1], generator=gen).to(H)
    gOG = torch.randn(gshape, generator=gen).to(H)
    gOgG = torch.randn(gshape, generator=gen).to(H)
    gOI = torch.randn(inp.shape, generator=gen).to(H)
    d = lambda t: t.to(cuda)
    f = lambda t: t.double()
    for multicell in (True, False):
        off = cell_offsets(N, multicell, torch.device("cpu")).clone()
        for pad, align in ((0, True), (1, True), (2, True), (0, False)):
            # the 2D forward maps with align_corners=True whatever the flag says (cu2d:307-308): the points must be
            # safe for both maps there
            xyz = sizes[::-1]
            c = _safe_half_coords(4 * P, dim, xyz, N, multicell, gen, align=align)
            if dim == 2 and not align:
                offs = np.linspace(0, 1 - 1 / N, N) if multicell else np.zeros(1)
                keep = torch.ones(c.shape[0], dtype=torch.bool)
                for a in range(dim):
                    i = (c[:, a].double().numpy()[:, None] + 1) / 2 * (xyz[a] - 1 - (1 if multicell else 0)) + offs[None, :]
                    fr = i - np.floor(i)
                    keep &= torch.from_numpy(((fr > 0.02) & (fr < 0.98)).all(1))
                c = c[keep]
            assert c.shape[0] >= P
            grid = c[:P].reshape((1,) + gshape[1:]).expand(gshape).contiguous()
            kw = dict(pad=pad, align=align, kernel=kernel, multicell=multicell, index_mode=2)
            what = "f16 %dD k=%d pad=%d align=%s mc=%s " % (dim, kernel, pad, align, multicell)
            o = ops.forward(d(inp), d(grid), d(off), pad, align, kernel, multicell)
            assert o.dtype == H
            assert_close_scaled(o, so.forward(f(inp), f(grid), off, **kw), what + "F", **TOL)
            gI, gG = ops.backward(d(gOut), d(inp), d(grid), d(off), pad, align, True, kernel, multicell)
            rI, rG = so.backward(f(gOut), f(inp), f(grid), off, input_requires_grad=True, **kw)
            assert gI.dtype == H and gG.dtype == H
            assert_close_scaled(gI, rI, what + "B gInput", **TOL)
            assert_close_scaled(gG, rG, what + "B gGrid", **TOL)
            for use in (False, True):
                a = ops.backward_backward(d(gOI) if use else None, d(gOG), d(inp), d(grid), d(gOut), d(off), pad, align,
                                          use, kernel, multicell)
                b = so.backward_backward(f(gOI) if use else None, f(gOG), f(inp), f(grid), f(gOut), off,
                                         input_requires_grad=use, **kw)
                for nm, x, y in zip(("gInput", "gGrid", "ggOut"), a, b):
                    assert_close_scaled(x, y, what + "BB %s U=%s" % (nm, use), **TOL)
            a = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad, align, False,
                                               kernel, multicell)
            b = so.backward_backward_backward(f(inp), f(grid), f(gOut), f(gOG), f(gOgG), off, **kw)
            for nm, x, y in zip(("gInput", "ggOut"), a, b):
                assert_close_scaled(x, y, what + "BBB " + nm, **TOL)
            a2 = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad, align, False,
                                                kernel, multicell, gOutggOut=d(gOut2))
            bI = so.backward_backward(None, f(gOG), f(inp), f(grid), f(gOut2), off, input_requires_grad=False, **kw)[0]
            assert_close_scaled(a2[0], b[0] + bI, what + "BBB+X2 gInput", **TOL)


@pytest.mark.parametrize("dim", [2, 3])
def test_f16_autograd_function_value_and_first_gradients(cuda, dim):
    """The reference-facing Function on half tensors (what `AT_DISPATCH..._AND_HALF` promises a caller)."""
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    from cosinesampler_b200.autograd import cell_offsets
    S = CosineSampler2d if dim == 2 else CosineSampler3d
    gen = torch.Generator().manual_seed(5 + dim)
    shape = (4, 4, 16, 16) if dim == 2 else (4, 4, 10, 10, 10)
    N, P = shape[0], 1500
    cells0 = torch.rand(shape, generator=gen).half()
    c = _safe_half_coords(P, dim, shape[2:][::-1], N, True, gen)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid0 = c.reshape((1,) + gshape[1:]).expand(gshape).contiguous()
    cells = cells0.to(cuda).requires_grad_(True)
    grid = grid0.to(cuda).requires_grad_(True)
    out = S.apply(cells, grid, "zeros", True, "cosine", True)
    assert out.dtype == torch.float16
    w = torch.randn(out.shape, generator=gen).half()
    gc, gg = torch.autograd.grad(out, (cells, grid), grad_outputs=w.to(cuda))
    off = cell_offsets(N, True, torch.device("cpu")).clone()
    kw = dict(pad=0, align=True, kernel=0, multicell=True, index_mode=2)
    assert_close_scaled(out, so.forward(cells0.double(), grid0.double(), off, **kw), "f16 apply value", **TOL)
    rI, rG = so.backward(w.double(), cells0.double(), grid0.double(), off, input_requires_grad=True, **kw)
    assert_close_scaled(gc, rI, "f16 apply d/d cells", **TOL)
    assert_close_scaled(gg, rG, "f16 apply d/d grid", **TOL)
