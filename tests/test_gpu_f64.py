"""Double-precision path (csrc/cs_scalar.cuh, ops_f64.py): what the reference's AT_DISPATCH_FLOATING_TYPES_AND_HALF
(cu2d:905) promises for float64 and cannot deliver (its float offset meets TensorInfo<double>, cu2d:914).

  * every stage against the stage oracle computed in float64 with the same fp32 offsets (1e-11);
  * the reference's derivative chain (test_2d.py / test_3d.py) through CosineSampler2d / 3d on float64 tensors
    against nested autograd over the oracle sampler in float64 (1e-9): loss and d loss / d cells included."""
import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, grid_sample_3d, make_head
from util import assert_close_scaled

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-11, atol_scale=1e-11)


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_f64_stages_match_the_stage_oracle(cuda, dim, kernel):
    from cosinesampler_b200 import ops
    from cosinesampler_b200.autograd import cell_offsets
    gen = torch.Generator().manual_seed(17 * dim + kernel)
    N, C, P = 3, 5, 777
    sizes = (9, 12) if dim == 2 else (6, 7, 8)
    inp = torch.rand((N, C) + sizes, generator=gen, dtype=torch.float64)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid = torch.rand(gshape, generator=gen, dtype=torch.float64) * 2.6 - 1.3          # some points out of range
    gOut = torch.randn((N, C) + gshape[1:-1], generator=gen, dtype=torch.float64)
    gOut2 = torch.randn((N, C) + gshape[1:-1], generator=gen, dtype=torch.float64)
    gOG = torch.randn(gshape, generator=gen, dtype=torch.float64)
    gOgG = torch.randn(gshape, generator=gen, dtype=torch.float64)
    gOI = torch.randn(inp.shape, generator=gen, dtype=torch.float64)
    d = lambda t: t.to(cuda)
    for multicell in (True, False):
        off = cell_offsets(N, multicell, torch.device("cpu")).clone()
        for pad, align in ((0, True), (1, True), (2, True), (0, False)):
            kw = dict(pad=pad, align=align, kernel=kernel, multicell=multicell, index_mode=2)
            what = "f64 %dD k=%d pad=%d align=%s mc=%s " % (dim, kernel, pad, align, multicell)
            o = ops.forward(d(inp), d(grid), d(off), pad, align, kernel, multicell)
            assert o.dtype == torch.float64
            assert_close_scaled(o, so.forward(inp, grid, off, **kw), what + "F", **TOL)
            gI, gG = ops.backward(d(gOut), d(inp), d(grid), d(off), pad, align, True, kernel, multicell)
            rI, rG = so.backward(gOut, inp, grid, off, input_requires_grad=True, **kw)
            assert_close_scaled(gI, rI, what + "B gInput", **TOL)
            assert_close_scaled(gG, rG, what + "B gGrid", **TOL)
            for use in (False, True):
                a = ops.backward_backward(d(gOI) if use else None, d(gOG), d(inp), d(grid), d(gOut), d(off), pad, align,
                                          use, kernel, multicell)
                b = so.backward_backward(gOI if use else None, gOG, inp, grid, gOut, off, input_requires_grad=use, **kw)
                for nm, x, y in zip(("gInput", "gGrid", "ggOut"), a, b):
                    assert_close_scaled(x, y, what + "BB %s U=%s" % (nm, use), **TOL)
            a = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad, align, False,
                                               kernel, multicell)
            b = so.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, **kw)
            for nm, x, y in zip(("gInput", "ggOut"), a, b):
                assert_close_scaled(x, y, what + "BBB " + nm, **TOL)
            # fused b_input pass: gInput + scatter(gOutggOut * A)
            a2 = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad, align, False,
                                                kernel, multicell, gOutggOut=d(gOut2))
            bI = so.backward_backward(None, gOG, inp, grid, gOut2, off, input_requires_grad=False, **kw)[0]
            assert_close_scaled(a2[0], b[0] + bI, what + "BBB+X2 gInput", **TOL)


@pytest.mark.parametrize("dim", [2, 3])
def test_f64_derivative_chain_matches_the_oracle_sampler(cuda, dim):
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    S = CosineSampler2d if dim == 2 else CosineSampler3d
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    gen = torch.Generator().manual_seed(dim)
    shape = (4, 4, 16, 16) if dim == 2 else (3, 4, 10, 10, 10)
    cells0 = torch.rand(shape, generator=gen, dtype=torch.float64)
    coords0 = torch.rand(1500, dim, generator=gen, dtype=torch.float64) * 1.9 - 0.95
    residual = "t2d" if dim == 2 else "laplace"

    def run(sampler, device):
        cells = cells0.to(device).requires_grad_(True)
        coords = [coords0[:, a:a + 1].to(device).requires_grad_(True) for a in range(dim)]
        head = make_head(shape[1], seed=1, dtype=torch.float64).to(device)
        return derivative_chain(sampler, cells, coords, head, residual=residual)

    ours = run(lambda c, g: S.apply(c, g, "zeros", True, "cosine", True), cuda)
    ref = run(lambda c, g: fn(c, g, step="cosine", offset=True), "cpu")
    for k in ref:
        assert ours[k].dtype == torch.float64
        assert_close_scaled(ours[k], ref[k].reshape(ours[k].shape), "f64 %dD chain %s" % (dim, k), rtol=1e-9, atol_scale=1e-9)
