"""Parity against the reference's ORIGINAL CUDA op, recompiled for sm_100 from the sources
under /root/reference into oracle/_ref/ (oracle/build_ref.py; skipped when not built).

The reference is built with --use_fast_math (setup.py:37): its index map is a single fma and
its sin/cos are MUFU approximations.  We therefore run our kernels with index_mode='fused'
(same cell decisions) and allow the MUFU-sized error: rtol 1e-4 + 2e-5 of scale, with at most
1e-4 of the elements as cell-flip outliers."""
import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu


def _ref(dim):
    from oracle import build_ref
    mod = build_ref.load("_cosine_%dd" % dim)
    if mod is None:
        pytest.skip("oracle/_ref/_cosine_%dd.so not built" % dim)
    return mod


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("multicell", [True, False])
@pytest.mark.parametrize("dim", [2, 3])
def test_every_stage_matches_the_reference_cuda_op(cuda, dim, kernel, multicell):
    ref = _ref(dim)
    from cosinesampler_b200 import ops
    from cosinesampler_b200.autograd import cell_offsets
    torch.manual_seed(dim * 10 + kernel)
    N, C, P = 4, 16, 2 ** 16
    shape = (N, C, 128, 128) if dim == 2 else (N, C, 32, 32, 32)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    inp = torch.rand(shape, device=cuda)
    grid = torch.rand(gshape, device=cuda) * 2 - 1
    gOut = torch.randn((N, C) + gshape[1:-1], device=cuda)
    gOG = torch.randn(gshape, device=cuda)
    gOgG = torch.randn(gshape, device=cuda)
    gOI = torch.randn(shape, device=cuda)
    off = cell_offsets(N, multicell, cuda)
    tol = dict(rtol=1e-4, atol_scale=2e-5, max_outlier_frac=1e-4)
    ops.set_index_mode("fused")
    try:
        o = ops.forward(inp, grid, off, 0, True, kernel, multicell)
        r = ref.forward(inp, grid, off, 0, True, kernel, multicell)
        assert_close_scaled(o, r, "F", **tol)

        gI, gG = ops.backward(gOut, inp, grid, off, 0, True, True, kernel, multicell)
        rI, rG = ref.backward(gOut, inp, grid, off, 0, True, True, kernel, multicell)
        assert_close_scaled(gI, rI, "B gInput", **tol)
        assert_close_scaled(gG, rG, "B gGrid", **tol)

        for use in (False, True):
            a = ops.backward_backward(gOI if use else None, gOG, inp, grid, gOut, off, 0, True, use,
                                      kernel, multicell)
            b = ref.backward_backward(gOI if use else torch.zeros(1, device=cuda), gOG, inp, grid, gOut, off,
                                      0, True, use, kernel, multicell)
            for nm, x, y in zip(("gInput", "gGrid", "ggOut"), a, b):
                assert_close_scaled(x, y, "BB %s (gOutInput=%s)" % (nm, use), **tol)

        a = ops.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False, kernel, multicell)
        b = ref.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False, kernel, multicell)
        assert_close_scaled(a[0], b[0], "BBB gInput", **tol)
        assert_close_scaled(a[1], b[1], "BBB ggOut", **tol)
    finally:
        ops.set_index_mode("separate")


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_linear_and_smoothstep_match_the_reference_op_at_1e_5(cuda, dim, kernel):
    """No transcendental in these kernels, so the reference build's MUFU approximations play no part: every
    stage must meet the north star's rtol 1e-5 (+ 1e-5 of scale for the sums; cell flips between the two
    fp32 index maps excepted, <= 1e-4 of the elements) against the reference's own CUDA op."""
    ref = _ref(dim)
    from cosinesampler_b200 import ops
    from cosinesampler_b200.autograd import cell_offsets
    torch.manual_seed(dim * 7 + kernel)
    N, C, P = 4, 16, 2 ** 16
    shape = (N, C, 128, 128) if dim == 2 else (N, C, 32, 32, 32)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    inp = torch.rand(shape, device=cuda)
    grid = torch.rand(gshape, device=cuda) * 2 - 1
    gOut = torch.randn((N, C) + gshape[1:-1], device=cuda)
    gOG = torch.randn(gshape, device=cuda)
    gOgG = torch.randn(gshape, device=cuda)
    off = cell_offsets(N, True, cuda)
    tol = dict(rtol=1e-5, atol_scale=1e-5, max_outlier_frac=1e-4)
    ops.set_index_mode("fused")
    try:
        assert_close_scaled(ops.forward(inp, grid, off, 0, True, kernel, True),
                            ref.forward(inp, grid, off, 0, True, kernel, True), "F", **tol)
        a = ops.backward(gOut, inp, grid, off, 0, True, True, kernel, True)
        b = ref.backward(gOut, inp, grid, off, 0, True, True, kernel, True)
        for nm, x, y in zip(("gInput", "gGrid"), a, b):
            assert_close_scaled(x, y, "B " + nm, **tol)
        a = ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, kernel, True)
        b = ref.backward_backward(torch.zeros(1, device=cuda), gOG, inp, grid, gOut, off, 0, True, False, kernel, True)
        for nm, x, y in zip(("gInput", "gGrid", "ggOut"), a, b):
            assert_close_scaled(x, y, "BB " + nm, **tol)
        a = ops.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False, kernel, True)
        b = ref.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, 0, True, False, kernel, True)
        for nm, x, y in zip(("gInput", "ggOut"), a, b):
            assert_close_scaled(x, y, "BBB " + nm, **tol)
    finally:
        ops.set_index_mode("separate")
