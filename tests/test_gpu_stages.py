"""Stage-level parity on the GPU: every C-ABI entry point (through cosinesampler_b200.ops,
the mirror of the reference's pybind surface) against oracle/stage_oracle.py on identical
seeded inputs.  The oracle evaluates the index map in fp32 exactly like the kernels and
everything else in fp64, so the comparison has no cell-flip ambiguity.

Tolerance (north star: rtol 1e-5 in fp32, every derivative order):
    |a - b| <= 1e-5 |b| + 1e-5 max|b|        (tests/util.py)
"""
import itertools

import pytest
import torch

from oracle import stage_oracle as so
from util import assert_close_scaled

pytestmark = pytest.mark.gpu

KERNELS = [0, 1, 2]   # cosine, linear, smoothstep


def _inputs(dim, N, C, P, seed, lo=-1.0, hi=1.0, sizes=None):
    gen = torch.Generator().manual_seed(seed)
    if sizes is None:
        sizes = (9, 12) if dim == 2 else (6, 7, 8)
    inp = torch.rand((N, C) + tuple(sizes), generator=gen)
    shape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid = torch.rand(shape, generator=gen) * (hi - lo) + lo
    gOut = torch.randn((N, C) + shape[1:-1], generator=gen)
    gOG = torch.randn(shape, generator=gen)
    gOgG = torch.randn(shape, generator=gen)
    gOI = torch.randn(inp.shape, generator=gen)
    gOggO = torch.randn(gOut.shape, generator=gen)
    return inp, grid, gOut, gOG, gOgG, gOI, gOggO


def _offset(N, multicell):
    from cosinesampler_b200.autograd import cell_offsets
    return cell_offsets(N, multicell, torch.device("cpu")).clone()


def _run_all(cuda, dim, N, C, P, kernel, multicell, pad=0, align=True, seed=0, lo=-1.0, hi=1.0,
             sizes=None, index_mode="separate", small_cell="auto"):
    from cosinesampler_b200 import ops
    ops.set_index_mode(index_mode)
    ops.set_small_cell(small_cell)
    im = 0 if index_mode == "separate" else 1
    inp, grid, gOut, gOG, gOgG, gOI, gOggO = _inputs(dim, N, C, P, seed, lo, hi, sizes)
    off = _offset(N, multicell)
    d = lambda t: t.to(cuda)
    kw = dict(pad=pad, align=align, kernel=kernel, multicell=multicell, index_mode=im)
    try:
        # F
        out = ops.forward(d(inp), d(grid), d(off), pad, align, kernel, multicell)
        assert out.shape == gOut.shape and out.is_contiguous()
        assert_close_scaled(out, so.forward(inp, grid, off, **kw), "F out")
        # B
        gI, gG = ops.backward(d(gOut), d(inp), d(grid), d(off), pad, align, True, kernel, multicell)
        r_gI, r_gG = so.backward(gOut, inp, grid, off, input_requires_grad=True, **kw)
        assert gI.shape == inp.shape and gI.is_contiguous() and gG.shape == grid.shape
        assert_close_scaled(gI, r_gI, "B gInput")
        assert_close_scaled(gG, r_gG, "B gGrid")
        gI2, gG2 = ops.backward(d(gOut), d(inp), d(grid), d(off), pad, align, False, kernel, multicell)
        assert gI2 is None
        assert_close_scaled(gG2, r_gG, "B gGrid (no gInput)")
        # BB without / with gOutInput
        for use in (False, True):
            res = ops.backward_backward(d(gOI) if use else None, d(gOG), d(inp), d(grid), d(gOut), d(off),
                                        pad, align, use, kernel, multicell)
            ref = so.backward_backward(gOI if use else None, gOG, inp, grid, gOut, off,
                                       input_requires_grad=use, **kw)
            for name, a, b in zip(("gInput", "gGrid", "ggOut"), res, ref):
                assert_close_scaled(a, b, "BB %s (gOutInput=%s)" % (name, use))
        # BB elided outputs
        only = ops.backward_backward(None, d(gOG), d(inp), d(grid), d(gOut), d(off), pad, align, False,
                                     kernel, multicell, want=(True, False, False))
        assert only[1] is None and only[2] is None
        ref = so.backward_backward(None, gOG, inp, grid, gOut, off, input_requires_grad=False, **kw)
        assert_close_scaled(only[0], ref[0], "BB gInput only")
        # BBB
        res = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad, align,
                                             False, kernel, multicell)
        ref3 = so.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, **kw)
        assert_close_scaled(res[0], ref3[0], "BBB gInput")
        assert_close_scaled(res[1], ref3[1], "BBB ggOut")
        # BBB fused with the b_input pass (mod2d:106-111): gInput + b_input
        fused = ops.backward_backward_backward(d(inp), d(grid), d(gOut), d(gOG), d(gOgG), d(off), pad,
                                               align, False, kernel, multicell, gOutggOut=d(gOggO))
        b_input = so.backward_backward(None, gOG, inp, grid, gOggO, off, input_requires_grad=False, **kw)[0]
        assert_close_scaled(fused[0], ref3[0] + b_input, "BBB fused gInput + b_input")
        assert_close_scaled(fused[1], ref3[1], "BBB fused ggOut")
    finally:
        ops.set_index_mode("separate")
        ops.set_small_cell("auto")


# both kernels of the library: the cp.async-pipelined global-memory kernel ("never") and the
# shared-memory small-cell kernel with privatised scatter ("always")
PATHS = ["never", "always"]


@pytest.mark.parametrize("small_cell", PATHS)
@pytest.mark.parametrize("multicell", [True, False])
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("dim", [2, 3])
def test_all_stages_channel_last_c16(cuda, dim, kernel, multicell, small_cell):
    _run_all(cuda, dim, N=3, C=16, P=1000, kernel=kernel, multicell=multicell, small_cell=small_cell)


@pytest.mark.parametrize("small_cell", PATHS)
@pytest.mark.parametrize("C,P", [(4, 257), (8, 512), (32, 130), (12, 64), (6, 100), (1, 33), (64, 40), (20, 7)])
@pytest.mark.parametrize("dim", [2, 3])
def test_channel_counts_and_ragged_point_counts(cuda, dim, C, P, small_cell):
    """C % 4 == 0 takes the channel-last vector path with 1/2/4/8 lanes per quad; other C the
    scalar channel-first path; P % 4 != 0 the scalar stream path; tiles end ragged."""
    _run_all(cuda, dim, N=2, C=C, P=P, kernel=0, multicell=True, seed=C * 1000 + P, small_cell=small_cell)


@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
def test_lane_override(cuda, lanes):
    from cosinesampler_b200 import ops
    ops.set_lanes(lanes)
    try:
        _run_all(cuda, 2, N=2, C=16, P=300, kernel=0, multicell=True, seed=lanes)
        _run_all(cuda, 3, N=2, C=8, P=300, kernel=2, multicell=True, seed=lanes)
    finally:
        ops.set_lanes(0)


@pytest.mark.parametrize("small_cell", PATHS)
@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("pad", [0, 1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_padding_modes_and_align_corners_out_of_range_points(cuda, dim, pad, align, small_cell):
    """coordinates in [-1.4, 1.4]: zeros padding drops out-of-range corners, border clips,
    reflection reflects (over [0,S-2] when align_corners, cu2d:184-188)."""
    for kernel, multicell in itertools.product((0, 2, 1), (True, False)):
        _run_all(cuda, dim, N=2, C=8, P=400, kernel=kernel, multicell=multicell, pad=pad, align=align,
                 lo=-1.4, hi=1.4, seed=17 + pad, small_cell=small_cell)


def test_contended_scatter_on_a_tiny_cell(cuda):
    """The reference's own shapes (test_2d.py:26-38): many points on 16x16 cells, so thousands of
    contributions land on each texel; both scatter paths must agree with the oracle."""
    for path in PATHS:
        _run_all(cuda, 2, N=6, C=4, P=20000, kernel=0, multicell=True, sizes=(16, 16), seed=3, small_cell=path)
        _run_all(cuda, 3, N=3, C=4, P=20000, kernel=0, multicell=True, sizes=(8, 8, 8), seed=4, small_cell=path)


def test_fused_index_mode(cuda):
    """index_mode 'fused' reproduces the single-rounding fma of the reference's fast-math build."""
    _run_all(cuda, 2, N=4, C=16, P=2000, kernel=0, multicell=True, index_mode="fused",
             sizes=(64, 64))
    _run_all(cuda, 3, N=4, C=16, P=2000, kernel=2, multicell=True, index_mode="fused",
             sizes=(16, 16, 16))


def test_expanded_gout_and_grid_are_read_in_place(cuda):
    """PIXEL's val.sum(0) produces a gOut expanded over the cell axis (stride 0); the reference
    copies it (mod2d:42).  Same numbers either way."""
    from cosinesampler_b200 import ops
    inp, grid, gOut, gOG, *_ = _inputs(2, 4, 16, 512, seed=5)
    off = _offset(4, True).to(cuda)
    g1 = gOut[:1].to(cuda)
    gexp = g1.expand(4, -1, -1, -1)
    assert gexp.stride(0) == 0
    grid1 = grid[:1].to(cuda)
    grid_exp = grid1.expand(4, -1, -1, -1)
    gI_a, gG_a = ops.backward(gexp, inp.to(cuda), grid_exp, off, 0, True, True, 0, True)
    gI_b, gG_b = ops.backward(gexp.contiguous(), inp.to(cuda), grid_exp.contiguous(), off, 0, True, True, 0, True)
    assert torch.equal(gG_a, gG_b)
    assert_close_scaled(gI_a, gI_b, "gInput, expanded vs materialised gOut")
    r_gI, r_gG = so.backward(gexp.cpu(), inp, grid_exp.cpu(), off.cpu(), input_requires_grad=True,
                             pad=0, align=True, kernel=0, multicell=True)
    assert_close_scaled(gG_a, r_gG, "gGrid vs oracle")
    assert_close_scaled(gI_a, r_gI, "gInput vs oracle")


def test_empty_and_degenerate_inputs(cuda):
    from cosinesampler_b200 import ops
    off = _offset(2, True).to(cuda)
    inp = torch.rand(2, 8, 5, 5, device=cuda)
    empty_grid = torch.zeros(2, 1, 0, 2, device=cuda)
    out = ops.forward(inp, empty_grid, off, 0, True, 0, True)
    assert out.shape == (2, 8, 1, 0)
    gI, gG = ops.backward(torch.zeros(2, 8, 1, 0, device=cuda), inp, empty_grid, off, 0, True, True, 0, True)
    assert gG.shape == empty_grid.shape and float(gI.abs().sum()) == 0.0
    # non-finite coordinates must not fault: the reference leaves them undefined (cu2d:310-311)
    grid = torch.tensor([[[[float("nan"), 0.0], [float("inf"), 0.2], [0.1, -float("inf")], [0.3, 0.3]]]],
                        device=cuda).repeat(2, 1, 1, 1)
    out = ops.forward(inp, grid, off, 0, True, 0, True)
    torch.cuda.synchronize()
    assert torch.isfinite(out[..., 3]).all() and float(out[..., :3].abs().sum()) == 0.0
    # huge finite coordinates: border padding clips them to the edge texel (cu2d:90-93), zeros drops them
    far = torch.tensor([[[[1e20, -1e20], [3.0, -3.0], [-1e20, 0.5], [0.25, 1e20]]]], device=cuda).repeat(2, 1, 1, 1)
    for pad in (0, 1):
        o = ops.forward(inp, far, off, pad, True, 0, True)
        ref = so.forward(inp.cpu(), far.cpu(), off.cpu(), pad=pad, kernel=0, multicell=True)
        assert_close_scaled(o, ref, "huge coordinates, padding %d" % pad)
    # a 2x2 cell with multicell: index range collapses to a single interval
    tiny = torch.rand(2, 4, 2, 2, device=cuda)
    g = torch.rand(2, 1, 16, 2, device=cuda) * 2 - 1
    o = ops.forward(tiny, g, off, 0, True, 0, True)
    ref = so.forward(tiny.cpu(), g.cpu(), off.cpu(), kernel=0, multicell=True)
    assert_close_scaled(o, ref, "2x2 cell")


def test_non_contiguous_and_wrong_device_inputs_raise(cuda):
    from cosinesampler_b200 import ops
    off = _offset(2, True).to(cuda)
    inp = torch.rand(2, 8, 6, 12, device=cuda)[..., ::2]
    grid = torch.rand(2, 1, 8, 2, device=cuda)
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.forward(inp, grid, off, 0, True, 0, True)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.forward(inp.contiguous(), grid.cpu(), off, 0, True, 0, True)
    with pytest.raises(RuntimeError, match="float32"):
        ops.forward(inp.contiguous().double(), grid, off, 0, True, 0, True)


def test_layout_staging_round_trip(cuda):
    from cosinesampler_b200 import ops
    for shape in [(3, 16, 7, 9), (2, 5, 4, 4, 5), (1, 40, 33, 2), (2, 4, 1, 1)]:
        x = torch.rand(shape, device=cuda)
        cl = ops.to_channel_last(x)
        N, C = shape[:2]
        assert torch.equal(cl, x.reshape(N, C, -1).transpose(1, 2).contiguous())
        back = ops.from_channel_last(cl, shape)
        assert torch.equal(back, x)
        acc = torch.ones(shape, device=cuda)
        ops.from_channel_last(cl, shape, out=acc, accumulate=True)
        assert torch.equal(acc, x + 1)
