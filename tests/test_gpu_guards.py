"""compute-sanitizer is closed on this GPU pool, so out-of-bounds *writes* are checked the
hard way: every output of every C-ABI entry point is placed inside a larger buffer filled with a
canary pattern, the kernels run on ragged / tiny / out-of-range problems, and the canaries on both
sides of each output must be untouched (and the outputs must match the regular path)."""
import ctypes

import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu
CANARY = -1234.5
PAD = 4096          # floats of guard on each side


MISALIGN = 0        # set to 1 by the unaligned test: outputs start 4 bytes off a 16-byte boundary


class Guarded:
    def __init__(self, shape, device, zero=False):
        self.n = 1
        for s in shape:
            self.n *= s
        self.shape = shape
        self.lo = PAD + MISALIGN
        self.buf = torch.full((self.n + 2 * PAD + 8,), CANARY, device=device)
        self.view = self.buf[self.lo:self.lo + self.n]
        if zero:
            self.view.zero_()

    def ptr(self):
        return self.view.data_ptr()

    def tensor(self):
        return self.view.view(self.shape)

    def intact(self):
        lo = self.buf[:self.lo]
        hi = self.buf[self.lo + self.n:]
        return bool((lo == CANARY).all()) and bool((hi == CANARY).all())


CASES = [
    # dim, N, C, sizes, P, coordinate range, small_cell
    (2, 3, 16, (9, 12), 1001, 1.5, 1), (2, 3, 16, (9, 12), 1001, 1.5, 2),
    (2, 2, 4, (16, 16), 777, 1.0, 1), (2, 2, 4, (16, 16), 777, 1.0, 2),
    (2, 2, 6, (5, 7), 130, 1.3, 1),
    (3, 2, 16, (6, 7, 8), 515, 1.4, 1), (3, 2, 16, (6, 7, 8), 515, 1.4, 2),
    (3, 2, 8, (4, 4, 4), 64, 1.0, 1), (3, 1, 5, (3, 4, 5), 33, 1.2, 1),
]


@pytest.mark.parametrize("dim,N,C,sizes,P,rng,small", CASES)
def test_no_output_is_written_out_of_bounds(cuda, dim, N, C, sizes, P, rng, small):
    from cosinesampler_b200 import _lib, ops
    from cosinesampler_b200.autograd import cell_offsets
    lib = _lib.load()
    torch.manual_seed(P)
    inp = torch.rand((N, C) + sizes, device=cuda)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid = (torch.rand(gshape, device=cuda) * 2 - 1) * rng
    sshape = (N, C) + gshape[1:-1]
    gOut = torch.randn(sshape, device=cuda)
    gOut2 = torch.randn(sshape, device=cuda)
    gOG = torch.randn(gshape, device=cuda)
    gOgG = torch.randn(gshape, device=cuda)
    gOI = torch.randn(inp.shape, device=cuda)
    off = cell_offsets(N, True, cuda)
    T = 1
    for s in sizes:
        T *= s
    cl = ops.uses_channel_last(C)
    field = ops.to_channel_last(inp) if cl else inp
    goi = ops.to_channel_last(gOI) if cl else gOI
    fshape = (N, T, C) if cl else inp.shape

    pb = _lib.Problem()
    pb.dim, pb.N, pb.C = dim, N, C
    pb.D, pb.H, pb.W = (1,) + sizes if dim == 2 else sizes
    pb.P = P
    pb.padding_mode, pb.align_corners, pb.kernel, pb.multicell = 0, 1, 0, 1
    pb.index_mode = 0
    pb.field_layout = _lib.LAYOUT_CHANNEL_LAST if cl else _lib.LAYOUT_CHANNEL_FIRST
    pb.grid_stride_n = P * dim
    pb.lanes, pb.small_cell = 0, small
    stream = torch.cuda.current_stream().cuda_stream
    s1 = _lib.Stream3(gOut.data_ptr(), gOut.stride(0), gOut.stride(1))
    s2 = _lib.Stream3(gOut2.data_ptr(), gOut2.stride(0), gOut2.stride(1))
    outs = []

    out = Guarded(sshape, cuda)
    _lib.check(lib.cs_forward(pb, field.data_ptr(), grid.data_ptr(), off.data_ptr(), out.ptr(), stream), "F")
    outs.append(out)

    gI, gG = Guarded(fshape, cuda, zero=True), Guarded(gshape, cuda)
    _lib.check(lib.cs_backward(pb, s1, field.data_ptr(), grid.data_ptr(), off.data_ptr(), gI.ptr(), gG.ptr(),
                               stream), "B")
    outs += [gI, gG]

    bI, bG, bO = Guarded(fshape, cuda, zero=True), Guarded(gshape, cuda), Guarded(sshape, cuda)
    _lib.check(lib.cs_backward_backward(pb, goi.data_ptr(), gOG.data_ptr(), field.data_ptr(), grid.data_ptr(), s1,
                                        off.data_ptr(), bI.ptr(), bG.ptr(), bO.ptr(), stream), "BB")
    outs += [bI, bG, bO]

    cI, cO = Guarded(fshape, cuda, zero=True), Guarded(sshape, cuda)
    _lib.check(lib.cs_backward_backward_backward(pb, field.data_ptr(), grid.data_ptr(), s1, gOG.data_ptr(),
                                                 gOgG.data_ptr(), s2, off.data_ptr(), cI.ptr(), cO.ptr(), stream),
               "BBB")
    outs += [cI, cO]
    torch.cuda.synchronize()
    for g in outs:
        assert g.intact(), "canary overwritten next to an output of shape %s" % (g.shape,)
        assert torch.isfinite(g.tensor()).all()
        assert not bool((g.tensor() == CANARY).all())

    # same numbers as the regular host path
    ops.set_small_cell({1: "never", 2: "always"}[small])
    try:
        ref = ops.forward(inp, grid, off, 0, True, 0, True)
        r = ops.backward_backward(gOI, gOG, inp, grid, gOut, off, 0, True, True, 0, True)
        if MISALIGN:     # unaligned fields run the scalar kernel: another channel summation order
            assert_close_scaled(out.tensor(), ref, "F")
            assert_close_scaled(bG.tensor(), r[1], "BB gGrid")
            assert_close_scaled(bO.tensor(), r[2], "BB ggOut")
        else:
            assert torch.equal(out.tensor(), ref)
            assert torch.equal(bG.tensor(), r[1]) and torch.equal(bO.tensor(), r[2])
    finally:
        ops.set_small_cell("auto")


def test_unaligned_outputs_take_the_scalar_paths_without_overrun(cuda):
    """outputs that start 4 bytes off a 16-byte boundary: stream / gGrid stores fall back to scalar
    code, the accumulator to scalar reds"""
    global MISALIGN
    MISALIGN = 1
    try:
        test_no_output_is_written_out_of_bounds(cuda, 2, 3, 16, (9, 12), 1001, 1.5, 1)
        test_no_output_is_written_out_of_bounds(cuda, 3, 2, 16, (6, 7, 8), 515, 1.4, 1)
    finally:
        MISALIGN = 0
