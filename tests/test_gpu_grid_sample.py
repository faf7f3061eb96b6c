"""BASELINE.json config 2: linear kernel, multicell off == torch.nn.functional.grid_sample
(README.md:26-27 of the reference claims it; the reference never tests it).

Forward must be bit-identical to ATen's compiled kernel (same weights (ix - ix_nw) etc., same
fma accumulation order nw, ne, sw, se).  gGrid is a sum over channels whose order differs
(channels are split over lanes), gInput is accumulated with atomics in both implementations:
both are compared with the fp32 tolerance and the fraction of bit-equal elements is reported."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close_scaled

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,shape,P", [(2, (1, 32, 512, 512), 2 ** 20), (2, (2, 16, 37, 53), 10007),
                                         (3, (1, 16, 32, 32, 32), 2 ** 17), (3, (2, 8, 9, 10, 11), 4099)])
def test_linear_no_multicell_equals_grid_sample(cuda, dim, shape, P):
    if dim == 2:
        from cosine_sampler_2d import CosineSampler2d as S
        name = "bilinear"
    else:
        from cosine_sampler_3d import CosineSampler3d as S
        name = "trilinear"
    gen = torch.Generator().manual_seed(7)
    N = shape[0]
    inp = torch.rand(shape, generator=gen).to(cuda).requires_grad_(True)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid = (torch.rand(gshape, generator=gen) * 2 - 1).to(cuda).requires_grad_(True)
    out = S.apply(inp, grid, "zeros", True, name, False)
    # F.grid_sample hands 4-D bilinear/zeros/align_corners=True inputs to cuDNN when it can;
    # ATen's own kernel (GridSampler.cu, the one the reference is derived from, cu2d:1-3)
    # is what bit-equality is defined against.  The cuDNN result is checked with the tolerance.
    with torch.backends.cudnn.flags(enabled=False):
        ref = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.equal(out, ref), "forward must be bit-identical to ATen's grid_sample kernel"
    ref_cudnn = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert_close_scaled(out, ref_cudnn, "forward vs F.grid_sample (cuDNN path when taken)")
    gOut = torch.randn(ref.shape, generator=gen).to(cuda)
    gI, gG = torch.autograd.grad(out, [inp, grid], gOut)
    rI, rG = torch.autograd.grad(ref, [inp, grid], gOut)
    assert_close_scaled(gG, rG, "gGrid vs grid_sample")
    assert_close_scaled(gI, rI, "gInput vs grid_sample")
    frac = float((gG == rG).float().mean())
    print("gGrid bit-equal fraction (fast order): %.4f" % frac)
    # reference order: channels in sequence, ATen's products and fmas -> bit-identical gGrid
    from cosinesampler_b200 import ops
    ops.set_grad_order("reference")
    try:
        gG2 = torch.autograd.grad(S.apply(inp, grid, "zeros", True, name, False), grid, gOut)[0]
    finally:
        ops.set_grad_order("fast")
    frac2 = float((gG2 == rG).float().mean())
    print("gGrid bit-equal fraction (reference order): %.6f" % frac2)
    assert torch.equal(gG2, rG), "reference-order gGrid must be bit-identical to ATen (%.6f equal)" % frac2


def test_zero_padding_out_of_range_equals_grid_sample(cuda):
    from cosine_sampler_2d import CosineSampler2d as S
    gen = torch.Generator().manual_seed(8)
    inp = torch.rand(2, 8, 16, 16, generator=gen).to(cuda)
    grid = (torch.rand(2, 1, 4096, 2, generator=gen) * 3 - 1.5).to(cuda)
    out = S.apply(inp, grid, "zeros", True, "bilinear", False)
    with torch.backends.cudnn.flags(enabled=False):
        ref = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.equal(out, ref)
    outb = S.apply(inp, grid, "border", True, "bilinear", False)
    refb = F.grid_sample(inp, grid, mode="bilinear", padding_mode="border", align_corners=True)
    assert_close_scaled(outb, refb, "border padding")
