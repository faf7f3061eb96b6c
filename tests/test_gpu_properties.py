"""Size-independent properties at BASELINE.json's full sizes (configs 3 and 4), where the
CPU oracle would take minutes: partition of unity, linearity, the adjoint identity
<F(x), g> == <x, B(g)>, directional finite differences for the coordinate gradients, and
agreement between the vector (channel-last) and scalar (channel-first) code paths."""
import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu

CONFIGS = [
    ("cfg3", 2, (4, 16, 256, 256), 2 ** 20, "cosine"),
    ("cfg4", 3, (4, 16, 64, 64, 64), 2 ** 22, "smooth-step"),
]


def _sampler(dim):
    if dim == 2:
        from cosine_sampler_2d import CosineSampler2d as S
    else:
        from cosine_sampler_3d import CosineSampler3d as S
    return S


@pytest.mark.parametrize("name,dim,shape,P,kernel", CONFIGS)
def test_full_size_properties(cuda, name, dim, shape, P, kernel):
    S = _sampler(dim)
    torch.manual_seed(0)
    N, C = shape[:2]
    cells = torch.rand(shape, device=cuda)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    pts = torch.rand((1,) + gshape[1:], device=cuda) * 2 - 1
    grid = pts.repeat((N,) + (1,) * (len(gshape) - 1))
    f = lambda c, g=grid: S.apply(c, g, "zeros", True, kernel, True)

    # partition of unity: a constant field samples to the same constant
    ones = torch.ones(shape, device=cuda)
    out1 = f(ones)
    assert float((out1 - 1).abs().max()) < 2e-6

    # linearity in the field
    a, b = cells, torch.rand(shape, device=cuda)
    lhs = f(2.5 * a - 0.5 * b)
    rhs = 2.5 * f(a) - 0.5 * f(b)
    assert_close_scaled(lhs, rhs, "linearity", rtol=1e-5, atol_scale=1e-6)

    # range: an interpolant of values in [0,1] stays in [0,1]
    out = f(cells)
    assert float(out.min()) >= -1e-6 and float(out.max()) <= 1 + 1e-6

    # adjoint identity between F and the gInput half of B
    g = torch.randn(out.shape, device=cuda)
    c = cells.clone().requires_grad_(True)
    gI = torch.autograd.grad(f(c), c, g)[0]
    lhs = (out.double() * g.double()).sum()
    rhs = (cells.double() * gI.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * float((out.double() * g.double()).abs().sum())
    # total mass: each pair spreads exactly gOut over its corners
    assert abs(float(gI.double().sum() - g.double().sum())) <= 1e-5 * float(g.double().abs().sum())

    # coordinate gradient against a central finite difference along a random direction,
    # on a subset (first 2^16 points), in the interior of cells
    sub = 2 ** 16
    gsub = grid[..., :sub, :].contiguous().requires_grad_(True)
    osub = S.apply(cells, gsub, "zeros", True, kernel, True)
    w = torch.randn(osub.shape, device=cuda)
    gG = torch.autograd.grad(osub, gsub, w)[0]
    d = torch.randn(gsub.shape, device=cuda)
    h = 1e-3 / max(shape[2:])
    with torch.no_grad():
        fp = S.apply(cells, gsub + h * d, "zeros", True, kernel, True)
        fm = S.apply(cells, gsub - h * d, "zeros", True, kernel, True)
        fd = ((fp.double() - fm.double()) * w.double()).sum() / (2 * h)
        an = (gG.double() * d.double()).sum()
    assert abs(float(fd - an)) <= 2e-2 * abs(float(an)) + 1e-3 * float((gG.double() * d.double()).abs().sum()) / sub ** 0.5


@pytest.mark.parametrize("dim,shape,P", [(2, (4, 16, 64, 64), 2 ** 16), (3, (2, 8, 16, 16, 16), 2 ** 15)])
def test_vector_and_scalar_paths_agree(cuda, dim, shape, P):
    """channel-last/red.v4 path vs channel-first scalar path of the same library"""
    from cosinesampler_b200 import ops, _lib
    from cosinesampler_b200.autograd import cell_offsets
    torch.manual_seed(1)
    N, C = shape[:2]
    inp = torch.rand(shape, device=cuda)
    gshape = (N, 1, P, 2) if dim == 2 else (N, 1, 1, P, 3)
    grid = torch.rand(gshape, device=cuda) * 2 - 1
    gOut = torch.randn((N, C) + gshape[1:-1], device=cuda)
    gOG = torch.randn(gshape, device=cuda)
    off = cell_offsets(N, True, cuda)
    fast = ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, 0, True)
    real = ops.uses_channel_last
    ops.uses_channel_last = lambda C: False
    try:
        slow = ops.backward_backward(None, gOG, inp, grid, gOut, off, 0, True, False, 0, True)
    finally:
        ops.uses_channel_last = real
    for nm, a, b in zip(("gInput", "gGrid", "ggOut"), fast, slow):
        assert_close_scaled(a, b, "vector vs scalar path: " + nm)
