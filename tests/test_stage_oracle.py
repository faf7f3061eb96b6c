"""Pin oracle/stage_oracle.py (what the four CUDA entry points compute) to PyTorch autograd
over oracle/grid_sampler_oracle.py (itself bit-equal to the real reference), wherever the
reference CUDA kernels are mathematically exact:

  F    == sampler value
  B    == d<out,gOut>/d(input, grid)
  BB   == d(<gInput,gOI> + <gGrid,gOG>)/d(input, grid, gOut); 2D gGrid only on the axis where
          gOG is hot (the 2D kernel drops mixed terms, cu2d:675-678,705-706)
  BBB  == d<gGrid', gOgG>/d(input, gOut) with gOG, gOgG hot on the same axis
          (pure second derivatives only, cu2d:876-885 / cu3d:1054-1065)
"""
import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import cell_offsets, grid_sample_2d, grid_sample_3d
from util import safe_coords

KERNELS = [("cosine", so.K_COSINE), ("smoothstep", so.K_SMOOTHSTEP), ("linear", so.K_LINEAR)]


def _setup(dim, multicell, seed=0, P=40, N=3, C=2):
    gen = torch.Generator().manual_seed(seed)
    shape = (N, C, 6, 7) if dim == 2 else (N, C, 5, 5, 5)
    sizes = [shape[-1 - a] for a in range(dim)]
    inp = torch.rand(shape, generator=gen, dtype=torch.float64)
    pts = safe_coords(P, dim, sizes, N, multicell, gen).float().double()   # fp32-representable
    grid = pts.reshape((1,) * dim + (P, dim)).repeat((N,) + (1,) * (dim + 1)).contiguous()
    off = cell_offsets(N, multicell, dtype=torch.float32)
    return inp, grid, off, gen


def _sampler(dim, name, multicell):
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    step = {"linear": "bilinear" if dim == 2 else "trilinear"}.get(name, name)
    return lambda c, g: fn(c, g, step=step, offset=multicell)


@pytest.mark.parametrize("multicell", [True, False])
@pytest.mark.parametrize("name,kcode", KERNELS)
@pytest.mark.parametrize("dim", [2, 3])
def test_stage_oracle_matches_autograd(dim, name, kcode, multicell):
    inp, grid, off, gen = _setup(dim, multicell)
    N, C = inp.shape[:2]
    P = grid.shape[-2]
    sampler = _sampler(dim, name, multicell)
    inp_r = inp.clone().requires_grad_(True)
    grid_r = grid.clone().requires_grad_(True)
    out = sampler(inp_r, grid_r)
    kw = dict(pad=0, align=True, kernel=kcode, multicell=multicell, index_mode=2)

    # F
    f = so.forward(inp, grid, off, **kw)
    torch.testing.assert_close(f.reshape(out.shape), out.detach(), rtol=1e-10, atol=1e-12)

    # B
    gOut = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    gI_t, gG_t = torch.autograd.grad((out * gOut).sum(), [inp_r, grid_r], create_graph=True)
    gI, gG = so.backward(gOut, inp, grid, off, input_requires_grad=True, **kw)
    torch.testing.assert_close(gI, gI_t.detach(), rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(gG, gG_t.detach(), rtol=1e-10, atol=1e-11)

    # BB: gOG hot on one axis per point
    gOut_r = gOut.clone().requires_grad_(True)
    out2 = sampler(inp_r, grid_r)
    gI_t, gG_t = torch.autograd.grad((out2 * gOut_r).sum(), [inp_r, grid_r], create_graph=True)
    hot = torch.randint(0, dim, (P,), generator=gen)
    onehot = torch.nn.functional.one_hot(hot, dim).double()
    gOG = (torch.randn(N, P, 1, generator=gen, dtype=torch.float64) * onehot).reshape(grid.shape)
    gOI = torch.randn(inp.shape, generator=gen, dtype=torch.float64)
    for use_goi in (False, True):
        L = (gG_t * gOG).sum() + ((gI_t * gOI).sum() if use_goi else 0.0)
        t_gI, t_gG, t_ggO = torch.autograd.grad(L, [inp_r, grid_r, gOut_r], create_graph=True)
        s_gI, s_gG, s_ggO = so.backward_backward(gOI if use_goi else None, gOG, inp, grid, gOut, off,
                                                 input_requires_grad=use_goi, **kw)
        torch.testing.assert_close(s_gI, t_gI.detach(), rtol=1e-9, atol=1e-10)
        torch.testing.assert_close(s_ggO, t_ggO.detach(), rtol=1e-9, atol=1e-10)
        if dim == 3:
            torch.testing.assert_close(s_gG, t_gG.detach(), rtol=1e-9, atol=1e-9)
        elif not use_goi:
            m = onehot.reshape((1,) * dim + (P, dim)).expand_as(s_gG)
            torch.testing.assert_close(s_gG * m, t_gG.detach() * m, rtol=1e-9, atol=1e-9)
            assert (s_gG * (1 - m)).abs().max() == 0     # mixed terms dropped by the 2D kernel

    # BBB: gOgG hot on the same axis as gOG
    L = (gG_t * gOG).sum()
    t_gI, t_gG, t_ggO = torch.autograd.grad(L, [inp_r, grid_r, gOut_r], create_graph=True)
    gOgG = (torch.randn(N, P, 1, generator=gen, dtype=torch.float64) * onehot).reshape(grid.shape)
    M = (t_gG * gOgG).sum()
    if M.requires_grad and name != "linear":
        u_gI, u_ggO = torch.autograd.grad(M, [inp_r, gOut_r], allow_unused=True)
        s_gI, s_ggO = so.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, **kw)
        torch.testing.assert_close(s_gI, u_gI, rtol=1e-8, atol=1e-8)
        torch.testing.assert_close(s_ggO, u_ggO, rtol=1e-8, atol=1e-8)
    else:
        s_gI, s_ggO = so.backward_backward_backward(inp, grid, gOut, gOG, gOgG, off, **kw)
        assert s_gI.abs().max() == 0 and s_ggO.abs().max() == 0   # linear: k'' = 0


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("pad,pname", [(so.PAD_ZEROS, "zeros"), (so.PAD_BORDER, "border")])
def test_linear_stage_oracle_equals_grid_sample_with_padding(pad, pname, align):
    """zeros / border padding and align_corners follow ATen for the linear kernel without
    multicell (reflection deliberately does not: cu2d:184-188 reflects over [0, S-2]).
    The 2D forward ignores align_corners (cu2d:307-308), so 2D is checked with align=True only."""
    gen = torch.Generator().manual_seed(5)
    for dim in (2, 3):
        if dim == 2 and not align:
            continue
        shape = (2, 3, 6, 7) if dim == 2 else (2, 3, 4, 5, 6)
        inp = torch.rand(shape, generator=gen, dtype=torch.float64)
        P = 64
        grid = ((torch.rand((2,) + (1,) * (dim - 1) + (P, dim), generator=gen) * 2.6 - 1.3)).double()
        off = torch.zeros(2)
        ref_in = inp.clone().requires_grad_(True)
        ref_g = grid.clone().requires_grad_(True)
        ref = torch.nn.functional.grid_sample(ref_in, ref_g, mode="bilinear", padding_mode=pname,
                                              align_corners=align)
        out = so.forward(inp, grid, off, pad=pad, align=align, kernel=so.K_LINEAR, multicell=False, index_mode=2)
        torch.testing.assert_close(out.reshape(ref.shape), ref.detach(), rtol=1e-6, atol=1e-6)
        gOut = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
        rI, rG = torch.autograd.grad((ref * gOut).sum(), [ref_in, ref_g])
        gI, gG = so.backward(gOut, inp, grid, off, pad=pad, align=align, kernel=so.K_LINEAR,
                             multicell=False, index_mode=2)
        torch.testing.assert_close(gI, rI, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(gG, rG, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------
# jet oracle (cosinesampler_b200/jet.py) pinned to the four stage oracles above
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("multicell", [True, False])
@pytest.mark.parametrize("name,kcode", KERNELS)
@pytest.mark.parametrize("dim", [2, 3])
def test_jet_oracle_matches_stage_oracles(dim, name, kcode, multicell):
    inp, grid, off, gen = _setup(dim, multicell, seed=3, C=4)
    N, C = inp.shape[:2]
    P = grid.shape[-2]
    coords = grid[0].reshape(P, dim)
    kw = dict(pad=0, align=True, kernel=kcode, multicell=multicell, index_mode=2)
    jets = so.jet_forward(inp, coords, off, order=2, **kw)
    assert jets.shape == (1 + 2 * dim, C, P)
    close = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-11)

    # value = F summed over the cells
    close(jets[0], so.forward(inp, grid, off, **kw).reshape(N, C, P).sum(0))
    gOut = torch.randn(C, P, generator=gen, dtype=torch.float64)
    gOutN = gOut.reshape((1, C) + (1,) * (dim - 1) + (P,)).repeat((N,) + (1,) * (dim + 1))
    # first derivatives: B's gGrid is the contraction of z_a with gOut over the channels
    _, gG = so.backward(gOutN, inp, grid, off, input_requires_grad=False, **kw)
    for a in range(dim):
        close((jets[1 + a] * gOut).sum(0), gG.reshape(N, P, dim)[..., a].sum(0))
    # pure second derivatives: BB's gGrid with gOutGrid hot on axis a
    for a in range(dim):
        hot = torch.zeros(grid.shape, dtype=torch.float64)
        hot[..., a] = 1.0
        _, gG2, ggO = so.backward_backward(None, hot, inp, grid, gOutN, off, input_requires_grad=False, **kw)
        close((jets[1 + dim + a] * gOut).sum(0), gG2.reshape(N, P, dim)[..., a].sum(0))
        close(jets[1 + a], ggO.reshape(N, C, P).sum(0))          # ggOut of BB = z_a

    # adjoint: each jet's share of jet_backward is the gInput of the matching stage
    G = torch.randn(jets.shape, generator=gen, dtype=torch.float64)
    gI = so.jet_backward(G, inp.shape, coords, off, order=2, **kw)
    rep = lambda t: t.reshape((1, C) + (1,) * (dim - 1) + (P,)).repeat((N,) + (1,) * (dim + 1))
    want = so.backward(rep(G[0]), inp, grid, off, input_requires_grad=True, **kw)[0]
    for a in range(dim):
        hot = torch.zeros(grid.shape, dtype=torch.float64)
        hot[..., a] = 1.0
        want = want + so.backward_backward(None, hot, inp, grid, rep(G[1 + a]), off, **kw)[0]
        want = want + so.backward_backward_backward(inp, grid, rep(G[1 + dim + a]), hot, hot, off, **kw)[0]
    close(gI, want)
    # <jets(V), G> == <V, jet_backward(G)>
    close((jets * G).sum(), (inp * gI).sum())
    # order 1 is a prefix of order 2
    close(so.jet_forward(inp, coords, off, order=1, **kw), jets[:1 + dim])


@pytest.mark.parametrize("name,kcode", KERNELS)
def test_mixed_jets_match_the_3d_double_backward(name, kcode):
    """order 3: the mixed second derivatives are what the reference's 3D double backward contracts
    (cu3d:836-856): BB's gGrid_b with gOutGrid hot on axis a is <z_ab, gOut> for b != a; order 2 is a prefix;
    and the adjoint identity <jets(V), G> = <V, jet_backward(G)> holds with the mixed rows included."""
    dim = 3
    inp, grid, off, gen = _setup(dim, True, seed=5, C=4)
    N, C = inp.shape[:2]
    P = grid.shape[-2]
    coords = grid[0].reshape(P, dim)
    kw = dict(pad=0, align=True, kernel=kcode, multicell=True, index_mode=2)
    jets = so.jet_forward(inp, coords, off, order=3, **kw)
    assert jets.shape == (so.jet_count(dim, 3), C, P) == (10, C, P)
    close = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-11)
    close(jets[:1 + 2 * dim], so.jet_forward(inp, coords, off, order=2, **kw))
    gOut = torch.randn(C, P, generator=gen, dtype=torch.float64)
    gOutN = gOut.reshape((1, C) + (1,) * (dim - 1) + (P,)).repeat((N,) + (1,) * (dim + 1))
    for m, (a, b) in enumerate(so.mixed_pairs(dim)):
        hot = torch.zeros(grid.shape, dtype=torch.float64)
        hot[..., a] = 1.0
        _, gG2, _ = so.backward_backward(None, hot, inp, grid, gOutN, off, input_requires_grad=False, **kw)
        close((jets[1 + 2 * dim + m] * gOut).sum(0), gG2.reshape(N, P, dim)[..., b].sum(0))
    G = torch.randn(jets.shape, generator=gen, dtype=torch.float64)
    gI = so.jet_backward(G, inp.shape, coords, off, order=3, **kw)
    close((jets * G).sum(), (inp * gI).sum())
    # 2D has one mixed jet
    inp2, grid2, off2, _ = _setup(2, True, seed=6, C=4)
    j2 = so.jet_forward(inp2, grid2[0].reshape(-1, 2), off2, order=3, **kw)
    assert j2.shape[0] == so.jet_count(2, 3) == 6
