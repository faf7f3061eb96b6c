"""CPU checks of the fused path's host-side pieces (no kernel runs here):

  * `jet.jet_mlp` -- the head's chain rule applied to jets (second-order Taylor mode, plain torch
    ops) -- fed with jets from the jet oracle reproduces u, u_a, u_aa of the reference's nested
    autograd chain (`oracle.grid_sampler_oracle.derivative_chain`, test_2d.py:55-127) in fp64;
  * the residual table handed to cs_pde_head_step encodes the residuals of `chain.pde_loss`;
  * a whole jet-based loss (`chain._residual` over jet_mlp) equals the chain's loss and its
    gradient w.r.t. the cells (through the jet oracle's adjoint) equals the chain's dloss.
"""
import math

import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import (cell_offsets, derivative_chain, grid_sample_2d, grid_sample_3d,
                                        make_head)
from util import safe_coords


@pytest.mark.parametrize("dim,step,kcode,residual", [(2, "cosine", so.K_COSINE, "helmholtz"),
                                                     (2, "smoothstep", so.K_SMOOTHSTEP, "t2d"),
                                                     (3, "cosine", so.K_COSINE, "laplace")])
def test_jet_mlp_reproduces_the_nested_autograd_chain(dim, step, kcode, residual):
    from cosinesampler_b200.chain import _residual
    from cosinesampler_b200.jet import jet_mlp
    gen = torch.Generator().manual_seed(3 + dim)
    N, C, P = 3, 4, 60
    shape = (N, C, 7, 8) if dim == 2 else (N, C, 6, 6, 6)
    sizes = [shape[-1 - a] for a in range(dim)]
    cells = torch.rand(shape, generator=gen, dtype=torch.float64)
    pts = safe_coords(P, dim, sizes, N, True, gen).float().double()
    head = make_head(C, seed=2, dtype=torch.float64)
    fn = grid_sample_2d if dim == 2 else grid_sample_3d

    cells_r = cells.clone().requires_grad_(True)
    coords = [pts[:, a:a + 1].clone().requires_grad_(True) for a in range(dim)]
    ref = derivative_chain(lambda c, g: fn(c, g, step=step, offset=True), cells_r, coords, head, residual=residual)

    off = cell_offsets(N, True, dtype=torch.float32)
    kw = dict(pad=0, align=True, kernel=kcode, multicell=True, index_mode=2)
    jets = so.jet_forward(cells, pts, off, order=2, **kw).requires_grad_(True)
    u, first, second = jet_mlp(head, jets, dim)
    names = "xyz"[:dim]
    close = lambda a, b: torch.testing.assert_close(a, b, rtol=1e-9, atol=1e-11)
    close(u, ref["u"])
    for a in range(dim):
        close(first[a], ref["u_" + names[a]])
        close(second[a], ref["u_%s%s" % (names[a], names[a])])
    loss = torch.mean(_residual(u, first, second, residual, math.pi ** 2) ** 2)
    close(loss, ref["loss"])
    gJets, = torch.autograd.grad(loss, jets)
    dloss = so.jet_backward(gJets, cells.shape, pts, off, order=2, **kw)
    close(dloss, ref["dloss"])


def test_residual_coefficients():
    from cosinesampler_b200.jet import residual_coefficients
    r = residual_coefficients("helmholtz", 2, k2=9.0)
    assert (r.c_u, r.c_u3, list(r.c1), list(r.c2)) == (9.0, 0.0, [0.0, 0.0, 0.0], [1.0, 1.0, 0.0])
    r = residual_coefficients("laplace", 3)
    assert (r.c_u, list(r.c2)) == (1.0, [1.0, 1.0, 1.0])
    r = residual_coefficients("t2d", 2)                       # test_2d.py:221
    assert r.c_u == -5.0 and r.c_u3 == 5.0 and list(r.c1) == [0.0, 2.0, 0.0]
    assert abs(r.c2[0] + 1e-4) < 1e-10 and r.c2[1] == 0.0
    r = residual_coefficients({"c_u": 2.0, "c1": [1.0, -1.0], "c2": [0.5, 0.25]}, 2)
    assert (r.c_u, list(r.c1)[:2], list(r.c2)[:2]) == (2.0, [1.0, -1.0], [0.5, 0.25])
    with pytest.raises(ValueError):
        residual_coefficients("t2d", 3)
    with pytest.raises(ValueError):
        residual_coefficients("burgers", 2)


def test_fused_path_rejects_cpu_tensors_and_odd_channel_counts():
    from cosinesampler_b200 import jet
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        jet.SamplerJet2d.apply(torch.rand(2, 8, 8, 8), torch.rand(16, 2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        jet.fused_pde_step(torch.rand(2, 8, 8, 8), torch.rand(16, 2), make_head(8))
    assert jet.jet_bytes(2, 4, 16, 1 << 20, 256 * 256, 2) == 4 * ((1 << 20) * (2 + 5 * 16) + 4 * 16 * 65536)
