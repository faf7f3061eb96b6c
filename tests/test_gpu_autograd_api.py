"""Behaviour of the drop-in autograd Functions that the reference's users rely on: which inputs
get gradients, plain .backward(), repeated use of one graph, no_grad, argument errors."""
import pytest
import torch

from oracle.grid_sampler_oracle import grid_sample_2d, grid_sample_3d
from util import assert_close_scaled

pytestmark = pytest.mark.gpu


def _data(cuda, dim=2):
    gen = torch.Generator().manual_seed(9)
    shape = (3, 8, 10, 11) if dim == 2 else (2, 8, 5, 6, 7)
    cells = torch.rand(shape, generator=gen).to(cuda)
    gshape = (shape[0], 1, 200, 2) if dim == 2 else (shape[0], 1, 1, 200, 3)
    grid = (torch.rand(gshape, generator=gen) * 1.8 - 0.9).to(cuda)
    return cells, grid


def test_signature_defaults_match_the_reference(cuda):
    """apply(input, grid) == apply(input, grid, 'zeros', True, 'cosine', True)  (mod2d:22)"""
    from cosine_sampler_2d import CosineSampler2d
    cells, grid = _data(cuda)
    assert torch.equal(CosineSampler2d.apply(cells, grid),
                       CosineSampler2d.apply(cells, grid, "zeros", True, "cosine", True))
    with pytest.raises(TypeError):
        CosineSampler2d.apply(cells, grid, "zeros", True, "trilinear", True)   # 3D name in 2D (mod2d:12-18)


@pytest.mark.parametrize("req_cells,req_grid", [(True, False), (False, True), (True, True)])
def test_only_requested_gradients_are_produced(cuda, req_cells, req_grid):
    from cosine_sampler_2d import CosineSampler2d
    from cosinesampler_b200 import _lib
    cells, grid = _data(cuda)
    c = cells.clone().requires_grad_(req_cells)
    g = grid.clone().requires_grad_(req_grid)
    out = CosineSampler2d.apply(c, g, "zeros", True, "cosine", True)
    n0 = _lib.launch_count()
    out.square().sum().backward()
    launches = _lib.launch_count() - n0
    assert (c.grad is not None) == req_cells and (g.grad is not None) == req_grid
    # gInput costs a scatter + a layout conversion, gGrid rides in the same or its own launch
    assert launches == (2 if req_cells else 1)
    cr = cells.double().cpu().requires_grad_(True)
    gr = grid.double().cpu().requires_grad_(True)
    grid_sample_2d(cr, gr, step="cosine", offset=True).square().sum().backward()
    if req_cells:
        assert_close_scaled(c.grad, cr.grad, "cells.grad", max_outlier_frac=1e-3)
    if req_grid:
        assert_close_scaled(g.grad, gr.grad, "grid.grad", max_outlier_frac=1e-3)


def test_graph_can_be_differentiated_repeatedly_and_under_no_grad(cuda):
    from cosine_sampler_3d import CosineSampler3d
    cells, grid = _data(cuda, 3)
    c = cells.clone().requires_grad_(True)
    g = grid.clone().requires_grad_(True)
    out = CosineSampler3d.apply(c, g, "zeros", True, "smooth-step", True)
    a1 = torch.autograd.grad(out.sum(), g, retain_graph=True)[0]
    a2 = torch.autograd.grad(out.sum(), g, retain_graph=True)[0]
    assert torch.equal(a1, a2)
    with torch.no_grad():
        o2 = CosineSampler3d.apply(c, g, "zeros", True, "smooth-step", True)
    assert not o2.requires_grad and torch.equal(o2, out.detach())
    ref = grid_sample_3d(cells.cpu().double(), grid.cpu().double(), step="smoothstep", offset=True)
    assert_close_scaled(out, ref.reshape(out.shape), "3D forward", max_outlier_frac=1e-3)


def test_input_modified_in_place_between_calls_is_restaged(cuda):
    """the channel-last staging is tied to (storage, version) of `input` at forward time"""
    from cosine_sampler_2d import CosineSampler2d
    cells, grid = _data(cuda)
    p = torch.nn.Parameter(cells.clone())
    o1 = CosineSampler2d.apply(p, grid).detach().clone()
    with torch.no_grad():
        p.mul_(2.0)
    o2 = CosineSampler2d.apply(p, grid).detach()
    assert_close_scaled(o2, 2 * o1, "output after in-place update of the cells", rtol=1e-6, atol_scale=1e-7)
    p.data.mul_(0.5)          # bypasses the version counter: staging is redone at every forward anyway
    o3 = CosineSampler2d.apply(p, grid).detach()
    assert_close_scaled(o3, o1, "output after .data update", rtol=1e-6, atol_scale=1e-7)
