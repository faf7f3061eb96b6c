"""Pin the algebra of the one-pass fused step on the CPU: `oracle/onepass_oracle.py` restates, in float64 torch ops,
what `cs_head_premix` / `cs_pde_fused_kernel` / `cs_head_postmix` compute (W1-mixed cells, jets of the hidden
pre-activations, the closed-form head / residual / gradient stage, the adjoint scatter, the back-mix) and must
reproduce the reference's own formulation -- nested `autograd.grad` over its pure-PyTorch sampler and the
Linear-Tanh-Linear head (`test/test_2d.py:36-127`) -- for the loss, d loss / d cells and every head gradient.
The GPU tests (`tests/test_gpu_onepass.py`) then check the CUDA kernels against the same chain."""
import math

import pytest
import torch

from oracle import onepass_oracle, stage_oracle as so
from oracle.grid_sampler_oracle import cell_offsets, derivative_chain, grid_sample_2d, grid_sample_3d, make_head
from util import safe_coords

RESIDUALS = {
    "helmholtz": lambda dim: dict(c_u=math.pi ** 2, c_u3=0.0, c1=[0.0] * dim, c2=[1.0] * dim),
    "laplace": lambda dim: dict(c_u=1.0, c_u3=0.0, c1=[0.0] * dim, c2=[1.0] * dim),
    "t2d": lambda dim: dict(c_u=-5.0, c_u3=5.0, c1=[0.0, 2.0], c2=[-0.0001, 0.0]),            # test_2d.py:221
}


@pytest.mark.parametrize("dim,step,kcode,residual,K", [(2, "cosine", so.K_COSINE, "helmholtz", 16),
                                                       (2, "smoothstep", so.K_SMOOTHSTEP, "t2d", 8),
                                                       (3, "cosine", so.K_COSINE, "laplace", 16),
                                                       (3, "smoothstep", so.K_SMOOTHSTEP, "laplace", 4)])
def test_one_pass_algebra_reproduces_the_nested_autograd_chain(dim, step, kcode, residual, K):
    gen = torch.Generator().manual_seed(11 * dim + K)
    N, C, P = 3, 5, 80
    shape = (N, C, 7, 9) if dim == 2 else (N, C, 5, 6, 7)
    sizes = [shape[-1 - a] for a in range(dim)]
    cells0 = torch.rand(shape, generator=gen, dtype=torch.float64)
    pts = safe_coords(P, dim, sizes, N, True, gen).double()
    head = make_head(C, hidden=K, seed=4, dtype=torch.float64)
    fn = grid_sample_2d if dim == 2 else grid_sample_3d

    cells = cells0.clone().requires_grad_(True)
    coords = [pts[:, a:a + 1].clone().requires_grad_(True) for a in range(dim)]
    ref = derivative_chain(lambda c, g: fn(c, g, step=step, offset=True), cells, coords, head, residual=residual)
    ref_pg = torch.autograd.grad(ref["loss"], list(head.parameters()), retain_graph=True)

    W1, b1, w2, b2 = [p.detach() for p in head.parameters()]
    off = cell_offsets(N, True, dtype=torch.float32)
    got = onepass_oracle.one_pass_step(cells0, pts, W1, b1, w2, b2, RESIDUALS[residual](dim), 1.0 / P, off,
                                       kernel=kcode)
    close = lambda a, b, what: torch.testing.assert_close(a.reshape(b.shape), b, rtol=1e-9, atol=1e-11,
                                                          msg=lambda m: "%s: %s" % (what, m))
    close(got["u"], ref["u"].detach(), "u")
    close(got["loss"], ref["loss"].detach(), "loss")
    close(got["gInput"], ref["dloss"].detach(), "d loss / d cells")
    for name, a, b in zip(("gW1", "gb1", "gw2", "gb2"), (got["gW1"], got["gb1"], got["gw2"], got["gb2"]), ref_pg):
        close(a, b, name)


def test_premix_commutes_with_the_sampler():
    """W1 . jets(V) == jets(W1 . V): the identity that removes the per-point matrix products."""
    gen = torch.Generator().manual_seed(3)
    N, C, K, P = 4, 6, 8, 50
    V = torch.rand(N, C, 6, 7, generator=gen, dtype=torch.float64)
    W1 = torch.randn(K, C, generator=gen, dtype=torch.float64)
    pts = safe_coords(P, 2, [7, 6], N, True, gen).double()
    off = cell_offsets(N, True, dtype=torch.float32)
    kw = dict(order=2, pad=0, align=True, kernel=so.K_COSINE, multicell=True, index_mode=2)
    a = torch.einsum("kc,jcp->jkp", W1, so.jet_forward(V, pts, off, **kw))
    b = so.jet_forward(torch.einsum("kc,nc...->nk...", W1, V), pts, off, **kw)
    torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-12)
