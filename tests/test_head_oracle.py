"""Pin oracle/head_oracle.py (closed-form level-0..3 passes of the Linear-Tanh-Linear head) to PyTorch
autograd over the plain nn.Sequential head: the whole quantity list of the reference's scripts
(u, u_cell, u_a, u_aa, u_a_cell, u_aa_cell, loss, dloss; test_2d.py:55-127,221-230) and the head's
parameter gradients of the loss must agree when the head is swapped for the closed-form chain."""
import pytest
import torch

from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, grid_sample_3d, make_head
from oracle.head_oracle import ClosedFormHead, level1, level2, level3
from util import safe_coords


@pytest.mark.parametrize("dim,residual", [(2, "helmholtz"), (2, "t2d"), (3, "laplace")])
def test_closed_form_head_reproduces_the_nested_autograd_chain(dim, residual):
    gen = torch.Generator().manual_seed(9 + dim)
    N, C, P = 3, 4, 50
    shape = (N, C, 7, 8) if dim == 2 else (N, C, 6, 6, 6)
    sizes = [shape[-1 - a] for a in range(dim)]
    cells0 = torch.rand(shape, generator=gen, dtype=torch.float64)
    pts = safe_coords(P, dim, sizes, N, True, gen).double()
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    sampler = lambda c, g: fn(c, g, step="cosine", offset=True)
    out = {}
    for kind in ("sequential", "closed_form"):
        seq = make_head(C, seed=5, dtype=torch.float64)
        head = seq if kind == "sequential" else ClosedFormHead(seq)
        cells = cells0.clone().requires_grad_(True)
        coords = [pts[:, a:a + 1].clone().requires_grad_(True) for a in range(dim)]
        q = derivative_chain(sampler, cells, coords, head, residual=residual)
        pg = torch.autograd.grad(q["loss"], list(seq.parameters()), retain_graph=True)
        out[kind] = ({k: v.detach() for k, v in q.items()}, pg)
    for k, v in out["sequential"][0].items():
        torch.testing.assert_close(out["closed_form"][0][k], v, rtol=1e-9, atol=1e-11, msg=lambda m, k=k: "%s: %s" % (k, m))
    for a, b in zip(out["closed_form"][1], out["sequential"][1]):
        torch.testing.assert_close(a, b, rtol=1e-9, atol=1e-11)


def test_each_level_is_the_gradient_of_the_previous_one():
    gen = torch.Generator().manual_seed(1)
    P, C, K = 23, 5, 16
    f = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    z, gu, ggz, g_gz2, g_ggu = f(P, C), f(P), f(P, C), f(P, C), f(P)
    W1, b1, w2 = f(K, C) * 0.4, f(K) * 0.4, f(1, K) * 0.4
    leaves = [t.requires_grad_(True) for t in (z, gu, ggz, W1, b1, w2)]
    z, gu, ggz, W1, b1, w2 = leaves
    # level 2 = gradient of <ggz, level1.gz>
    gz = level1(z, gu, W1, b1, w2)[0]
    ref = torch.autograd.grad((ggz.detach() * gz).sum(), [z, gu, W1, b1, w2], create_graph=True)
    got = level2(z, gu, ggz, W1, b1, w2)
    for a, b in zip(got, ref):
        torch.testing.assert_close(a.reshape(b.shape), b, rtol=1e-10, atol=1e-12)
    # level 3 = gradient of <g_gz2, gz2> + <g_ggu, ggu>
    T = (g_gz2 * got[0]).sum() + (g_ggu * got[1]).sum()
    ref3 = torch.autograd.grad(T, [z, gu, ggz, W1, b1, w2])
    got3 = level3(z, gu, ggz, g_gz2, g_ggu, W1, b1, w2)
    for a, b in zip(got3, ref3):
        torch.testing.assert_close(a.reshape(b.shape), b, rtol=1e-10, atol=1e-12)
