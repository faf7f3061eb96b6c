"""Caller-side glue of a PIXEL-style training step, as the reference's comparison
scripts drive the operator (`test/test_2d.py:36-127,221-230`,
`test/test_3d.py:31-156,270-280`): replicate the coordinates over the N cells,
sample, sum over cells, a small MLP head, then nested `autograd.grad` calls for
u_a and u_aa, a PDE residual and its gradient w.r.t. the cells.

Generic over the sampler callable, so bench.py can time the same step through
this package's CUDA op and (for the CPU baseline leg) through the oracle.
This is bench / test scaffolding around the operator, not part of the operator.
"""
import math

import torch


def make_head(c_in, hidden=16, seed=0, device="cpu", dtype=torch.float32):
    """Linear(C,16)-Tanh-Linear(16,1) (`test_2d.py:42-47`, widened to C inputs)."""
    gen = torch.Generator().manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(c_in, hidden), torch.nn.Tanh(),
                              torch.nn.Linear(hidden, 1))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.empty_like(p).uniform_(-0.5, 0.5, generator=gen))
    return net.to(device=device, dtype=dtype)


def _grad(y, x):
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), retain_graph=True,
                               create_graph=True)[0]


def replicate_grid(coords, n_cells):
    """[P,1] coordinate columns -> grid [N,1,(1,)P,dim] (`test_2d.py:36-38`, `test_3d.py:31-32`)."""
    nd = len(coords)
    P = coords[0].shape[0]
    g = torch.cat(coords, -1)
    return g.reshape((1,) * nd + (P, nd)).repeat((n_cells,) + (1,) * (nd + 1))


def pde_loss(sampler, cells, coords, head, residual="helmholtz", k2=math.pi ** 2):
    """Residual loss of one batch of collocation points.

    residual: 'helmholtz'  f = sum_a u_aa + k2 * u        (README "Helmholtz equation")
              'laplace'    f = sum_a u_aa + u             (`test_3d.py:270`)
              't2d'        f = 2 u_y + 5 u^3 - 5 u - 1e-4 u_xx   (`test_2d.py:221`)
    """
    N, C = cells.shape[:2]
    grid = replicate_grid(coords, N)
    val = sampler(cells, grid)
    u = head(val.sum(0).reshape(C, -1).t())
    first = [None] * len(coords)
    second = [None] * len(coords)
    if residual == "t2d":
        first[1] = _grad(u, coords[1])
        first[0] = _grad(u, coords[0])
        second[0] = _grad(first[0], coords[0])
        f = first[1] * 2 + 5 * (u ** 3) - 5 * u - 0.0001 * second[0]
    else:
        for a in range(len(coords)):
            first[a] = _grad(u, coords[a])
            second[a] = _grad(first[a], coords[a])
        f = (k2 if residual == "helmholtz" else 1.0) * u
        for a in range(len(coords)):
            f = f + second[a]
    return torch.mean(f ** 2)


def _residual(u, first, second, residual, k2):
    if residual == "t2d":
        return first[1] * 2 + 5 * (u ** 3) - 5 * u - 0.0001 * second[0]
    f = (k2 if residual == "helmholtz" else 1.0) * u
    for s2 in second:
        f = f + s2
    return f


def jet_pde_loss(jet_sampler, cells, coords, head, residual="helmholtz", k2=math.pi ** 2):
    """The same loss as `pde_loss`, through the fused jet operator (`jet.SamplerJet2d/3d`):
    one gather pass yields z, z_a, z_aa summed over the cells, the head's chain rule is applied
    to the jets (`jet.jet_mlp`), and a single first-order backward scatters into the cells.
    `jet_sampler(cells, coords[P,dim]) -> jets [1+2*dim, C, P]`; `coords` is the list of [P,1]
    columns `pde_loss` takes."""
    from .jet import jet_mlp
    dim = len(coords)
    jets = jet_sampler(cells, torch.cat([c.detach() for c in coords], -1))
    u, first, second = jet_mlp(head, jets, dim, order=2)
    return torch.mean(_residual(u, first, second, residual, k2) ** 2)


def training_step(sampler, cells, coords, head, residual="helmholtz", k2=math.pi ** 2,
                  chunk=None, loss_scale=1.0, jet=False):
    """One fwd -> triple-bwd step (jet=True: `sampler` is a jet sampler, see `jet_pde_loss`): accumulates d loss / d cells (and head grads) into
    `.grad`.  `coords` is a list of [P,1] tensors (no grad needed); points are processed
    in chunks of `chunk` so that the [N,C,P] streams stay bounded.  Returns the loss
    (a 0-dim tensor, mean over all points, scaled by loss_scale)."""
    P = coords[0].shape[0]
    chunk = P if not chunk else min(chunk, P)
    # gradients are wanted for the cells and the head only, not for the coordinates:
    # naming them lets the engine (and the operator) skip the coordinate-gradient paths
    params = [cells] + [p for p in head.parameters() if p.requires_grad]
    total = None
    for s in range(0, P, chunk):
        e = min(P, s + chunk)
        if jet:
            loss = jet_pde_loss(sampler, cells, [c[s:e] for c in coords], head, residual, k2)
        else:
            cs = [c[s:e].detach().requires_grad_(True) for c in coords]
            loss = pde_loss(sampler, cells, cs, head, residual, k2)
        loss = loss * (loss_scale * (e - s) / P)
        loss.backward(inputs=params)
        total = loss.detach() if total is None else total + loss.detach()
    return total
