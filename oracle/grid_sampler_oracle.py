"""ORACLE (test infrastructure, not product code).

Device-agnostic restatement of the reference's pure-PyTorch sampler
(`/root/reference/test/grid_sampler.py`): `grid_sample_2d` (:3-89) and
`grid_sample_3d` (:91-239).  All derivative orders come from PyTorch autograd
over these functions, exactly as the reference's own comparison scripts
(`test/test_2d.py:130-206`, `test/test_3d.py:159-253`) obtain them.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this module.  The product path
(`cosinesampler_b200/`) never does.

Parity pin: `oracle/make_golden.py` imports the *real* reference module in the
build container and stores its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against them bit for bit
(fp64 and fp32).

What is restated (reference line numbers in brackets):
  * index map  i = ((g + 1) / 2) * (S - 1 - multicell) + offset[n], with
    offset = linspace(0, 1 - 1/N, N)                       [:33-41, :120-128]
  * corner indices floor(i), floor(i)+1 taken without grad  [:46-50, :132-164]
  * weights  k(right - i) for the low corner, 1 - k for the high corner, with
    k = identity | x^2(3-2x) | 0.5(1-cos(pi x))             [:18-25, :52-63]
  * index clamp to the valid range instead of zero padding  [:66-72, :180-211]
  * gather + blend, summed low-x first                      [:75-87, :214-238]

Differences on purpose:
  * the two hard-coded `.to("cuda")` calls (:34, :121) become `input.device`;
  * 3D: the reference scales/clamps/strides with a mix of IT/IH/IW that is only
    self-consistent for cubic grids (:123-125, :181-183, :217).  This
    restatement uses the CUDA kernels' axis convention (grid[...,0] -> W,
    grid[...,1] -> H, grid[...,2] -> D; `cosine_sampler_3d_kernel.cu:295-301`),
    which is identical to the reference for cubic grids (the only shapes the
    reference and BASELINE.json use).
"""
import math

import torch

_STEP_ALIASES = {
    "cosine": "cosine",
    "smoothstep": "smoothstep", "smooth-step": "smoothstep",
    "bilinear": "linear", "trilinear": "linear", "linear": "linear",
}


def _step_fn(step):
    kind = _STEP_ALIASES.get(step)
    if kind is None:
        raise NotImplementedError(step)
    if kind == "linear":
        return lambda r: r
    if kind == "smoothstep":
        return lambda r: (r ** 2) * (3 - 2 * r)
    return lambda r: 0.5 * (1 - torch.cos(torch.pi * r))


def cell_offsets(n_cells, multicell, device=None, dtype=torch.float32):
    """offset[n] of `modules_2d.py:24-27` / `grid_sampler.py:34`.

    linspace is evaluated in fp32 first (as the reference does) and only then
    cast, so an fp64 run sees the same offsets as the fp32 run."""
    if multicell:
        off = torch.linspace(0, 1 - (1 / n_cells), n_cells)
    else:
        off = torch.zeros(n_cells)
    return off.to(device=device, dtype=dtype)


def _sample_nd(input, coords, sizes, step, offset, corner_axis_order, shared_high_weight):
    """coords[a]: [N, P] normalised coordinate along the axis with extent
    sizes[a] and flat stride strides[a]; corner_axis_order lists the axes from
    the fastest-varying corner bit to the slowest (sum order of the blend).
    shared_high_weight: the 2D reference forms `1 - k` once per axis (:54-56),
    the 3D reference re-forms it inside every corner product (:170-177); the
    autograd accumulation order, hence the last bit, follows from that."""
    N, C = input.shape[:2]
    nd = len(coords)
    k = _step_fn(step)
    off = cell_offsets(N, bool(offset), input.device, coords[0].dtype).reshape(N, 1)
    shrink = 2 if offset else 1
    strides = [1] * nd
    for a in range(1, nd):
        strides[a] = strides[a - 1] * sizes[a - 1]

    lo_w, hi_w, lo_i, hi_i = [], [], [], []
    for a in range(nd):
        i = ((coords[a] + 1) / 2) * (sizes[a] - shrink) + off
        with torch.no_grad():
            left = torch.floor(i)
            right = left + 1
        lo_w.append(k(right - i))
        hi_w.append((1 - lo_w[a]) if shared_high_weight else None)
        with torch.no_grad():
            lo_i.append(torch.clamp(left, 0, sizes[a] - 1))
            hi_i.append(torch.clamp(right, 0, sizes[a] - 1))

    flat = input.reshape(N, C, -1)
    P = coords[0].shape[1]
    out = None
    for corner in range(1 << nd):
        w = None
        idx = 0
        # weight product in axis order 0..nd-1, as the reference writes it
        bits = {}
        for pos, a in enumerate(corner_axis_order):
            bits[a] = (corner >> pos) & 1
        for a in range(nd):
            if bits[a]:
                wa = hi_w[a] if shared_high_weight else (1 - lo_w[a])
            else:
                wa = lo_w[a]
            w = wa if w is None else w * wa
            idx = idx + (hi_i[a] if bits[a] else lo_i[a]) * strides[a]
        val = torch.gather(flat, 2, idx.long().view(N, 1, P).expand(N, C, P))
        term = val * w.view(N, 1, P)
        out = term if out is None else out + term
    return out


def grid_sample_2d(input, grid, step="cosine", offset=True):
    """input [N,C,IH,IW], grid [N,H,W,2] -> [N,C,H,W]  (`grid_sampler.py:3-89`).
    grid[...,0] indexes IW, grid[...,1] indexes IH; blend order nw, ne, sw, se."""
    N, C, IH, IW = input.shape
    _, H, W, _ = grid.shape
    gx = grid[..., 0].reshape(N, H * W)
    gy = grid[..., 1].reshape(N, H * W)
    out = _sample_nd(input, [gx, gy], [IW, IH], step, offset, corner_axis_order=[0, 1],
                     shared_high_weight=True)
    return out.view(N, C, H, W)


def grid_sample_3d(input, grid, step="cosine", offset=True):
    """input [N,C,ID,IH,IW], grid [N,1,1,P,3] or [N,1,P,3] -> [N,C,1,P]
    (`grid_sampler.py:91-239`).  Blend order: the H-axis bit varies fastest,
    then the D-axis bit, then the W-axis bit (nw/ne/sw/se front, then back)."""
    N, C, ID, IH, IW = input.shape
    P = grid.shape[-2]
    g = grid.reshape(N, P, 3)
    out = _sample_nd(input, [g[..., 0], g[..., 1], g[..., 2]], [IW, IH, ID], step, offset,
                     corner_axis_order=[1, 2, 0], shared_high_weight=False)
    return out.view(N, C, 1, P)


# --------------------------------------------------------------------------
# The derivative chains of test/test_2d.py:42-127 and test/test_3d.py:34-156,
# expressed once for any sampler callable.  Used for parity tests (our op vs
# this oracle) and for the CPU baseline timing in bench.py.
# --------------------------------------------------------------------------

def _g(y, x, grad_outputs=None, create_graph=True):
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, x, grad_outputs=grad_outputs, retain_graph=True,
                               create_graph=create_graph)[0]


def make_head(c_in, hidden=16, seed=0, device="cpu", dtype=torch.float32):
    """Linear(C,16)-Tanh-Linear(16,1) head of test_2d.py:42-47."""
    gen = torch.Generator().manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(c_in, hidden), torch.nn.Tanh(),
                              torch.nn.Linear(hidden, 1))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.empty_like(p).uniform_(-0.5, 0.5, generator=gen))
    return net.to(device=device, dtype=dtype)


def derivative_chain(sampler, cells, coords, head, residual="t2d", orders=3):
    """Rebuild the quantity list of test_2d.py:55-127,221-230 (2 coords) or
    test_3d.py:52-156,270-280 (3 coords).

    sampler(cells, grid) -> [N,C,...,P];  coords: list of [P,1] leaf tensors with
    requires_grad, in *grid channel order* (coords[0] -> grid[...,0] -> W axis).
    Returns an ordered dict name -> tensor."""
    nd = len(coords)
    N, C = cells.shape[:2]
    P = coords[0].shape[0]
    grid = torch.cat(coords, -1)
    grid = grid.reshape((1,) * nd + (P, nd)).repeat((N,) + (1,) * (nd + 1))
    val = sampler(cells, grid)
    u = head(val.sum(0).reshape(C, -1).t())
    names = "xyz"[:nd]
    q = {"val": val, "u": u}
    q["u_cell"] = _g(u, cells)
    if orders >= 1:
        for a in range(nd):
            q["u_" + names[a]] = _g(u, coords[a])
    if orders >= 2:
        for a in range(nd):
            q["u_%s%s" % (names[a], names[a])] = _g(q["u_" + names[a]], coords[a])
        for a in range(nd):
            q["u_%s_cell" % names[a]] = _g(q["u_" + names[a]], cells)
    if orders >= 3:
        for a in range(nd):
            nm = "u_%s%s" % (names[a], names[a])
            q[nm + "_cell"] = _g(q[nm], cells)
        if residual == "t2d":      # test_2d.py:221
            f = q["u_y"] * 2 + 5 * (u ** 3) - 5 * u - 0.0001 * q["u_xx"]
        elif residual == "helmholtz":   # README "Second-order PDE (Helmholtz equation)"
            f = q["u_xx"] + q["u_yy"] + (math.pi ** 2) * u
        elif residual == "laplace":     # test_3d.py:270
            f = u
            for a in range(nd):
                f = f + q["u_%s%s" % (names[a], names[a])]
        else:
            raise ValueError(residual)
        loss = torch.mean(f ** 2)
        q["loss"] = loss
        q["dloss"] = _g(loss, cells)
    return q
