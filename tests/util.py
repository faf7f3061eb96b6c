"""Shared helpers for the parity tests."""
import numpy as np
import torch

RTOL = 1e-5          # north star: "within rtol 1e-5 in fp32 for every derivative order"
ATOL_SCALE = 1e-5    # plus 1e-5 of the tensor's own scale (SURVEY 7.1 tolerance recipe)


def assert_close_scaled(actual, expected, what="", rtol=RTOL, atol_scale=ATOL_SCALE, max_outlier_frac=0.0):
    """|a-b| <= rtol*|b| + atol_scale*max|b| elementwise; a tiny fraction of outliers may be
    allowed when the two sides round the index map differently (cell flips)."""
    a = actual.detach().double().cpu().reshape(-1)
    b = expected.detach().double().cpu().reshape(-1)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(actual.shape), tuple(expected.shape))
    if a.numel() == 0:
        return
    assert torch.isfinite(a).all(), "%s: non-finite values in actual" % what
    scale = b.abs().max().item()
    tol = rtol * b.abs() + atol_scale * scale
    bad = (a - b).abs() > tol
    nbad = int(bad.sum())
    if nbad > max_outlier_frac * a.numel():
        worst = ((a - b).abs() / (tol + 1e-300)).argmax().item()
        raise AssertionError(
            "%s: %d / %d elements outside rtol=%g + %g*max|ref| (scale %.3e); worst at %d: %r vs %r"
            % (what, nbad, a.numel(), rtol, atol_scale, scale, worst, a[worst].item(), b[worst].item()))


def safe_coords(P, dim, sizes, n_cells, multicell, gen, margin=0.02, lo=-1.0, hi=1.0):
    """Random normalised coordinates in (lo, hi) whose fractional cell position stays at
    least `margin` away from a cell edge for every cell offset, so that fp32 and fp64
    evaluations of the index map pick the same cell."""
    offs = np.linspace(0, 1 - 1 / n_cells, n_cells) if multicell else np.zeros(1)
    out = np.empty((P, dim), dtype=np.float64)
    for a in range(dim):
        s = sizes[a] - 1 - (1 if multicell else 0)
        got = 0
        while got < P:
            g = torch.rand(4 * P, generator=gen, dtype=torch.float64).numpy() * (hi - lo) + lo
            i = (g[:, None] + 1) / 2 * s + offs[None, :]
            fr = i - np.floor(i)
            ok = ((fr > margin) & (fr < 1 - margin)).all(1)
            g = g[ok][:P - got]
            out[got:got + len(g), a] = g
            got += len(g)
    return torch.from_numpy(out)
