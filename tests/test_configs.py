"""BASELINE.json `configs`, one test each at the named shapes (configs 3-5 at full size are covered
by test_gpu_properties.py and bench.py; config 2 by test_gpu_grid_sample.py).

config 1: 2D cosine, multicell=False, grid [1,16,64,64] fp32, 4096 points: forward + backward
parity with the PyTorch reference (test/grid_sampler.py)."""
import os
import sys
import types

import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, make_head
from util import assert_close_scaled


def _config1(dtype=torch.float32):
    gen = torch.Generator().manual_seed(1)
    cells = torch.rand(1, 16, 64, 64, generator=gen).to(dtype)
    coords = (torch.rand(4096, 2, generator=gen) * 2 - 1).to(dtype)
    return cells, coords


def _import_real_reference():
    if not os.path.isdir("/root/reference/test"):
        return None
    stub_pkg = types.ModuleType("cosine_sampler_2d")
    stub_mod = types.ModuleType("cosine_sampler_2d.modules_2d")
    stub_mod.CosineSampler2d = None
    stub_pkg.modules_2d = stub_mod
    saved = {k: sys.modules.get(k) for k in ("cosine_sampler_2d", "cosine_sampler_2d.modules_2d", "grid_sampler")}
    sys.modules["cosine_sampler_2d"] = stub_pkg
    sys.modules["cosine_sampler_2d.modules_2d"] = stub_mod
    sys.modules.pop("grid_sampler", None)
    sys.path.insert(0, "/root/reference/test")
    try:
        import grid_sampler as ref
        return ref
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_config1_cpu_oracle_forward_backward():
    """the oracle restatement == the real reference (when it is present in this container), and the
    stage oracle's F / B == autograd over it, at config 1's shape (offset=False runs on the CPU)"""
    cells, coords = _config1()
    grid = coords.reshape(1, 1, 4096, 2).clone()
    c = cells.clone().requires_grad_(True)
    g = grid.clone().requires_grad_(True)
    out = grid_sample_2d(c, g, step="cosine", offset=False)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    gI, gG = torch.autograd.grad((out * w).sum(), [c, g])
    ref = _import_real_reference()
    if ref is not None:
        c2 = cells.clone().requires_grad_(True)
        g2 = grid.clone().requires_grad_(True)
        out2 = ref.grid_sample_2d(c2, g2, step="cosine", offset=False)
        rI, rG = torch.autograd.grad((out2 * w).sum(), [c2, g2])
        assert torch.equal(out, out2) and torch.equal(gI, rI) and torch.equal(gG, rG)
    off = torch.zeros(1)
    f = so.forward(cells, grid, off, kernel=so.K_COSINE, multicell=False)
    assert_close_scaled(f.reshape(out.shape), out, "stage oracle F vs reference sampler", rtol=1e-5, atol_scale=2e-6)
    sI, sG = so.backward(w, cells, grid, off, kernel=so.K_COSINE, multicell=False)
    assert_close_scaled(sI, gI, "stage oracle gInput", rtol=1e-5, atol_scale=2e-6)
    assert_close_scaled(sG, gG, "stage oracle gGrid", rtol=1e-5, atol_scale=2e-6)


@pytest.mark.gpu
def test_config1_gpu_forward_backward_and_chain(cuda):
    from cosine_sampler_2d import CosineSampler2d
    cells, coords = _config1()
    grid = coords.reshape(1, 1, 4096, 2).clone()
    c = cells.to(cuda).requires_grad_(True)
    g = grid.to(cuda).requires_grad_(True)
    out = CosineSampler2d.apply(c, g, "zeros", True, "cosine", False)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
    gI, gG = torch.autograd.grad((out * w.to(cuda)).sum(), [c, g])
    c2 = cells.clone().double().requires_grad_(True)
    g2 = grid.clone().double().requires_grad_(True)
    ref = grid_sample_2d(c2, g2, step="cosine", offset=False)
    rI, rG = torch.autograd.grad((ref * w.double()).sum(), [c2, g2])
    assert_close_scaled(out, ref, "config 1 forward", max_outlier_frac=5e-4)
    assert_close_scaled(gI, rI, "config 1 gInput", max_outlier_frac=5e-4)
    assert_close_scaled(gG, rG, "config 1 gGrid", max_outlier_frac=5e-4)
    # and every derivative order through the test_2d.py chain on the same shapes
    head = make_head(16, seed=4).to(cuda)
    head64 = make_head(16, seed=4, dtype=torch.float64)

    def run(sampler, dev, dt, hd):
        cc = cells.to(device=dev, dtype=dt).requires_grad_(True)
        xs = [coords[:, a:a + 1].to(device=dev, dtype=dt).requires_grad_(True) for a in range(2)]
        return derivative_chain(sampler, cc, xs, hd, residual="helmholtz")
    ours = run(lambda a, b: CosineSampler2d.apply(a, b, "zeros", True, "cosine", False), cuda, torch.float32, head)
    ref = run(lambda a, b: grid_sample_2d(a, b, step="cosine", offset=False), "cpu", torch.float64, head64)
    for k in ours:
        # fp32 vs fp64 index maps may pick different cells for a handful of the 4096 points
        assert_close_scaled(ours[k].reshape(ref[k].shape), ref[k], "config 1 chain " + k, rtol=1e-4,
                            atol_scale=2e-5, max_outlier_frac=2e-3)
