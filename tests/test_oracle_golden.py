"""The oracle restatement (oracle/grid_sampler_oracle.py) against golden vectors minted
from the REAL reference sampler (`/root/reference/test/grid_sampler.py`, via
oracle/make_golden.py): values and every derivative of the test_2d / test_3d chains,
fp64 and fp32, cosine / smoothstep / linear, multicell on and off.  Bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, grid_sample_3d, make_head

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("2d", 2, "t2d", ["cosine", "smoothstep", "bilinear"]),
         ("3d", 3, "laplace", ["cosine", "smoothstep", "trilinear"])]


@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
@pytest.mark.parametrize("name,nd,residual,steps", CASES)
def test_restatement_matches_reference_bit_for_bit(name, nd, residual, steps, tag, dtype):
    z = np.load(os.path.join(GOLDEN, "ref_sampler_%s_%s.npz" % (name, tag)))
    fn = grid_sample_2d if nd == 2 else grid_sample_3d
    checked = 0
    for step in steps:
        for offset in (1, 0):
            cells = torch.tensor(z["cells"]).requires_grad_(True)
            coords = [torch.tensor(z["coords"][:, a:a + 1]).requires_grad_(True) for a in range(nd)]
            head = make_head(cells.shape[1], seed=7, dtype=dtype)
            q = derivative_chain(lambda c, g: fn(c, g, step=step, offset=bool(offset)),
                                 cells, coords, head, residual=residual)
            for k, v in q.items():
                ref = z["%s|%d|%s" % (step, offset, k)]
                assert np.array_equal(v.detach().numpy(), ref), (name, tag, step, offset, k)
                checked += 1
    assert checked == len(steps) * 2 * (13 if nd == 2 else 17)


def test_linear_no_offset_equals_torch_grid_sample():
    """README.md:26-27 of the reference: linear kernel without multicell is F.grid_sample."""
    g = torch.Generator().manual_seed(3)
    inp = torch.rand(2, 3, 7, 9, generator=g, dtype=torch.float64)
    grid = torch.rand(2, 1, 50, 2, generator=g, dtype=torch.float64) * 2 - 1
    ours = grid_sample_2d(inp, grid, step="bilinear", offset=False)
    ref = torch.nn.functional.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.allclose(ours, ref, rtol=1e-12, atol=1e-12)
    inp3 = torch.rand(2, 3, 5, 5, 5, generator=g, dtype=torch.float64)
    grid3 = torch.rand(2, 1, 1, 40, 3, generator=g, dtype=torch.float64) * 2 - 1
    ours3 = grid_sample_3d(inp3, grid3, step="trilinear", offset=False)
    ref3 = torch.nn.functional.grid_sample(inp3, grid3, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.allclose(ours3.reshape(ref3.shape), ref3, rtol=1e-12, atol=1e-12)
