"""The whole PIXEL step at BASELINE config-3 / config-4 sizes against the reference's ORIGINAL CUDA op
(oracle/_ref/_cosine_{2,3}d.so, built from /root/reference by oracle/build_ref.py; skipped when not built),
driven with the reference's own call pattern (tools/refop_chain.py: every first backward scatters gInput, the
triple backward is a BBB call plus a second BB call, modules_2d.py:20-111): loss and d loss / d cells of
  * the drop-in operator (CosineSampler2d / 3d),
  * the round-1 fused path (jets -> tensor-core head -> scatter),
  * the one-pass fused step (binned points, W1-mixed cells),
with index_mode='fused' (the reference build's single-fma index map).  The cosine kernel of the reference
build uses MUFU sin/cos (--use_fast_math): rtol 1e-4 + 2e-5 of scale; smoothstep has no transcendental."""
import os
import sys

import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CONFIGS = {
    "cfg3": (2, (4, 16, 256, 256), 2 ** 20, 0, "cosine", "helmholtz"),
    "cfg4": (3, (4, 16, 64, 64, 64), 2 ** 22, 2, "smooth-step", "laplace"),
}


@pytest.mark.parametrize("name", ["cfg3", "cfg4"])
def test_step_matches_reference_cuda_op_at_baseline_size(cuda, name):
    from oracle import build_ref
    dim, shape, P, kcode, kname, residual = CONFIGS[name]
    ref = build_ref.load("_cosine_%dd" % dim)
    if ref is None:
        pytest.skip("oracle/_ref/_cosine_%dd.so not built" % dim)
    from refop_chain import make_refop_sampler
    from cosinesampler_b200 import chain, fused, jet, ops
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    torch.manual_seed(0)
    cells0 = torch.rand(shape, device=cuda)
    coords = (torch.rand(P, dim, device=cuda) * 2 - 1).contiguous()
    cols = [coords[:, a:a + 1] for a in range(dim)]
    S = CosineSampler2d if dim == 2 else CosineSampler3d
    chunk = 2 ** 20
    ops.set_index_mode("fused")
    try:
        res = {}
        for mode in ("reference_op", "dropin", "jets", "onepass"):
            cells = torch.nn.Parameter(cells0.clone())
            head = chain.make_head(shape[1], seed=0, device=cuda)
            if mode == "reference_op":
                loss = chain.training_step(make_refop_sampler(ref, kcode), cells, cols, head, residual, chunk=chunk)
            elif mode == "dropin":
                loss = chain.training_step(lambda c, g: S.apply(c, g, "zeros", True, kname, True), cells, cols, head,
                                           residual, chunk=chunk)
            elif mode == "jets":
                loss = jet.fused_pde_step(cells, coords, head, residual, kernel=kname, mode="jets", chunk=chunk)
            else:
                loss = fused.one_pass_pde_step(cells, coords, head, residual, kernel=kname)
            res[mode] = (loss.detach().clone(), cells.grad.clone(), [p.grad.clone() for p in head.parameters()])
            del cells, head, loss
            torch.cuda.empty_cache()
    finally:
        ops.set_index_mode("separate")
    rl, rg, rh = res["reference_op"]
    for mode in ("dropin", "jets", "onepass"):
        what = "%s %s vs reference CUDA op: " % (name, mode)
        assert_close_scaled(res[mode][0], rl, what + "loss", rtol=1e-4)
        assert_close_scaled(res[mode][1], rg, what + "d loss / d cells", rtol=1e-4, atol_scale=2e-5, max_outlier_frac=1e-4)
        # the head gradients are fp32 sums of 2^20 - 2^22 terms of both signs (gb2 is a single such sum): the two
        # sides order them differently (torch reductions vs fp32 atomics), which is worth a few 1e-4
        for a, b in zip(res[mode][2], rh):
            assert_close_scaled(a, b, what + "head grad", rtol=5e-4, atol_scale=1e-4)
