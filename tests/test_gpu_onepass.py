"""The one-pass fused step (`cosinesampler_b200.fused`, csrc/cs_fused.cuh): W1-mixed cells, binned
points, gather -> head -> scatter in one kernel with shared-memory aggregation windows.

  * cs_bin_points is a permutation that groups the points by texel;
  * cs_head_premix / cs_head_postmix equal the einsums they stand for;
  * the step equals the reference's nested-autograd chain through the oracle sampler in fp64 (loss,
    d loss / d cells, head gradients), the drop-in operator's `chain.training_step`, and the round-1
    fused path (`jet.FusedPdeStep`), for every kernel / padding / hidden width / aggregation mode;
  * binned == unbinned on duplicate and edge coordinates.
"""
import math

import pytest
import torch

from oracle.grid_sampler_oracle import grid_sample_2d, grid_sample_3d, make_head
from util import assert_close_scaled, safe_coords

pytestmark = pytest.mark.gpu


def _head(C, K, seed, device="cpu", dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(C, K), torch.nn.Tanh(), torch.nn.Linear(K, 1))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.empty_like(p).uniform_(-0.5, 0.5, generator=gen))
    return net.to(device=device, dtype=dtype)


def _oracle_step(cells0, coords0, head64, dim, kernel, residual, k2=math.pi ** 2):
    """loss = mean f^2 and its gradients through the pure-PyTorch sampler in fp64 (the reference's own
    definition of every derivative order: test_2d.py:130-240)."""
    from cosinesampler_b200.chain import training_step
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    step = kernel
    cells = torch.nn.Parameter(cells0.double())
    cols = [coords0[:, a:a + 1].double() for a in range(dim)]
    loss = training_step(lambda c, g: fn(c, g, step=step, offset=True), cells, cols, head64, residual, k2)
    return loss.detach(), cells.grad, [p.grad for p in head64.parameters()]


@pytest.mark.parametrize("dim", [2, 3])
def test_bin_points_is_a_texel_grouping_permutation(cuda, dim):
    from cosinesampler_b200 import fused
    from cosinesampler_b200.autograd import cell_offsets
    gen = torch.Generator().manual_seed(dim)
    shape = (4, 16, 40, 56) if dim == 2 else (4, 16, 12, 20, 24)
    cells = torch.rand(shape, generator=gen).to(cuda)
    for P in (1, 7, 5000, 100003):
        coords = (torch.rand(P, dim, generator=gen) * 2.2 - 1.1).to(cuda)       # some out of range
        coords[::17] = coords[0].clone()                                                # duplicates
        off = cell_offsets(4, True, cuda)
        binned, perm = fused.bin_points(cells, coords, off, want_perm=True)
        assert torch.equal(binned, coords[perm.long()])
        assert torch.equal(torch.sort(perm.long()).values, torch.arange(P, device=cuda))
        # points of one texel are contiguous: the texel id sequence of the binned points has no
        # texel appearing in two separate runs
        sizes = shape[2:][::-1]                                                 # W, H(, D)
        tex = torch.zeros(P, dtype=torch.long, device=cuda)
        mul = 1
        for a in range(dim):
            s = sizes[a] - 2
            i = ((binned[:, a] + 1) * 0.5) * s + off[0]
            l = torch.floor(i).clamp(0, sizes[a] - 1).long()
            tex += l * mul
            mul *= sizes[a]
        change = torch.ones(P, dtype=torch.bool, device=cuda)
        change[1:] = tex[1:] != tex[:-1]
        runs = tex[change]
        assert runs.numel() == torch.unique(runs).numel(), "a texel appears in two separate runs"


@pytest.mark.parametrize("K", [4, 8, 16, 32, 64])
def test_premix_postmix_match_einsum(cuda, K):
    from cosinesampler_b200 import fused
    gen = torch.Generator().manual_seed(K)
    for shape in ((3, 16, 17, 23), (2, 5, 6, 7, 9), (1, 64, 8, 8)):
        N, C = shape[:2]
        cells = torch.rand(shape, generator=gen).to(cuda)
        W1 = (torch.rand(K, C, generator=gen) - 0.5).to(cuda)
        T = cells[0, 0].numel()
        Vh = fused.head_premix(cells, W1)
        ref = torch.einsum("kc,nct->ntk", W1.double(), cells.double().reshape(N, C, T))
        assert float(Vh[N * T].abs().max()) == 0.0
        Vh = Vh[:N * T].view(N, T, K)
        assert_close_scaled(Vh, ref, "premix K=%d %s" % (K, shape))
        g = torch.randn(N, T, K, generator=gen).to(cuda)
        for hidden_first in (False, True):
            gW1 = torch.full((K, C), 0.5, device=cuda)
            gin = fused.head_postmix(g.transpose(1, 2).contiguous() if hidden_first else g, cells, W1, gW1,
                                     hidden_first)
            assert gin.shape == cells.shape
            assert_close_scaled(gin.reshape(N, C, T), torch.einsum("kc,ntk->nct", W1.double(), g.double()),
                                "postmix gInput K=%d %s hf=%s" % (K, shape, hidden_first))
            assert_close_scaled(gW1, 0.5 + torch.einsum("ntk,nct->kc", g.double(), cells.double().reshape(N, C, T)),
                                "postmix gW1 K=%d %s hf=%s" % (K, shape, hidden_first), rtol=2e-5, atol_scale=2e-5)


CASES = [
    # dim, cells shape, kernel, residual, K
    (2, (4, 16, 32, 32), "cosine", "helmholtz", 16),
    (2, (4, 16, 32, 32), "smooth-step", "t2d", 16),
    (2, (4, 8, 24, 40), "bilinear", "helmholtz", 32),
    (2, (3, 4, 20, 20), "cosine", "helmholtz", 8),
    (2, (2, 20, 16, 16), "cosine", "laplace", 4),
    (3, (4, 16, 12, 12, 12), "smooth-step", "laplace", 16),
    (3, (2, 8, 10, 10, 10), "cosine", "helmholtz", 32),
    (2, (4, 8, 24, 24), "cosine", "helmholtz", 64),
    (3, (3, 4, 9, 9, 9), "smooth-step", "laplace", 64),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dD-%s-%s-K%d-C%d" % (c[0], c[2], c[3], c[4], c[1][1]))
def test_one_pass_step_matches_fp64_oracle_chain(cuda, case):
    from cosinesampler_b200 import fused
    dim, shape, kernel, residual, K = case
    gen = torch.Generator().manual_seed(3 * dim + K)
    N, C = shape[:2]
    P = 6000
    cells0 = torch.rand(shape, generator=gen)
    sizes = shape[2:][::-1]
    coords0 = safe_coords(P, dim, sizes, N, True, gen).float()
    head64 = _head(C, K, seed=5, dtype=torch.float64)
    ref_loss, ref_gc, ref_gh = _oracle_step(cells0, coords0, head64, dim, kernel, residual)
    for bin_, agg, chunk in ((True, "auto", None), (False, "off", None), (True, "auto", 2500), (False, "auto", 1111)):
        cells = torch.nn.Parameter(cells0.clone().to(cuda))
        head = _head(C, K, seed=5, device=cuda)
        loss = fused.one_pass_pde_step(cells, coords0.to(cuda).contiguous(), head, residual, kernel=kernel,
                                       chunk=chunk, bin=bin_, aggregate=agg)
        what = "onepass %dD %s %s K=%d bin=%s agg=%s chunk=%s " % (dim, kernel, residual, K, bin_, agg, chunk)
        # the north star's tolerance: rtol 1e-5 (+ 1e-5 of the tensor's scale for sums); measured
        # (profiles/r2/onepass_errors.jsonl): loss 2e-7, d loss / d cells 3.5e-6 of scale, head gradients 2.2e-6
        assert_close_scaled(loss, ref_loss, what + "loss", rtol=1e-5)
        assert_close_scaled(cells.grad, ref_gc, what + "cells.grad", rtol=1e-5, atol_scale=1e-5)
        for a, b in zip([p.grad for p in head.parameters()], ref_gh):
            assert_close_scaled(a, b, what + "head grad", rtol=1e-5, atol_scale=1e-5)


@pytest.mark.parametrize("dim", [2, 3])
def test_one_pass_step_equals_dropin_and_round1_fused(cuda, dim):
    """Unrestricted coordinates (cell edges, out of range, duplicates) and every padding mode: the one-pass
    step against the drop-in operator's chain and the jets + tensor-core-head path of round 1."""
    from cosinesampler_b200 import fused, jet
    from cosinesampler_b200.chain import training_step
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    D = CosineSampler2d if dim == 2 else CosineSampler3d
    shape = (4, 16, 48, 48) if dim == 2 else (4, 16, 14, 14, 14)
    gen = torch.Generator().manual_seed(40 + dim)
    P = 20000
    cells0 = torch.rand(shape, generator=gen)
    coords = torch.rand(P, dim, generator=gen) * 2.1 - 1.05
    coords[:64] = 1.0                       # the last texel
    coords[64:128] = -1.0
    coords[128:4096] = coords[128]          # 4000 copies of one point: a long run in one texel
    coords = coords.to(cuda).contiguous()
    cols = [coords[:, a:a + 1].contiguous() for a in range(dim)]
    kernel = "cosine" if dim == 2 else "smooth-step"
    residual = "helmholtz" if dim == 2 else "laplace"
    for pad in ("zeros", "border", "reflection"):
        res = {}
        for mode in ("dropin", "jets", "onepass", "onepass_unbinned"):
            cells = torch.nn.Parameter(cells0.clone().to(cuda))
            head = make_head(shape[1], seed=3).to(cuda)
            if mode == "dropin":
                loss = training_step(lambda c, g: D.apply(c, g, pad, True, kernel, True), cells, cols, head, residual)
            elif mode == "jets":
                loss = jet.fused_pde_step(cells, coords, head, residual, padding_mode=pad, kernel=kernel, mode="jets")
            else:
                loss = fused.one_pass_pde_step(cells, coords, head, residual, padding_mode=pad, kernel=kernel,
                                               bin=(mode == "onepass"), chunk=7000)
            res[mode] = (loss.detach(), cells.grad, [p.grad for p in head.parameters()])
        for mode in ("jets", "onepass", "onepass_unbinned"):
            what = "%dD pad=%s %s " % (dim, pad, mode)
            assert_close_scaled(res[mode][0], res["dropin"][0], what + "loss", rtol=1e-4)
            assert_close_scaled(res[mode][1], res["dropin"][1], what + "cells.grad", rtol=1e-4, atol_scale=2e-5)
            for a, b in zip(res[mode][2], res["dropin"][2]):
                assert_close_scaled(a, b, what + "head grad", rtol=1e-4, atol_scale=2e-5)


def test_one_pass_step_saturated_tanh_and_nonfinite_coordinates(cuda):
    """Large |h| (tanh saturates: s1 -> 0 without NaN) and NaN / inf coordinates (contribute zero samples,
    exactly like the jets path)."""
    from cosinesampler_b200 import fused, jet
    gen = torch.Generator().manual_seed(9)
    shape = (4, 16, 32, 32)
    cells0 = torch.rand(shape, generator=gen) * 40.0 - 20.0            # pre-activations of order +-100
    coords = (torch.rand(4096, 2, generator=gen) * 2 - 1)
    coords[5] = float("nan")
    coords[6, 0] = float("inf")
    coords = coords.to(cuda).contiguous()
    res = {}
    for mode in ("jets", "onepass"):
        cells = torch.nn.Parameter(cells0.clone().to(cuda))
        head = make_head(16, seed=2).to(cuda)
        if mode == "jets":
            loss = jet.fused_pde_step(cells, coords, head, "helmholtz", mode="jets")
        else:
            loss = fused.one_pass_pde_step(cells, coords, head, "helmholtz")
        assert torch.isfinite(loss) and torch.isfinite(cells.grad).all()
        res[mode] = (loss.detach(), cells.grad, [p.grad for p in head.parameters()])
    assert_close_scaled(res["onepass"][0], res["jets"][0], "saturated loss", rtol=1e-4)
    assert_close_scaled(res["onepass"][1], res["jets"][1], "saturated cells.grad", rtol=1e-4, atol_scale=2e-5)
    for a, b in zip(res["onepass"][2], res["jets"][2]):
        assert_close_scaled(a, b, "saturated head grad", rtol=1e-4, atol_scale=2e-5)


def test_one_pass_step_config3_size_properties(cuda):
    """BASELINE config 3 at full size (cells [4,16,256,256], 2^20 points): binned + aggregated ==
    unbinned + direct reds, and the gradient is linear in the loss scale."""
    from cosinesampler_b200 import fused
    gen = torch.Generator().manual_seed(0)
    cells0 = torch.rand(4, 16, 256, 256, generator=gen).to(cuda)
    coords = (torch.rand(2 ** 20, 2, generator=gen) * 2 - 1).to(cuda)
    res = {}
    for mode, kw in (("agg", dict(bin=True, aggregate="auto")), ("direct", dict(bin=False, aggregate="off")),
                     ("agg_x3", dict(bin=True, aggregate="on", loss_scale=3.0))):
        cells = torch.nn.Parameter(cells0.clone())
        head = make_head(16, seed=0).to(cuda)
        loss = fused.one_pass_pde_step(cells, coords, head, "helmholtz", **kw)
        res[mode] = (loss.detach(), cells.grad, [p.grad for p in head.parameters()])
    assert_close_scaled(res["agg"][0], res["direct"][0], "cfg3 loss", rtol=1e-5)
    assert_close_scaled(res["agg"][1], res["direct"][1], "cfg3 cells.grad", rtol=1e-4, atol_scale=1e-5)
    assert_close_scaled(res["agg_x3"][1], 3.0 * res["agg"][1], "cfg3 linearity", rtol=1e-4, atol_scale=1e-5)
    for a, b in zip(res["agg"][2], res["direct"][2]):
        assert_close_scaled(a, b, "cfg3 head grad", rtol=1e-4, atol_scale=2e-5)


def test_one_pass_rejects_what_it_does_not_implement(cuda):
    from cosinesampler_b200 import fused
    cells = torch.rand(2, 8, 16, 16, device=cuda)
    coords = torch.rand(64, 2, device=cuda)
    wide = torch.nn.Sequential(torch.nn.Linear(8, 24), torch.nn.Tanh(), torch.nn.Linear(24, 1)).to(cuda)
    with pytest.raises(NotImplementedError):
        fused.one_pass_pde_step(cells, coords, wide)
    ok = _head(8, 16, 1, device=cuda)
    with pytest.raises(NotImplementedError):
        fused.one_pass_pde_step(cells, coords, ok, align_corners=False)
    with pytest.raises(RuntimeError):
        fused.one_pass_pde_step(cells.cpu(), coords, ok)
    many = torch.rand(40, 8, 8, 8, device=cuda)                      # more cells than the records have room for
    with pytest.raises(NotImplementedError):
        fused.one_pass_pde_step(many, coords, ok)
    from cosinesampler_b200 import jet
    assert jet.fused_mode(many, ok) == "jets" and jet.fused_mode(cells, ok) == "onepass"
    # an empty chunk is a no-op
    loss = fused.one_pass_pde_step(torch.nn.Parameter(cells), coords[:0], ok)
    assert float(loss) == 0.0


def test_bin_cache_reuses_and_invalidates(cuda):
    """cache_bins: the binned copy is reused while the coordinate tensor is unchanged and rebuilt after an
    in-place modification (version counter) -- same results either way."""
    from cosinesampler_b200 import fused
    gen = torch.Generator().manual_seed(3)
    cells0 = torch.rand(4, 16, 32, 32, generator=gen).to(cuda)
    coords = (torch.rand(5000, 2, generator=gen) * 2 - 1).to(cuda)
    head = make_head(16, seed=1).to(cuda)
    fused.bin_cache.clear()

    def run(cache):
        cells = torch.nn.Parameter(cells0.clone())
        for p in head.parameters():
            p.grad = None
        loss = fused.one_pass_pde_step(cells, coords, head, "helmholtz", cache_bins=cache)
        return loss.detach().clone(), cells.grad.clone()
    ref = run(False)
    a = run(True)
    assert len(fused.bin_cache.items) == 1
    binned_first = fused.bin_cache.items[0][2]
    b = run(True)
    assert fused.bin_cache.items[0][2] is binned_first                  # reused
    for x in (a, b):
        assert_close_scaled(x[0], ref[0], "cached loss", rtol=1e-5)
        assert_close_scaled(x[1], ref[1], "cached grad", rtol=1e-4, atol_scale=1e-5)
    coords.mul_(0.5)                                                    # in place: version changes
    ref2 = run(False)
    c = run(True)
    assert fused.bin_cache.items[-1][2] is not binned_first
    assert_close_scaled(c[0], ref2[0], "invalidated loss", rtol=1e-5)
    assert_close_scaled(c[1], ref2[1], "invalidated grad", rtol=1e-4, atol_scale=1e-5)
    fused.bin_cache.clear()
