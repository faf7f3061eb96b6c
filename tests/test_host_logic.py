"""Host-side logic that needs no GPU: enums, offsets, argument checks, stream views,
dead-output queries, point sharding."""
import pytest
import torch

import cosine_sampler_2d
import cosine_sampler_3d
from cosinesampler_b200 import autograd as ag
from cosinesampler_b200 import dp, ops
from cosinesampler_b200.modules_2d import kernel_enum as kernel_enum_2d
from cosinesampler_b200.modules_3d import kernel_enum as kernel_enum_3d


def test_reference_module_paths_and_names():
    # reference import paths (cosine_sampler_2d/__init__.py:1, cosine_sampler_3d/__init__.py:1)
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    from cosine_sampler_2d.modules_2d import (CosineSamplerBackward, CosineSamplerBackwardBackward,
                                              padding_mode_enum, kernel_enum)
    assert CosineSampler2d.__name__ == "CosineSampler2d"
    assert CosineSampler3d.__name__ == "CosineSampler3d"
    assert issubclass(CosineSampler2d, torch.autograd.Function)
    assert issubclass(CosineSamplerBackward, torch.autograd.Function)
    assert issubclass(CosineSamplerBackwardBackward, torch.autograd.Function)
    assert cosine_sampler_3d.modules_3d.CosineSamplerBackward is not CosineSamplerBackward


def test_enums_follow_the_reference():
    pm = ag.padding_mode_enum
    assert (pm("zeros"), pm("border"), pm("reflection"), pm("anything")) == (0, 1, 2, 2)   # mod2d:4-10
    assert [kernel_enum_2d(k) for k in ("cosine", "bilinear", "smooth-step")] == [0, 1, 2]  # mod2d:12-18
    assert [kernel_enum_3d(k) for k in ("cosine", "trilinear", "smooth-step")] == [0, 1, 2]  # mod3d:12-18
    assert kernel_enum_2d("trilinear") is None and kernel_enum_3d("bilinear") is None
    assert kernel_enum_2d("smoothstep") is None


def test_offsets_are_the_reference_linspace():
    for n in (1, 2, 4, 50, 96):
        want = torch.linspace(0, 1 - (1 / n), n)            # mod2d:25
        got = ag.cell_offsets(n, True, torch.device("cpu"))
        assert torch.equal(got, want)
        assert torch.equal(ag.cell_offsets(n, False, torch.device("cpu")), torch.zeros(n))


def test_cpu_tensors_fail_loudly():
    x = torch.rand(1, 4, 8, 8)
    g = torch.rand(1, 1, 5, 2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        cosine_sampler_2d.CosineSampler2d.apply(x, g)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        cosine_sampler_3d.CosineSampler3d.apply(torch.rand(1, 4, 4, 4, 4), torch.rand(1, 1, 1, 5, 3))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ops.forward(x, g, torch.zeros(1), 0, True, 0, True)


def test_engine_wants_prunes_dead_outputs():
    seen = []

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, b):
            return a * b

        @staticmethod
        def backward(ctx, g):
            seen.append((ag._engine_wants(ctx, 0), ag._engine_wants(ctx, 1)))
            return g, g

    a = torch.randn(3, requires_grad=True)
    b = torch.randn(3, requires_grad=True)
    w = torch.randn(3, requires_grad=True)
    out = Probe.apply(a * w, b * 2).sum()
    torch.autograd.grad(out, a, retain_graph=True)
    torch.autograd.grad(out, b, retain_graph=True)
    torch.autograd.grad(out, [a, b], retain_graph=True)
    out.backward(retain_graph=True)
    out.backward(inputs=[w])
    assert seen == [(True, False), (False, True), (True, True), (True, True), (True, False)]
    # leaves directly on the edges (cells are a leaf Parameter in PIXEL)
    seen.clear()
    out = Probe.apply(a, b * 2).sum()
    torch.autograd.grad(out, b, retain_graph=True)
    assert seen == [(False, True)]


def test_shard_range_partitions_points():
    for P in (0, 1, 7, 2 ** 20, 2 ** 25 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [dp.shard_range(P, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_record_slots_of_the_one_pass_kernel_are_bank_conflict_free():
    """cs_fused.cuh `rec_slot(i) = i + (i >> 3)`: the L lanes of a walker read the same 16-byte record and the walkers
    of a warp read records PPQ points apart.  Model of an LDS.128 / STS.128: identical addresses are merged over
    the warp, the rest is served in wavefronts of one address per 16-byte bank group (8 groups).  With slot = point
    the reads of the K = 16 kernel need 4 wavefronts, with the spare slot per 8 points the minimum for every hidden
    width, and the one-point-per-lane writes of phase 1 stay at their minimum too."""
    from collections import Counter

    def wavefronts(slots):
        return max(Counter(s % 8 for s in set(slots)).values())

    for dim, ppq in ((2, 4), (3, 2)):
        for lshift in range(5):
            nw = 32 >> lshift
            pts = ppq * nw
            ideal_read = max(1, nw * 16 // 128)
            for name, slot in (("plain", lambda i: i), ("padded", lambda i: i + (i >> 3))):
                reads = [wavefronts([slot(ppq * (lane >> lshift) + t) for lane in range(32)]) for t in range(ppq)]
                lanes_used = min(32, pts * max(1, 32 // pts))
                writes = [wavefronts([slot((u * 32 + lane) % pts) for lane in range(lanes_used)])
                          for u in range((pts + 31) // 32)]
                if name == "padded":
                    assert all(r == ideal_read for r in reads), (dim, lshift, reads)
                    assert all(w <= 4 for w in writes)
                elif dim == 2 and lshift == 2:
                    assert reads == [4, 4, 4, 4]                # what ncu showed: 40 % of the wavefronts were conflicts
