"""cs_peer_allreduce_from_channel_last on ONE device: the W "ranks" are W separate buffers on cuda:0 and the
kernel is launched once per rank with the same pointer tables (each launch reduces the tiles that rank owns and
stores them into every rank's output).  Launches are stream-ordered, none waits for another, so this is the
multi-rank data path without the symmetric-memory rendezvous -- the part `tests/test_gpu_peer.py` can only
check on a multi-GPU box."""
import ctypes

import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(4, 16, 257), (3, 4, 1000), (2, 20, 96), (1, 32, 33), (5, 1, 7)])
def test_peer_reduce_single_device(cuda, world, shape):
    from cosinesampler_b200 import _lib, ops
    N, C, T = shape
    gen = torch.Generator().manual_seed(world * 100 + C)
    accs = [torch.randn(N, T, C, generator=gen).to(cuda) for _ in range(world)]
    outs = [torch.full((N, C, T), float("nan"), device=cuda) for _ in range(world)]
    n_small = 37
    smalls = [torch.randn(n_small, generator=gen).to(cuda) for _ in range(world)]
    small_outs = [torch.zeros(n_small, device=cuda) for _ in range(world)]
    arr = ctypes.c_void_p * 8
    acc_ptrs = arr(*[a.data_ptr() for a in accs])
    out_ptrs = arr(*[o.data_ptr() for o in outs])
    small_ptrs = arr(*[s.data_ptr() for s in smalls])
    lib = _lib.load()
    for r in range(world):
        rc = lib.cs_peer_allreduce_from_channel_last(world, r, acc_ptrs, out_ptrs, N, C, T, small_ptrs,
                                                     small_outs[r].data_ptr(), n_small,
                                                     ops._cur_stream(cuda))
        _lib.check(rc, "cs_peer_allreduce_from_channel_last")
    torch.cuda.synchronize()
    ref = torch.stack([a.double() for a in accs]).sum(0).permute(0, 2, 1)          # [N, C, T]
    sref = torch.stack([s.double() for s in smalls]).sum(0)
    for r in range(world):
        assert torch.isfinite(outs[r]).all(), "rank %d: tiles missing" % r
        assert_close_scaled(outs[r], ref, "world=%d rank=%d sum" % (world, r), rtol=1e-6, atol_scale=1e-6)
        assert torch.equal(outs[r], outs[0]), "ranks disagree bit-wise"
        assert_close_scaled(small_outs[r], sref, "world=%d rank=%d small" % (world, r), rtol=1e-6, atol_scale=1e-6)
    # same result as the single-rank path: sum, then cs_from_channel_last
    one = ops.from_channel_last(torch.stack(accs).sum(0), (N, C, T))
    assert_close_scaled(outs[0], one, "vs from_channel_last", rtol=1e-6, atol_scale=1e-6)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_flat_peer_allreduce_single_device(cuda, world):
    """cs_peer_allreduce (layout-preserving, 16-byte peer loads; the multimem path needs NVSwitch multicast
    objects and is exercised by the multi-GPU bench) with the ranks emulated as buffers on one device."""
    from cosinesampler_b200 import _lib, ops
    gen = torch.Generator().manual_seed(world)
    for n in (4, 1000, 4 * 16 * 257 * 4):
        accs = [torch.randn(n, generator=gen).to(cuda) for _ in range(world)]
        outs = [torch.full((n,), float("nan"), device=cuda) for _ in range(world)]
        smalls = [torch.randn(290, generator=gen).to(cuda) for _ in range(world)]
        small_outs = [torch.zeros(290, device=cuda) for _ in range(world)]
        arr = ctypes.c_void_p * 8
        acc_ptrs, out_ptrs = arr(*[a.data_ptr() for a in accs]), arr(*[o.data_ptr() for o in outs])
        small_ptrs = arr(*[s.data_ptr() for s in smalls])
        lib = _lib.load()
        for r in range(world):
            rc = lib.cs_peer_allreduce(world, r, acc_ptrs, out_ptrs, n, None, None, small_ptrs,
                                       small_outs[r].data_ptr(), 290, ops._cur_stream(cuda))
            _lib.check(rc, "cs_peer_allreduce")
        torch.cuda.synchronize()
        ref = torch.stack([a.double() for a in accs]).sum(0)
        for r in range(world):
            assert_close_scaled(outs[r], ref, "flat world=%d n=%d rank=%d" % (world, n, r), rtol=1e-6, atol_scale=1e-6)
            assert torch.equal(outs[r], outs[0])
            assert_close_scaled(small_outs[r], torch.stack([s.double() for s in smalls]).sum(0), "flat small",
                                rtol=1e-6, atol_scale=1e-6)
    # argument validation
    assert lib.cs_peer_allreduce(2, 0, acc_ptrs, out_ptrs, 6, None, None, None, None, 0, None) == -1
    assert lib.cs_peer_allreduce(2, 0, acc_ptrs, out_ptrs, 8, accs[0].data_ptr(), None, None, None, 0, None) == -1
