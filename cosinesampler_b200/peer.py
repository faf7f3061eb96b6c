"""Gradient all-reduce over peer memory, fused with the layout change.

The fused step scatters `d loss / d cells` into a channel-last accumulator and has to (a) sum it over
the ranks and (b) transpose it back to the reference layout.  With NCCL that is a transpose, a flat
copy, a 16 MiB ring all-reduce (~0.2 ms on 8 B200s, latency-bound) and a copy back.  Here the
accumulator lives in symmetric memory (torch.distributed._symmetric_memory: allocations mapped into
every peer over NVLink / NVSwitch) and one kernel, `cs_peer_allreduce_from_channel_last`, does
reduce-scatter + all-gather + transpose: each rank owns every world-th tile, sums it over all peers'
accumulators, and stores the channel-first result into all peers' outputs.  The head's gradients
and the loss ride along as a small vector.  torch is used for the symmetric allocations and the two
barriers that bracket the kernel; the data movement is ours.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops

MAX_PEERS = 8


class PeerReducer:
    """Symmetric buffers + the fused reduce for one `cells` shape and one process group."""

    def __init__(self, cells, n_small, group=None, channels=None, tail=0, transpose=True, multicast=None):
        """channels: channel count of the accumulator when it differs from the cells' (the one-pass step
        scatters into the W1-mixed cells: K hidden units per texel instead of C channels).
        tail: extra floats allocated behind the accumulator (the one-pass step's dump texel); they are zeroed
        with it and never reduced.
        transpose: True = the reduced output is channel-first [N,C,*S] (cs_peer_allreduce_from_channel_last);
        False = it keeps the accumulator's channel-last layout [N,T,C] (cs_peer_allreduce), and the sum is taken
        inside the NVSwitch with multimem.ld_reduce / multimem.st when the allocations have a multicast
        address (multicast=None: use it when available; False: peer loads; COSINE_SAMPLER_MULTIMEM=0 too)."""
        import os
        import torch.distributed._symmetric_memory as symm_mem
        if not dist.is_initialized():
            raise RuntimeError("PeerReducer needs an initialised process group")
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        if self.world > MAX_PEERS:
            raise RuntimeError("PeerReducer supports at most %d ranks, got %d" % (MAX_PEERS, self.world))
        ops._check(cells, "input")
        self.N = cells.shape[0]
        self.C = int(channels) if channels is not None else cells.shape[1]
        self.shape = (self.N, self.C) + tuple(cells.shape[2:])
        self.T = cells[0, 0].numel()
        self.n_small = int(n_small)
        dev = cells.device
        n = self.N * self.T * self.C
        self.acc = symm_mem.empty(n + int(tail), dtype=torch.float32, device=dev)
        self._n = n
        self.out = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.small = symm_mem.empty(max(self.n_small, 4), dtype=torch.float32, device=dev)
        self.h_acc = symm_mem.rendezvous(self.acc, group)
        self.h_out = symm_mem.rendezvous(self.out, group)
        self.h_small = symm_mem.rendezvous(self.small, group)
        self.rank = self.h_acc.rank
        self.transposed = bool(transpose)
        if n % 4 != 0 and not self.transposed:
            raise RuntimeError("the layout-preserving peer reduce needs N*T*C to be a multiple of 4")
        self.acc_mc = self.out_mc = 0
        if not self.transposed and multicast is not False and os.environ.get("COSINE_SAMPLER_MULTIMEM", "1") != "0":
            try:
                a, o = int(self.h_acc.multicast_ptr), int(self.h_out.multicast_ptr)
            except Exception:
                a = o = 0
            # every rank must take the same path: multimem only when all ranks have both addresses
            flag = torch.tensor([1.0 if (a and o) else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if float(flag.item()) > 0.5:
                self.acc_mc, self.out_mc = a, o
        self.kind = ("cs_peer_allreduce_from_channel_last (peer loads + layout change)" if self.transposed else
                     "cs_peer_allreduce (multimem.ld_reduce + multimem.st through the NVSwitch)" if self.acc_mc else
                     "cs_peer_allreduce (16-byte peer loads / stores)")
        self.small_out = torch.zeros(max(self.n_small, 4), dtype=torch.float32, device=dev)
        arr = ctypes.c_void_p * MAX_PEERS
        self._acc_ptrs = arr(*[int(x) for x in self.h_acc.buffer_ptrs])
        self._out_ptrs = arr(*[int(x) for x in self.h_out.buffer_ptrs])
        self._small_ptrs = arr(*[int(x) for x in self.h_small.buffer_ptrs])
        self.acc.zero_()
        self.small.zero_()
        # nobody may start scattering into / reading from a peer before everyone has zeroed
        self.h_acc.barrier(channel=0)

    def accumulator(self):
        """The zeroed channel-last accumulator [N, T, C] of this step (zeroed after the previous reduce)."""
        return self.acc[:self._n].view(self.N, self.T, self.C)

    def small_buffer(self):
        """The zeroed small vector (head gradients | loss) of this step."""
        return self.small[:self.n_small]

    def reduce(self):
        """-> (sum over ranks of the accumulators -- [N,C,*S] when `transposed`, else [N,T,C]; a view of the
        symmetric output buffer, valid until the next reduce --, sum over ranks of the small vectors).
        Stream-ordered; no host sync."""
        dev = self.acc.device
        self.h_acc.barrier(channel=0)                 # every rank has finished scattering
        with ops._on_device(dev):
            if self.transposed:
                rc = _lib.load().cs_peer_allreduce_from_channel_last(
                    self.world, self.rank, self._acc_ptrs, self._out_ptrs, self.N, self.C, self.T,
                    self._small_ptrs, self.small_out.data_ptr(), self.n_small, ops._cur_stream(dev))
            else:
                rc = _lib.load().cs_peer_allreduce(
                    self.world, self.rank, self._acc_ptrs, self._out_ptrs, self._n,
                    self.acc_mc or None, self.out_mc or None,
                    self._small_ptrs, self.small_out.data_ptr(), self.n_small, ops._cur_stream(dev))
        _lib.check(rc, "peer all-reduce")
        self.h_acc.barrier(channel=1)                 # every rank has finished reading / writing peers
        self.acc.zero_()                              # ready for the next step
        self.small.zero_()
        out = self.out.view(self.shape) if self.transposed else self.out.view(self.N, self.T, self.C)
        return out, self.small_out[:self.n_small]
