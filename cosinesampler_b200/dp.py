"""Data parallelism over collocation points: the points are
partitioned across ranks, the cells (and the small head) are replicated, and the
only exchange of a step is one SUM all-reduce of `cells.grad` plus the head's
gradients, packed in a single flat bucket.  One process per GPU; NCCL over
NVLink/NVSwitch on B200, gloo in the CPU tests.

The reference is single-process, single-device (no torch.distributed anywhere);
this is what the north star adds around the operator.
"""
import torch
import torch.distributed as dist


def shard_range(num_points, rank, world_size):
    """Contiguous, balanced split of [0, num_points): the first `num_points % world`
    ranks hold one extra point."""
    base, extra = divmod(num_points, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_grads(params, group=None):
    """Sum the `.grad` of `params` over the ranks of `group`, in place, with a single
    collective on one flat fp32 bucket.  Params without a grad contribute zeros."""
    params = [p for p in params]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    if len(grads) == 1 and grads[0].is_contiguous():
        dist.all_reduce(grads[0], op=dist.ReduceOp.SUM, group=group)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class PointShardedStep:
    """Runs `chain.training_step` on this rank's shard of the points and reduces the
    gradients.  loss_scale = shard/total makes the summed gradient the gradient of the
    mean over *all* points."""

    def __init__(self, sampler, cells, head, residual="helmholtz", chunk=None, group=None, fused=None,
                 peer_reduce=False, peer_multicast=None):
        """fused: None = `sampler` is a drop-in operator driven through `chain.training_step`
        (nested autograd, the reference's call pattern); a dict of `jet.fused_pde_step` keyword
        arguments (kernel=..., multicell=...) = the fused jet path (`sampler` is ignored).
        peer_reduce (fused path only): sum the gradients over the ranks with the fused peer-memory
        kernel (`peer.PeerReducer`, NVLink symmetric memory) instead of NCCL; the loss returned by
        `step` is then the loss over ALL ranks' points (`loss_is_global`).  peer_multicast: None = sum inside
        the NVSwitch (multimem) when the symmetric allocations have a multicast address, False = peer loads."""
        self.sampler, self.cells, self.head = sampler, cells, head
        self.residual, self.chunk, self.group = residual, chunk, group
        self.fused = fused
        self.reducer = None
        self.loss_is_global = False
        self.mode = None
        if fused is not None:
            from .jet import fused_mode
            self.mode = fused.get("mode", "auto")
            if self.mode == "auto":
                self.mode = fused_mode(cells, head, fused.get("align_corners", True))
        if peer_reduce and fused is not None and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(group) > 1:
            from .peer import PeerReducer
            if self.mode == "onepass":
                # the one-pass step scatters into the W1-mixed cells: K hidden units per texel
                from .fused import head_params, small_buffer_size
                K = head_params(head, cells.shape[1])[0].shape[0]
                self.reducer = PeerReducer(cells, small_buffer_size(cells.shape[1], K), group, channels=K, tail=K,
                                           transpose=False, multicast=peer_multicast)
            elif self.mode == "jets":
                from .jet import head_buffer_size
                self.reducer = PeerReducer(cells, head_buffer_size(cells.shape[1]), group)
            else:
                raise NotImplementedError("the peer-memory reduce needs a fused head")
            self.loss_is_global = True

    def params(self):
        return [self.cells] + list(self.head.parameters())

    def zero_grad(self):
        for p in self.params():
            p.grad = None

    def step(self, local_coords, total_points):
        if self.fused is not None:
            from .jet import fused_pde_step
            xy = local_coords if torch.is_tensor(local_coords) else torch.cat(list(local_coords), -1)
            kw = dict(self.fused)
            kw["mode"] = self.mode
            loss = fused_pde_step(self.cells, xy.contiguous(), self.head, self.residual, chunk=self.chunk,
                                  loss_scale=xy.shape[0] / float(total_points), reducer=self.reducer, **kw)
            if self.reducer is None:
                allreduce_grads(self.params(), self.group)
            return loss
        from .chain import training_step
        local = local_coords[0].shape[0]
        loss = training_step(self.sampler, self.cells, local_coords, self.head, self.residual,
                             chunk=self.chunk, loss_scale=local / float(total_points))
        allreduce_grads(self.params(), self.group)
        return loss
