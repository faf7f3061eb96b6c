"""Fused multi-cell jet sampler (SURVEY section 8f ranks 1 + 2; opt-in, not part of the reference's API).

The reference evaluates a PDE residual with nested `autograd.grad` calls through its operator:
per training step 1 forward, 5 first-backward, 6 double-backward and 2 triple-backward launches in
2D (1 / 7 / 9 / 3 in 3D; `modules_2d.py:38-111`), each streaming [N,C,P]
tensors, with the coordinates replicated over the N cells (`test_2d.py:38`) and the result summed
over the cells by the caller (`test_2d.py:51`).  All of those launches evaluate the same corner
gathers with different per-corner coefficients.  `SamplerJet2d / 3d` produce everything a second-
order residual needs in ONE gather pass and take the gradient back to the cells in ONE scatter pass:

    jets = SamplerJet2d.apply(cells, coords, 'zeros', True, 'cosine', True)     # [1 + 2*dim, C, P]
    z, z_x, z_y, z_xx, z_yy = jets                                              # each [C, P]

    cells  [N, C, (D,) H, W]      coords [P, dim]  (axis order of `grid[..., a]`: x, y(, z))

z = sum_n sample(cells[n], coords) is what `CosineSampler2d.apply(cells, grid).sum(0)` returns,
z_a / z_aa are its first / pure second derivatives along coordinate a (what the reference's
backward and double-backward kernels contract with the incoming gradient; mixed second
derivatives are not produced, as in the 2D reference, cu2d:675-678).  The caller applies the chain
rule of its head to the jets (`jet_mlp` below does it for Linear/Tanh stacks), so the training step
needs first-order autograd only; `jets.backward` scatters into `cells.grad` (the triple-backward
scatters of `modules_2d.py:98-111` in one kernel).  Gradients w.r.t. `coords` are not provided
(third-order coordinate derivatives: the reference returns None for them too, `mod2d:111`).

Needs C in {4, 8, 16, 32}; `align_corners` is honoured in 2D as well (the reference's 2D forward
ignores it, cu2d:307-308 -- with the default `True` there is no difference).
"""
import ctypes
import math

import torch

from . import _lib, ops
from .autograd import cell_offsets, padding_mode_enum, _kernel_enum, _require_kernel

SUPPORTED_CHANNELS = (4, 8, 16, 32)


def _check_args(input, coords, order):
    ops._check(input, "input")
    ops._check(coords, "coords")
    if order not in (1, 2, 3):
        raise ValueError("jet order must be 1, 2 or 3 (2 + mixed second derivatives), got %r" % (order,))
    if coords.dim() != 2 or coords.shape[1] not in (2, 3):
        raise RuntimeError("coords must be [P, dim] with dim 2 or 3, got %s" % (tuple(coords.shape),))
    dim = coords.shape[1]
    if input.dim() != dim + 2:
        raise RuntimeError("expected input [N,C,(D,)H,W] for %dD coords, got %s" % (dim, tuple(input.shape)))
    if input.shape[1] not in SUPPORTED_CHANNELS:
        raise RuntimeError("the jet operator needs C in %s, got %d (use CosineSampler%dd and .sum(0))"
                           % (SUPPORTED_CHANNELS, input.shape[1], dim))
    return dim


def _problem(input, coords, padding_mode, align_corners, kernel, multicell):
    dim = coords.shape[1]
    N, C = input.shape[:2]
    D, H, W = ((1,) + tuple(input.shape[2:])) if dim == 2 else tuple(input.shape[2:])
    return ops._problem(dim, N, C, D, H, W, coords.shape[0], padding_mode, align_corners, kernel, multicell,
                        _lib.LAYOUT_CHANNEL_LAST, 0)


def jet_count(dim, order):
    """Number of jets: value + dim first derivatives + (order >= 2) dim pure second derivatives + (order == 3)
    the dim(dim-1)/2 mixed ones, (x,y) in 2D and (x,y), (x,z), (y,z) in 3D -- what the reference's 3D double
    backward contracts (cu3d:836-856); its 2D kernels leave the mixed term out (cu2d:675-678)."""
    return 1 + min(order, 2) * dim + (dim * (dim - 1) // 2 if order >= 3 else 0)


def jet_bytes(dim, N, C, P, T, order):
    """Algorithmic bytes of one jet pass (forward or backward): coordinates + the jets + one field."""
    return 4 * (P * (dim + jet_count(dim, order) * C) + N * C * T)


def jet_forward(input, coords, offset, padding_mode, align_corners, kernel, multicell, order=2, staged=None):
    """jets [jet_count(dim, order), C, P] (cs_jet_forward).  `staged`: a channel-last copy of `input`
    (ops.stage) to reuse."""
    dim = _check_args(input, coords, order)
    ops._check(offset, "offset")
    field, layout = ops._field(input, staged)
    N, C = input.shape[:2]
    P = coords.shape[0]
    jets = torch.empty((jet_count(dim, order), C, P), dtype=input.dtype, device=input.device)
    pb = _problem(input, coords, padding_mode, align_corners, kernel, multicell)
    nbytes = jet_bytes(dim, N, C, P, input[0, 0].numel() if N and C else 0, order)
    with ops._on_device(input.device), ops._timed("JET%dd[fwd]" % dim, nbytes, input.device):
        rc = _lib.load().cs_jet_forward(pb, order, field.data_ptr(), coords.data_ptr(), offset.data_ptr(),
                                        jets.data_ptr(), ops._cur_stream(input.device))
    _lib.check(rc, "cs_jet_forward")
    return jets


def jet_backward_into(acc, gJets, input, coords, offset, padding_mode, align_corners, kernel, multicell,
                      order=2):
    """acc [N,T,C] (channel-last, caller-initialised) += adjoint of jet_forward applied to gJets.
    Several point chunks can share one accumulator (`new_accumulator` / `finish_accumulator`)."""
    dim = _check_args(input, coords, order)
    ops._check(offset, "offset")
    ops._check(gJets, "gJets", contiguous=False)
    N, C = input.shape[:2]
    P = coords.shape[0]
    if tuple(gJets.shape) != (jet_count(dim, order), C, P):
        raise RuntimeError("gJets must be %s, got %s" % ((jet_count(dim, order), C, P), tuple(gJets.shape)))
    gJets = gJets.contiguous()
    pb = _problem(input, coords, padding_mode, align_corners, kernel, multicell)
    nbytes = jet_bytes(dim, N, C, P, input[0, 0].numel() if N and C else 0, order)
    with ops._on_device(input.device), ops._timed("JET%dd[bwd]" % dim, nbytes, input.device):
        rc = _lib.load().cs_jet_backward(pb, order, gJets.data_ptr(), coords.data_ptr(), offset.data_ptr(),
                                         acc.data_ptr(), ops._cur_stream(input.device))
    _lib.check(rc, "cs_jet_backward")


def new_accumulator(input):
    """Zeroed channel-last gInput accumulator [N,T,C] for `jet_backward_into`."""
    return ops._new_accumulator(input, _lib.LAYOUT_CHANNEL_LAST)


def finish_accumulator(acc, input):
    """Channel-last accumulator -> gInput in the reference layout [N,C,(D,)H,W]."""
    return ops._finish_accumulator(acc, input, _lib.LAYOUT_CHANNEL_LAST)


def jet_backward(gJets, input, coords, offset, padding_mode, align_corners, kernel, multicell, order=2):
    """gInput [N,C,(D,)H,W] = adjoint of jet_forward applied to gJets (cs_jet_backward)."""
    acc = new_accumulator(input)
    jet_backward_into(acc, gJets, input, coords, offset, padding_mode, align_corners, kernel, multicell, order)
    return finish_accumulator(acc, input)


def _make_jet_function(dim):
    linear_name = "bilinear" if dim == 2 else "trilinear"

    class SamplerJet(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, coords, padding_mode="zeros", align_corners=True, kernel="cosine",
                    multicell=True, order=2):
            if coords.dim() != 2 or coords.shape[1] != dim:
                raise RuntimeError("SamplerJet%dd: coords must be [P, %d], got %s"
                                   % (dim, dim, tuple(coords.shape)))
            pm = padding_mode_enum(padding_mode)
            kn = _require_kernel(_kernel_enum(kernel, linear_name), kernel)
            coords = coords.detach().contiguous()
            offset = cell_offsets(input.shape[0], multicell, input.device)
            jets = jet_forward(input, coords, offset, pm, align_corners, kn, multicell, order)
            ctx.save_for_backward(input, coords, offset)
            ctx.args = (pm, align_corners, kn, multicell, order)
            return jets

        @staticmethod
        @torch.autograd.function.once_differentiable
        def backward(ctx, gJets):
            input, coords, offset = ctx.saved_tensors
            pm, align_corners, kn, multicell, order = ctx.args
            gInput = None
            if ctx.needs_input_grad[0] and gJets is not None:
                gInput = jet_backward(gJets, input, coords, offset, pm, align_corners, kn, multicell, order)
            return gInput, None, None, None, None, None, None

    SamplerJet.__name__ = SamplerJet.__qualname__ = "SamplerJet%dd" % dim
    return SamplerJet


SamplerJet2d = _make_jet_function(2)
SamplerJet3d = _make_jet_function(3)


# ---------------------------------------------------------------------------
# Chain rule of the caller's head, applied to jets (second-order Taylor mode)
# ---------------------------------------------------------------------------
def jet_mlp(head, jets, dim, order=2):
    """Propagate jets [J, C, P] of the head's input through `head`, an `nn.Sequential` of
    Linear / Tanh layers (the reference's test head, `test_2d.py:42-47`), and return
    (u, [u_a], [u_aa]) -- each [P, out_features] -- exactly what the nested
    `autograd.grad(u, x)`, `autograd.grad(u_x, x)` calls of `test_2d.py:55-127` produce; with order 3
    a fourth element {(a, b): u_ab} holds the mixed second derivatives (a < b).
    Plain torch ops in the jets' own feature-major layout ([features, P], no transposes):
    differentiable once more by autograd, which is all a training step needs."""
    J = jets.shape[0]
    assert J == jet_count(dim, order)
    pairs = [(a, b) for a in range(dim) for b in range(a + 1, dim)] if order >= 3 else []
    val = jets[0]                                   # [F, P]
    d1 = jets[1:1 + dim]                            # [dim, F, P]
    d2 = jets[1 + dim:1 + 2 * dim] if order >= 2 else None
    dm = jets[1 + 2 * dim:] if pairs else None      # [pairs, F, P]
    for layer in head:
        if isinstance(layer, torch.nn.Linear):
            w = layer.weight
            val = w @ val
            if layer.bias is not None:
                val = val + layer.bias[:, None]
            d1 = torch.matmul(w, d1)
            if d2 is not None:
                d2 = torch.matmul(w, d2)
            if dm is not None:
                dm = torch.matmul(w, dm)
        elif isinstance(layer, torch.nn.Tanh):
            t = torch.tanh(val)
            s1 = 1 - t * t                          # tanh'
            if d2 is not None:
                s2 = -2 * t * s1                    # tanh''
                if dm is not None:
                    dm = torch.stack([s2 * d1[a] * d1[b] + s1 * dm[m] for m, (a, b) in enumerate(pairs)])
                d2 = s2 * d1 * d1 + s1 * d2
            d1 = s1 * d1
            val = t
        else:
            raise NotImplementedError("jet_mlp supports Linear and Tanh layers, got %s" % type(layer).__name__)
    u = val.t()
    u_a = [d1[a].t() for a in range(dim)]
    u_aa = [d2[a].t() for a in range(dim)] if d2 is not None else None
    if order >= 3:
        return u, u_a, u_aa, {pr: dm[m].t() for m, pr in enumerate(pairs)}
    return u, u_a, u_aa


# ---------------------------------------------------------------------------
# Fused head + residual (cs_pde_head_step) and the fused training step
# ---------------------------------------------------------------------------
def residual_coefficients(residual, dim, k2=math.pi ** 2):
    """f = c_u u + c_u3 u^3 + sum_a (c1[a] u_a + c2[a] u_aa) for the residuals of `chain.pde_loss`
    (a in grid-channel order x, y(, z)), or pass a dict with those keys for another residual."""
    r = _lib.PdeResidual()
    if isinstance(residual, dict):
        r.c_u, r.c_u3 = float(residual.get("c_u", 0.0)), float(residual.get("c_u3", 0.0))
        for a in range(dim):
            r.c1[a] = float(residual.get("c1", [0.0] * dim)[a])
            r.c2[a] = float(residual.get("c2", [0.0] * dim)[a])
        return r
    if residual == "t2d":                            # test_2d.py:221
        if dim != 2:
            raise ValueError("the t2d residual is two-dimensional")
        r.c_u, r.c_u3 = -5.0, 5.0
        r.c1[1] = 2.0
        r.c2[0] = -0.0001
    elif residual in ("helmholtz", "laplace"):       # README Helmholtz; test_3d.py:270
        r.c_u = float(k2) if residual == "helmholtz" else 1.0
        for a in range(dim):
            r.c2[a] = 1.0
    else:
        raise ValueError(residual)
    return r


def _head_params(head, C):
    """(W1 [16,C], b1 [16], w2 [1,16], b2 [1]) of a Linear(C,16)-Tanh-Linear(16,1) head (test_2d.py:42-47)."""
    layers = list(head)
    ok = (len(layers) == 3 and isinstance(layers[0], torch.nn.Linear) and isinstance(layers[1], torch.nn.Tanh)
          and isinstance(layers[2], torch.nn.Linear) and layers[0].in_features == C
          and layers[0].out_features == 16 and layers[2].in_features == 16 and layers[2].out_features == 1
          and layers[0].bias is not None and layers[2].bias is not None)
    if not ok:
        raise NotImplementedError("the fused head is Linear(C,16)-Tanh-Linear(16,1) with biases; use "
                                  "jet_mlp (torch ops) for other heads")
    ps = (layers[0].weight, layers[0].bias, layers[2].weight, layers[2].bias)
    for t in ps:
        ops._check(t, "head parameter")
    return ps


def head_buffer_size(C):
    return 16 * C + 34


def head_buffer(C, device, dtype=torch.float32):
    """Zeroed accumulation buffer of cs_pde_head_step: gW1 [16,C] | gb1 [16] | gw2 [16] | gb2 [1] | loss_sum [1].
    The kernel adds into it, so one buffer can collect several chunks of points."""
    return torch.zeros(16 * C + 34, dtype=dtype, device=device)


def head_buffer_views(buf, C):
    """-> (loss_sum 0-dim, (gW1 [16,C], gb1 [16], gw2 [1,16], gb2 [1])) views of a `head_buffer`."""
    grads = (buf[:16 * C].view(16, C), buf[16 * C:16 * C + 16], buf[16 * C + 16:16 * C + 32].view(1, 16),
             buf[16 * C + 32:16 * C + 33])
    return buf[16 * C + 33], grads


def pde_head_step(jets, head, dim, residual="helmholtz", k2=math.pi ** 2, scale=1.0, in_place=False,
                  want_f=False, buf=None):
    """One launch of cs_pde_head_step on jets [1+2*dim, C, P]: returns
    (sum_p f^2 as a 0-dim tensor (unscaled), gJets, (gW1, gb1, gw2, gb2), f or None) where the
    gradients are those of  scale * sum_p f^2.  in_place=True overwrites `jets` with gJets.
    buf: a `head_buffer` to accumulate into (the returned loss / gradients are then running sums
    over every call that used it); default: a fresh one."""
    ops._check(jets, "jets")
    J, C, P = jets.shape
    if J != 1 + 2 * dim:
        raise RuntimeError("jets must be [1+2*dim, C, P], got %s" % (tuple(jets.shape),))
    W1, b1, w2, b2 = _head_params(head, C)
    res = residual_coefficients(residual, dim, k2)
    gJets = jets if in_place else torch.empty_like(jets)
    # one zeroed buffer for every accumulated output: gW1 | gb1 | gw2 | gb2 | loss_sum
    if buf is None:
        buf = head_buffer(C, jets.device, jets.dtype)
    elif buf.numel() != head_buffer_size(C) or buf.device != jets.device or buf.dtype != jets.dtype or not buf.is_contiguous():
        raise RuntimeError("buf must come from head_buffer(C, device)")
    f = torch.empty(P, dtype=jets.dtype, device=jets.device) if want_f else None
    base = buf.data_ptr()
    nbytes = 4 * 2 * J * C * P
    with ops._on_device(jets.device), ops._timed("HEAD%dd" % dim, nbytes, jets.device):
        rc = _lib.load().cs_pde_head_step(
            dim, C, P, jets.data_ptr(), W1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
            ctypes.byref(res), float(scale), gJets.data_ptr(), base, base + 4 * 16 * C,
            base + 4 * (16 * C + 16), base + 4 * (16 * C + 32), base + 4 * (16 * C + 33),
            f.data_ptr() if f is not None else None, ops._cur_stream(jets.device))
    _lib.check(rc, "cs_pde_head_step")
    loss_sum, grads = head_buffer_views(buf, C)
    return loss_sum, gJets, grads, f


class PdeHeadLoss(torch.autograd.Function):
    """loss = scale * sum_p f_p^2 as a differentiable (once) function of the jets and the head
    parameters: forward runs the fused kernel, which already produces every gradient."""

    @staticmethod
    def forward(ctx, jets, W1, b1, w2, b2, head, dim, residual, k2, scale):
        loss_sum, gJets, grads, _ = pde_head_step(jets.detach(), head, dim, residual, k2, scale)
        ctx.save_for_backward(gJets, *grads)
        return loss_sum * scale

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gl):
        gJets, gW1, gb1, gw2, gb2 = ctx.saved_tensors
        return gJets * gl, gW1 * gl, gb1 * gl, gw2 * gl, gb2 * gl, None, None, None, None, None


def pde_head_loss(jets, head, dim, residual="helmholtz", k2=math.pi ** 2, scale=None):
    """mean_p f_p^2 (scale defaults to 1/P), differentiable w.r.t. `jets` and the head parameters."""
    if scale is None:
        scale = 1.0 / max(1, jets.shape[-1])
    W1, b1, w2, b2 = _head_params(head, jets.shape[1])
    return PdeHeadLoss.apply(jets, W1, b1, w2, b2, head, dim, residual, k2, scale)


def _add_grad(param, g, owned=False):
    if param.grad is None:
        param.grad = g if owned else g.clone()
    else:
        param.grad.add_(g)


class FusedPdeStep:
    """A PIXEL training step without autograd, fed chunk by chunk:

        step = FusedPdeStep(cells, head, residual="helmholtz", kernel="cosine")
        step.begin()                      # stage `cells` channel-last once, zero one accumulator
        for xy in chunks:                 # [p, dim] each
            step.add(xy, 1.0 / P_total)   # jets -> head/residual/gradients -> scatter: 3 launches
        loss = step.finish()              # cells.grad, head .grad accumulated; 0-dim loss tensor

    which is what `chain.training_step` does with `loss.backward()` through the drop-in operator
    (14 operator launches and ~200 torch kernels per chunk).  Nothing synchronises with the host."""

    def __init__(self, cells, head, residual="helmholtz", k2=math.pi ** 2, padding_mode="zeros",
                 align_corners=True, kernel="cosine", multicell=True):
        ops._check(cells, "input")
        self.dim = cells.dim() - 2
        if self.dim not in (2, 3):
            raise RuntimeError("expected cells [N,C,(D,)H,W], got %s" % (tuple(cells.shape),))
        self.cells, self.head = cells, head
        self.residual, self.k2 = residual, k2
        self.pm = padding_mode_enum(padding_mode)
        self.kn = _require_kernel(_kernel_enum(kernel, "bilinear" if self.dim == 2 else "trilinear"), kernel)
        if self.dim == 2 and not align_corners:
            raise NotImplementedError(
                "2D with align_corners=False: the reference's 2D forward ignores the flag (cu2d:307-308) while its "
                "backward kernels honour it; the fused step refuses that inconsistent combination "
                "(SamplerJet2d honours the flag consistently and says so)")
        self.align_corners, self.multicell = align_corners, multicell
        self.params = _head_params(head, cells.shape[1])
        self._live = False

    def begin(self, reducer=None, scale=None):
        """reducer: a `peer.PeerReducer` whose symmetric-memory accumulators receive the scatters, so
        that `finish` can sum over the ranks with one kernel over NVLink instead of NCCL.
        scale: the step's loss scale, when known up front (a rank whose shard is empty never calls `add`
        but must still scale the reduced loss like the others)."""
        with torch.no_grad():
            self.cells_d = self.cells.detach()
            self.offset = cell_offsets(self.cells.shape[0], self.multicell, self.cells.device)
            self.staged = ops.stage(self.cells_d)
            self.reducer = reducer
            if reducer is not None:
                self.acc = reducer.accumulator()
                self.buf = reducer.small_buffer()
            else:
                self.acc = new_accumulator(self.cells_d)
                # the head kernel adds into this buffer: loss and head gradients of all chunks, no torch ops
                self.buf = head_buffer(self.cells.shape[1], self.cells.device)
        self.scale, self._live = (None if scale is None else float(scale)), True

    def add(self, xy, scale):
        """One chunk of points xy [p, dim]; its loss contribution is scale * sum_p f^2 (the same
        `scale` for every chunk of a step, e.g. loss_scale / total points)."""
        if not self._live:
            raise RuntimeError("FusedPdeStep.add before begin()")
        if self.scale is None:
            self.scale = float(scale)
        elif abs(float(scale) - self.scale) > 1e-12 * abs(self.scale):
            raise RuntimeError("FusedPdeStep.add: every chunk of a step must use the same scale")
        _check_args(self.cells_d, xy, 2)
        if xy.shape[1] != self.dim:
            raise RuntimeError("coords must be [p, %d], got %s" % (self.dim, tuple(xy.shape)))
        with torch.no_grad():
            a = (self.pm, self.align_corners, self.kn, self.multicell, 2)
            jets = jet_forward(self.cells_d, xy, self.offset, *a, staged=self.staged)
            _, gJets, _, _ = pde_head_step(jets, self.head, self.dim, self.residual, self.k2, scale,
                                           in_place=True, buf=self.buf)
            jet_backward_into(self.acc, gJets, self.cells_d, xy, self.offset, *a)

    def finish(self):
        if not self._live:
            raise RuntimeError("FusedPdeStep.finish before begin()")
        self._live = False
        with torch.no_grad():
            if self.reducer is not None:
                # sum over the ranks + layout change in one kernel over peer memory; the loss and the
                # head gradients returned are those of ALL ranks
                gcells, buf = self.reducer.reduce()
                if self.cells.requires_grad:
                    _add_grad(self.cells, gcells)
            else:
                buf = self.buf
                if self.cells.requires_grad:
                    _add_grad(self.cells, finish_accumulator(self.acc, self.cells_d), owned=True)
            loss_sum, pgrads = head_buffer_views(buf, self.cells.shape[1])
            for prm, g in zip(self.params, pgrads):
                if prm.requires_grad:
                    _add_grad(prm, g)
            loss = loss_sum * (self.scale if self.scale is not None else 0.0)
        self.acc = self.staged = self.buf = self.reducer = None
        return loss


def head_is_fusable(head, C):
    """True when `head` is the Linear(C,16)-Tanh-Linear(16,1) stack cs_pde_head_step implements."""
    try:
        _head_params(head, C)
        return True
    except NotImplementedError:
        return False


def jet_autograd_step(cells, coords, head, residual="helmholtz", k2=math.pi ** 2, padding_mode="zeros",
                      align_corners=True, kernel="cosine", multicell=True, chunk=None, loss_scale=1.0):
    """The same step for any Linear/Tanh head: jets from `SamplerJet2d/3d` (one gather pass, one
    scatter pass), the head's chain rule and the residual in torch (`jet_mlp`), first-order autograd."""
    from .chain import _residual
    if isinstance(residual, dict):
        raise NotImplementedError("a residual given as coefficients needs the fused head; pass a name")
    dim = coords.shape[1]
    S = SamplerJet2d if dim == 2 else SamplerJet3d
    P = coords.shape[0]
    chunk = max(1, P if not chunk else min(chunk, P))
    params = [cells] + [q for q in head.parameters() if q.requires_grad]
    total = None
    for s in range(0, P, chunk):
        xy = coords[s:s + chunk]
        jets = S.apply(cells, xy, padding_mode, align_corners, kernel, multicell)
        u, first, second = jet_mlp(head, jets, dim)
        loss = torch.sum(_residual(u, first, second, residual, k2) ** 2) * (loss_scale / P)
        loss.backward(inputs=params)
        total = loss.detach() if total is None else total + loss.detach()
    return total if total is not None else torch.zeros((), device=cells.device)


_warned_fallback = set()


def fused_mode(cells, head, align_corners=True):
    """Which implementation `fused_pde_step(mode='auto')` takes for this head:
    'onepass' (`fused.OnePassPdeStep`: Linear(C,K)-Tanh-Linear(K,1), K in {4,8,16,32,64}; one kernel per chunk),
    'jets' (`FusedPdeStep`: jets -> tensor-core head -> scatter, Linear(C,16)-Tanh-Linear(16,1)), or
    'torch_head' (`jet_autograd_step`: jet kernels + the head's chain rule in torch ops)."""
    from . import fused
    C = cells.shape[1]
    if fused.head_is_fusable(head, C) and (cells.dim() == 5 or align_corners) and cells.shape[0] <= fused.MAX_CELLS:
        return "onepass"
    if head_is_fusable(head, C) and C in SUPPORTED_CHANNELS:
        return "jets"
    return "torch_head"


def fused_pde_step(cells, coords, head, residual="helmholtz", k2=math.pi ** 2, padding_mode="zeros",
                   align_corners=True, kernel="cosine", multicell=True, chunk=None, loss_scale=1.0, reducer=None,
                   mode="auto", cache_bins=False):
    """One training step over coords [P, dim] in chunks of `chunk` points without nested autograd:
    accumulates `cells.grad` and the head parameters' `.grad`, returns loss_scale * mean_p f^2 as a
    0-dim tensor.  mode: 'auto' (see `fused_mode`), 'onepass', 'jets' or 'torch_head'.  Falling back to
    the torch head is announced once per head shape (it is several times slower)."""
    ops._check(cells, "input")
    if cells.dim() not in (4, 5):
        raise RuntimeError("expected cells [N,C,(D,)H,W], got %s" % (tuple(cells.shape),))
    if mode == "auto":
        mode = fused_mode(cells, head, align_corners)
        if mode == "torch_head":
            key = tuple((type(l).__name__, getattr(l, "in_features", 0), getattr(l, "out_features", 0)) for l in head)
            if key not in _warned_fallback:
                _warned_fallback.add(key)
                import warnings
                warnings.warn("cosinesampler_b200: head %s is not Linear(C,K)-Tanh-Linear(K,1) with K in {4,8,16,32,64}; "
                              "fused_pde_step runs the jet kernels with the head in torch ops (slower)" % (key,))
    if mode == "onepass":
        from . import fused
        return fused.one_pass_pde_step(cells, coords, head, residual, k2, padding_mode, align_corners, kernel,
                                       multicell, chunk, loss_scale, reducer, cache_bins=cache_bins)
    if mode == "torch_head":
        if reducer is not None:
            raise NotImplementedError("the peer-memory reduce needs a fused head")
        return jet_autograd_step(cells, coords, head, residual, k2, padding_mode, align_corners, kernel,
                                 multicell, chunk, loss_scale)
    if mode != "jets":
        raise ValueError("mode must be 'auto', 'onepass', 'jets' or 'torch_head', got %r" % (mode,))
    step = FusedPdeStep(cells, head, residual, k2, padding_mode, align_corners, kernel, multicell)
    P = coords.shape[0]
    chunk = max(1, P if not chunk else min(chunk, P))
    step.begin(reducer, scale=loss_scale / P if P else 0.0)
    for s in range(0, P, chunk):
        step.add(coords[s:s + chunk], loss_scale / P)
    return step.finish()


__all__ = ["jet_count", "fused_mode", "SamplerJet2d", "SamplerJet3d", "jet_forward", "jet_backward", "jet_backward_into", "jet_mlp",
           "jet_bytes", "pde_head_step", "pde_head_loss", "fused_pde_step", "FusedPdeStep", "jet_autograd_step",
           "head_is_fusable",
           "residual_coefficients"]
