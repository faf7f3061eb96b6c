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
import torch

from . import _lib, ops
from .autograd import cell_offsets, padding_mode_enum, _kernel_enum, _require_kernel

SUPPORTED_CHANNELS = (4, 8, 16, 32)


def _check_args(input, coords, order):
    ops._check(input, "input")
    ops._check(coords, "coords")
    if order not in (1, 2):
        raise ValueError("jet order must be 1 or 2, got %r" % (order,))
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


def jet_bytes(dim, N, C, P, T, order):
    """Algorithmic bytes of one jet pass (forward or backward): coordinates + the jets + one field."""
    return 4 * (P * (dim + (1 + order * dim) * C) + N * C * T)


def jet_forward(input, coords, offset, padding_mode, align_corners, kernel, multicell, order=2, staged=None):
    """jets [1 + order*dim, C, P] (cs_jet_forward).  `staged`: a channel-last copy of `input`
    (ops.stage) to reuse."""
    dim = _check_args(input, coords, order)
    ops._check(offset, "offset")
    field, layout = ops._field(input, staged)
    N, C = input.shape[:2]
    P = coords.shape[0]
    jets = torch.empty((1 + order * dim, C, P), dtype=input.dtype, device=input.device)
    pb = _problem(input, coords, padding_mode, align_corners, kernel, multicell)
    nbytes = jet_bytes(dim, N, C, P, input[0, 0].numel() if N and C else 0, order)
    with ops._on_device(input.device), ops._timed("JET%dd[fwd]" % dim, nbytes, input.device):
        rc = _lib.load().cs_jet_forward(pb, order, field.data_ptr(), coords.data_ptr(), offset.data_ptr(),
                                        jets.data_ptr(), ops._cur_stream(input.device))
    _lib.check(rc, "cs_jet_forward")
    return jets


def jet_backward(gJets, input, coords, offset, padding_mode, align_corners, kernel, multicell, order=2):
    """gInput [N,C,(D,)H,W] = adjoint of jet_forward applied to gJets (cs_jet_backward)."""
    dim = _check_args(input, coords, order)
    ops._check(offset, "offset")
    ops._check(gJets, "gJets", contiguous=False)
    N, C = input.shape[:2]
    P = coords.shape[0]
    if tuple(gJets.shape) != (1 + order * dim, C, P):
        raise RuntimeError("gJets must be %s, got %s" % ((1 + order * dim, C, P), tuple(gJets.shape)))
    gJets = gJets.contiguous()
    acc = ops._new_accumulator(input, _lib.LAYOUT_CHANNEL_LAST)
    pb = _problem(input, coords, padding_mode, align_corners, kernel, multicell)
    nbytes = jet_bytes(dim, N, C, P, input[0, 0].numel() if N and C else 0, order)
    with ops._on_device(input.device), ops._timed("JET%dd[bwd]" % dim, nbytes, input.device):
        rc = _lib.load().cs_jet_backward(pb, order, gJets.data_ptr(), coords.data_ptr(), offset.data_ptr(),
                                         acc.data_ptr(), ops._cur_stream(input.device))
    _lib.check(rc, "cs_jet_backward")
    return ops._finish_accumulator(acc, input, _lib.LAYOUT_CHANNEL_LAST)


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
    `autograd.grad(u, x)`, `autograd.grad(u_x, x)` calls of `test_2d.py:55-127` produce.
    Plain torch ops in the jets' own feature-major layout ([features, P], no transposes):
    differentiable once more by autograd, which is all a training step needs."""
    J = jets.shape[0]
    assert J == 1 + order * dim
    val = jets[0]                                   # [F, P]
    d1 = jets[1:1 + dim]                            # [dim, F, P]
    d2 = jets[1 + dim:1 + 2 * dim] if order >= 2 else None
    for layer in head:
        if isinstance(layer, torch.nn.Linear):
            w = layer.weight
            val = w @ val
            if layer.bias is not None:
                val = val + layer.bias[:, None]
            d1 = torch.matmul(w, d1)
            if d2 is not None:
                d2 = torch.matmul(w, d2)
        elif isinstance(layer, torch.nn.Tanh):
            t = torch.tanh(val)
            s1 = 1 - t * t                          # tanh'
            if d2 is not None:
                s2 = -2 * t * s1                    # tanh''
                d2 = s2 * d1 * d1 + s1 * d2
            d1 = s1 * d1
            val = t
        else:
            raise NotImplementedError("jet_mlp supports Linear and Tanh layers, got %s" % type(layer).__name__)
    u = val.t()
    u_a = [d1[a].t() for a in range(dim)]
    u_aa = [d2[a].t() for a in range(dim)] if d2 is not None else None
    return u, u_a, u_aa


__all__ = ["SamplerJet2d", "SamplerJet3d", "jet_forward", "jet_backward", "jet_mlp", "jet_bytes"]
