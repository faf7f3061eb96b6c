"""One-pass fused PIXEL training step (SURVEY section 8f ranks 1 + 2; opt-in, not part of the
reference's API): `csrc/cs_fused.cuh` behind torch tensors.

For a head `Linear(C,K)-Tanh-Linear(K,1)` (`test_2d.py:42-47`, K in {4, 8, 16, 32, 64}) the whole step of
`test_2d.py:36-127` + `loss.backward()` -- replicate the coordinates over the cells, sample, sum over
the cells, head, nested `autograd.grad` for u_a / u_aa, residual, loss, and the triple-backward
scatters of `modules_2d.py:98-111` -- is

    Vh   = W1 . cells                (cs_head_premix,  grid-sized, once per step)
    bin the points by texel          (cs_bin_points,   counting sort, once per chunk of points)
    one pass over the binned points  (cs_pde_fused_step: gather -> tanh / residual / gradients -> scatter)
    cells.grad, W1.grad              (cs_head_postmix, grid-sized, once per step)

The sampler is linear in the cells and so is the head's first layer, so they commute: gathering the
W1-mixed cells yields the hidden pre-activations directly and the three per-point matrix products
of the head become two grid-sized passes.  Same loss, same gradients (fp32, summation order aside)
as `chain.training_step` through the drop-in operator.

    step = OnePassPdeStep(cells, head, residual="helmholtz", kernel="cosine")
    step.begin()
    for xy in chunks:                   # [p, dim] each, any order
        step.add(xy, 1.0 / P_total)
    loss = step.finish()                # cells.grad, head .grad accumulated; 0-dim loss tensor
"""
import ctypes
import math

import torch

from . import _lib, ops
from .autograd import cell_offsets, padding_mode_enum, _kernel_enum, _require_kernel
from .jet import residual_coefficients, _add_grad

HIDDEN_WIDTHS = (4, 8, 16, 32, 64)
MAX_CHANNELS = 64
MAX_CELLS = 32          # the records of all cells of a tile of points live in shared memory


def head_params(head, C):
    """(W1 [K,C], b1 [K], w2 [1,K], b2 [1]) of a Linear(C,K)-Tanh-Linear(K,1) head, K in HIDDEN_WIDTHS."""
    layers = list(head)
    ok = (len(layers) == 3 and isinstance(layers[0], torch.nn.Linear) and isinstance(layers[1], torch.nn.Tanh)
          and isinstance(layers[2], torch.nn.Linear) and layers[0].in_features == C
          and layers[0].out_features in HIDDEN_WIDTHS and layers[2].in_features == layers[0].out_features
          and layers[2].out_features == 1 and layers[0].bias is not None and layers[2].bias is not None
          and C <= MAX_CHANNELS)
    if not ok:
        raise NotImplementedError("the one-pass step needs a Linear(C,K)-Tanh-Linear(K,1) head with biases, "
                                  "K in %s, C <= %d; use jet.jet_autograd_step for other heads"
                                  % (HIDDEN_WIDTHS, MAX_CHANNELS))
    ps = (layers[0].weight, layers[0].bias, layers[2].weight, layers[2].bias)
    for t in ps:
        ops._check(t, "head parameter")
    return ps


def head_is_fusable(head, C):
    try:
        head_params(head, C)
        return True
    except NotImplementedError:
        return False


def small_buffer_size(C, K):
    return K * C + 2 * K + 2


def small_buffer_views(buf, C, K):
    """-> (loss_sum 0-dim, (gW1 [K,C], gb1 [K], gw2 [1,K], gb2 [1])) views of gW1 | gb1 | gw2 | gb2 | loss_sum."""
    kc = K * C
    return buf[kc + 2 * K + 1], (buf[:kc].view(K, C), buf[kc:kc + K], buf[kc + K:kc + 2 * K].view(1, K),
                                 buf[kc + 2 * K:kc + 2 * K + 1])


def _geometry(cells):
    dim = cells.dim() - 2
    if dim not in (2, 3):
        raise RuntimeError("expected cells [N,C,(D,)H,W], got %s" % (tuple(cells.shape),))
    N, C = cells.shape[:2]
    D, H, W = ((1,) + tuple(cells.shape[2:])) if dim == 2 else tuple(cells.shape[2:])
    return dim, N, C, D, H, W


def _problem(cells, P, channels, padding_mode, align_corners, kernel, multicell):
    dim, N, C, D, H, W = _geometry(cells)
    return ops._problem(dim, N, channels, D, H, W, P, padding_mode, align_corners, kernel, multicell,
                        _lib.LAYOUT_CHANNEL_LAST, 0)


def bin_points(cells, coords, offset=None, align_corners=True, multicell=True, want_perm=False):
    """Counting sort of coords [P, dim] on the tile-major texel key of cell 0 (cs_bin_points):
    -> binned coords [P, dim] (and perm [P] int32 with binned[i] = coords[perm[i]] when want_perm).
    The order inside a texel is not deterministic."""
    ops._check(coords, "coords")
    dim = _geometry(cells)[0]
    if coords.dim() != 2 or coords.shape[1] != dim:
        raise RuntimeError("coords must be [P, %d], got %s" % (dim, tuple(coords.shape)))
    P = coords.shape[0]
    pb = _problem(cells, P, cells.shape[1], _lib.PAD_ZEROS, align_corners, _lib.KERNEL_COSINE, multicell)
    out = torch.empty_like(coords)
    perm = torch.empty(P, dtype=torch.int32, device=coords.device) if want_perm else None
    if P == 0:
        return (out, perm) if want_perm else out
    lib = _lib.load()
    nbytes = ctypes.c_int64(0)
    _lib.check(lib.cs_bin_workspace_bytes(pb, ctypes.byref(nbytes)), "cs_bin_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=coords.device)
    with ops._on_device(coords.device), ops._timed("BIN%dd" % dim, 4 * P * (2 * dim + 2), coords.device):
        rc = lib.cs_bin_points(pb, coords.data_ptr(), offset.data_ptr() if offset is not None else None,
                               out.data_ptr(), perm.data_ptr() if perm is not None else None,
                               ws.data_ptr(), nbytes.value, ops._cur_stream(coords.device))
    _lib.check(rc, "cs_bin_points")
    return (out, perm) if want_perm else out


class _BinCache:
    """Binned copies of coordinate tensors that are fed again and again (a PIXEL run keeps its collocation
    points for many steps): keyed on the tensor's storage pointer, version counter and the binning geometry.
    The cache keeps the ORIGINAL tensor alive, so its storage cannot be recycled under the same pointer."""

    def __init__(self, size=2):
        self.size, self.items = size, []

    def get(self, key, coords):
        for k, orig, binned in self.items:
            if k == key and orig._version == key[1]:      # `orig` (or a view of the same storage) is alive
                return binned
        return None

    def put(self, key, coords, binned):
        self.items = [it for it in self.items if it[0] != key][-(self.size - 1):] if self.size > 1 else []
        self.items.append((key, coords, binned))

    def clear(self):
        self.items = []


bin_cache = _BinCache()


def head_premix(cells, W1):
    """Vh [N*T + 1, K]: W1 applied to cells [N, C, *S] texel by texel, plus one zero texel that
    out-of-bounds corners read (cs_head_premix).  `Vh[:N*T].view(N, T, K)` is the mixed stack."""
    ops._check(cells, "input")
    ops._check(W1, "W1")
    N, C = cells.shape[:2]
    K = W1.shape[0]
    T = cells[0, 0].numel() if N and C else 0
    Vh = torch.empty((N * T + 1, K), dtype=cells.dtype, device=cells.device)
    with ops._on_device(cells.device), ops._timed("PREMIX", 4 * N * T * (C + K), cells.device):
        rc = _lib.load().cs_head_premix(N, C, T, K, cells.data_ptr(), W1.data_ptr(), Vh.data_ptr(),
                                        ops._cur_stream(cells.device))
    _lib.check(rc, "cs_head_premix")
    return Vh


def head_postmix(gVh, cells, W1, gW1=None, hidden_first=False, want_input_grad=True):
    """-> gInput [N, C, *S] = W1^T gVh (None when not wanted); gW1 [K, C] += sum_texels gVh (x) cells."""
    ops._check(gVh, "gVh")
    N, C = cells.shape[:2]
    K = W1.shape[0]
    T = cells[0, 0].numel() if N and C else 0
    gInput = torch.empty_like(cells, memory_format=torch.contiguous_format) if want_input_grad else None
    with ops._on_device(cells.device), ops._timed("POSTMIX", 4 * N * T * (2 * C + K), cells.device):
        rc = _lib.load().cs_head_postmix(N, C, T, K, gVh.data_ptr(), 1 if hidden_first else 0, cells.data_ptr(),
                                         W1.data_ptr(), gInput.data_ptr() if gInput is not None else None, 0,
                                         gW1.data_ptr() if gW1 is not None else None,
                                         ops._cur_stream(cells.device))
    _lib.check(rc, "cs_head_postmix")
    return gInput


def fused_bytes(dim, N, K, P, T):
    """Algorithmic bytes of one cs_pde_fused_step launch: the coordinates, Vh read, gVh accumulated."""
    return 4 * (P * dim + 2 * N * K * T)


class OnePassPdeStep:
    """See the module docstring.  aggregate: 'auto' / 'on' (runs of consecutive points with identical corners
    leave as one red per corner) or 'off' (one red per corner and point).  bin: sort every chunk by texel
    (and sub-texel quadrant) first."""

    def __init__(self, cells, head, residual="helmholtz", k2=math.pi ** 2, padding_mode="zeros",
                 align_corners=True, kernel="cosine", multicell=True, bin=True, aggregate="auto", cache_bins=False):
        """cache_bins: remember the binned copy of every coordinate tensor passed to `add` (`fused.bin_cache`,
        the last two tensors) and reuse it while the tensor object, its storage and its version are unchanged."""
        ops._check(cells, "input")
        self.dim = _geometry(cells)[0]
        if cells.shape[0] > MAX_CELLS:
            raise NotImplementedError("the one-pass step holds the records of all cells in shared memory: at most %d "
                                      "cells, got %d (jet.fused_pde_step falls back to the jets path)"
                                      % (MAX_CELLS, cells.shape[0]))
        self.cache_bins = bool(cache_bins)
        if self.dim == 2 and not align_corners:
            raise NotImplementedError(
                "2D with align_corners=False: the reference's 2D forward ignores the flag (cu2d:307-308) while its "
                "backward kernels honour it; the fused step refuses that inconsistent combination")
        self.cells, self.head = cells, head
        self.residual, self.k2 = residual, k2
        self.pm = padding_mode_enum(padding_mode)
        self.kn = _require_kernel(_kernel_enum(kernel, "bilinear" if self.dim == 2 else "trilinear"), kernel)
        self.align_corners, self.multicell = align_corners, multicell
        self.params = head_params(head, cells.shape[1])
        self.K = self.params[0].shape[0]
        self.bin = bool(bin)
        self.aggregate = {"off": 0, "auto": 1, "on": 1}[aggregate]
        self.res = residual_coefficients(residual, self.dim, k2)
        self._live = False

    def begin(self, reducer=None, scale=None):
        """reducer: a `peer.PeerReducer` built for the MIXED cells ([N, K, *S]): the scatters go into its
        symmetric-memory accumulator and `finish` sums over the ranks with one kernel over NVLink."""
        with torch.no_grad():
            self.cells_d = self.cells.detach()
            N, C = self.cells.shape[:2]
            self.T = self.cells_d[0, 0].numel()
            self.offset = cell_offsets(N, self.multicell, self.cells.device)
            W1 = self.params[0].detach()
            self.Vh = head_premix(self.cells_d, W1)
            self.reducer = reducer
            if reducer is not None:
                self.acc = reducer.accumulator()
                self.buf = reducer.small_buffer()
            else:
                # one extra texel absorbs the contributions of out-of-bounds corners
                self.acc = torch.zeros((N * self.T + 1, self.K), dtype=torch.float32, device=self.cells.device)
                self.buf = torch.zeros(small_buffer_size(C, self.K), dtype=torch.float32, device=self.cells.device)
        self.scale = None if scale is None else float(scale)
        self._live = True

    def add(self, xy, scale):
        """One chunk of points xy [p, dim]; its loss contribution is scale * sum_p f^2 (the same `scale`
        for every chunk of a step, e.g. loss_scale / total points)."""
        if not self._live:
            raise RuntimeError("OnePassPdeStep.add before begin()")
        if self.scale is None:
            self.scale = float(scale)
        elif abs(float(scale) - self.scale) > 1e-12 * abs(self.scale):
            raise RuntimeError("OnePassPdeStep.add: every chunk of a step must use the same scale")
        ops._check(xy, "coords")
        if xy.dim() != 2 or xy.shape[1] != self.dim:
            raise RuntimeError("coords must be [p, %d], got %s" % (self.dim, tuple(xy.shape)))
        P = xy.shape[0]
        if P == 0:
            return
        N, C = self.cells.shape[:2]
        dev = self.cells.device
        with torch.no_grad():
            if self.bin:
                key = None
                if self.cache_bins:
                    key = (xy.data_ptr(), xy._version, tuple(xy.shape), str(xy.device), tuple(self.cells.shape[2:]),
                           N, bool(self.align_corners), bool(self.multicell), ops.get_index_mode())
                    hit = bin_cache.get(key, xy)
                binned = hit if key is not None and hit is not None else \
                    bin_points(self.cells_d, xy, self.offset, self.align_corners, self.multicell)
                if key is not None and hit is None:
                    bin_cache.put(key, xy, binned)
                xy = binned
            pb = _problem(self.cells_d, P, self.K, self.pm, self.align_corners, self.kn, self.multicell)
            _, b1, w2, b2 = self.params
            base = self.buf.data_ptr()
            kc = self.K * C
            with ops._on_device(dev), ops._timed("ONEPASS%dd" % self.dim,
                                                  fused_bytes(self.dim, N, self.K, P, self.T), dev):
                rc = _lib.load().cs_pde_fused_step(
                    pb, self.Vh.data_ptr(), xy.data_ptr(), self.offset.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                    b2.data_ptr(), ctypes.byref(self.res), float(scale), self.acc.data_ptr(),
                    base + 4 * kc, base + 4 * (kc + self.K), base + 4 * (kc + 2 * self.K),
                    base + 4 * (kc + 2 * self.K + 1), self.aggregate, ops._cur_stream(dev))
            _lib.check(rc, "cs_pde_fused_step")

    def finish(self):
        if not self._live:
            raise RuntimeError("OnePassPdeStep.finish before begin()")
        self._live = False
        N, C = self.cells.shape[:2]
        W1 = self.params[0].detach()
        with torch.no_grad():
            if self.reducer is not None:
                # sum over the ranks in one kernel over peer memory (in the NVSwitch when the allocations have a
                # multicast address); the loss and the head gradients are those of ALL ranks
                gvh, buf = self.reducer.reduce()
                hidden_first = self.reducer.transposed
            else:
                gvh, buf, hidden_first = self.acc, self.buf, False
            loss_sum, pgrads = small_buffer_views(buf, C, self.K)
            want_w1 = self.params[0].requires_grad
            gcells = head_postmix(gvh, self.cells_d, W1, pgrads[0] if want_w1 else None, hidden_first,
                                  want_input_grad=self.cells.requires_grad)
            if self.cells.requires_grad:
                _add_grad(self.cells, gcells, owned=True)
            for prm, g in zip(self.params, pgrads):
                if prm.requires_grad:
                    _add_grad(prm, g)
            loss = loss_sum * (self.scale if self.scale is not None else 0.0)
        self.acc = self.Vh = self.buf = self.reducer = None
        return loss


def one_pass_pde_step(cells, coords, head, residual="helmholtz", k2=math.pi ** 2, padding_mode="zeros",
                      align_corners=True, kernel="cosine", multicell=True, chunk=None, loss_scale=1.0,
                      reducer=None, bin=True, aggregate="auto", cache_bins=False):
    """`OnePassPdeStep` over coords [P, dim] in chunks of `chunk` points: accumulates `cells.grad` and the
    head parameters' `.grad`, returns loss_scale * mean_p f^2 as a 0-dim tensor."""
    step = OnePassPdeStep(cells, head, residual, k2, padding_mode, align_corners, kernel, multicell, bin, aggregate,
                          cache_bins)
    P = coords.shape[0]
    chunk = max(1, P if not chunk else min(chunk, P))
    step.begin(reducer, scale=loss_scale / P if P else 0.0)
    for s in range(0, P, chunk):
        step.add(coords[s:s + chunk], loss_scale / P)
    return step.finish()


__all__ = ["OnePassPdeStep", "one_pass_pde_step", "bin_points", "head_premix", "head_postmix", "head_is_fusable",
           "head_params", "small_buffer_size", "small_buffer_views", "fused_bytes"]
