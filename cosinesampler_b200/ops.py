"""Stage-level operators on torch tensors: the host-side mirror of the
reference's native surface (pybind `_cosine_2d` / `_cosine_3d`:
`forward`, `backward`, `backward_backward`, `backward_backward_backward`,
`cosine_sampler_2d.cpp:47-135`, `cosine_sampler_3d.cpp:50-138`), same argument
order and meaning, dispatching to the C ABI in libcosine_sampler_b200.so.

torch is used here for device memory (outputs come from the caching
allocator on the current stream) and nothing else; every kernel is ours.
"""
import os

import torch

from . import _lib

_INDEX_MODE = _lib.INDEX_FUSED if os.environ.get("COSINE_SAMPLER_INDEX_MODE", "").lower() == "fused" \
    else _lib.INDEX_SEPARATE
_LANES = int(os.environ.get("COSINE_SAMPLER_LANES", "0"))


# Optional per-call timing: bench.py sets `profiler` to an object with
# `record(label, algorithmic_bytes, start_event, end_event)`; events are recorded on
# the stream the kernel is launched on.
profiler = None


class _timed:
    """Brackets one stage call with CUDA events when a profiler is installed.
    nbytes: algorithmic bytes (every boundary tensor once, an expanded gOut counted
    as the N-fold tensor the reference materialises); moved: the bytes the launch really has to move (an
    expanded stream once), reported beside it."""
    __slots__ = ("label", "nbytes", "device", "start", "moved")

    def __init__(self, label, nbytes, device, moved=None):
        self.label, self.nbytes, self.device, self.start = label, nbytes, device, None
        self.moved = nbytes if moved is None else moved

    def __enter__(self):
        if profiler is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *a):
        if self.start is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(torch.cuda.current_stream(self.device))
            try:
                profiler.record(self.label, self.nbytes, self.start, end, self.moved)
            except TypeError:                      # a profiler with the 4-argument signature
                profiler.record(self.label, self.nbytes, self.start, end)


def algorithmic_bytes(dim, N, C, P, T, streams, per_point, fields, expanded_streams=0):
    """Algorithmic bytes of one stage call: every tensor crossing the
    operator boundary counted once.  streams = number of [N,C,P] tensors read or written,
    per_point = number of [N,P,dim] tensors (coordinates included), fields = number of
    grid-shaped [N,C,T] tensors read or produced.  expanded_streams of the `streams` are stride-0
    over the cells (PIXEL's `val.sum(0)` gradient): they are counted once instead of N times, which
    gives the bytes a launch really moves."""
    full = streams - expanded_streams
    return 4 * (P * (C * (N * full + expanded_streams) + N * dim * per_point) + fields * N * C * T)


def set_index_mode(mode):
    """'separate' (default; rounds like test/grid_sampler.py:37-38) or 'fused'
    (one fma, like the reference CUDA build with --use_fast_math).  SURVEY 7.1."""
    global _INDEX_MODE
    _INDEX_MODE = {"separate": _lib.INDEX_SEPARATE, "fused": _lib.INDEX_FUSED}[mode]


def get_index_mode():
    return "fused" if _INDEX_MODE == _lib.INDEX_FUSED else "separate"


_SMALL_CELL = {"auto": 0, "never": 1, "always": 2}[os.environ.get("COSINE_SAMPLER_SMALL_CELL", "auto")]


def set_small_cell(mode):
    """'auto' (default): cells whose fields fit in shared memory take the shared-memory kernel
    (gathers from shared memory, privatised scatter); 'never' / 'always' force the choice
    ('always' still falls back when the cell does not fit)."""
    global _SMALL_CELL
    _SMALL_CELL = {"auto": 0, "never": 1, "always": 2}[mode]


_GRAD_ORDER = 0


def set_grad_order(mode):
    """'fast' (default) or 'reference': compute gGrid of the first backward channel by channel in
    the reference's / ATen's operation order (one thread per pair; slower, bit-comparable with
    torch.nn.functional.grid_sample for the linear kernel without multicell)."""
    global _GRAD_ORDER
    _GRAD_ORDER = {"fast": 0, "reference": 1}[mode]


def set_lanes(lanes):
    """0 = automatic; 1/2/4/8 lanes per point quad (1 = whole channel loop in one thread)."""
    global _LANES
    assert lanes in (0, 1, 2, 4, 8)
    _LANES = lanes


# ---------------------------------------------------------------------------
# argument checks (the reference's CHECK_INPUT, cosine_sampler_2d.cpp:4-6)
# ---------------------------------------------------------------------------
def _check(t, name, contiguous=True):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (or float64 / float16 together with input: the scalar paths), got %s"
                           % (name, t.dtype))
    if contiguous and not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _geometry(input, grid):
    dim = grid.shape[-1]
    if dim not in (2, 3) or input.dim() != dim + 2 or grid.dim() != dim + 2:
        raise RuntimeError("expected input [N,C,(D,)H,W] and grid [N,(Do,)Ho,Wo,dim]; got %s and %s"
                           % (tuple(input.shape), tuple(grid.shape)))
    if grid.shape[0] != input.shape[0]:
        raise RuntimeError("grid batch (%d) must equal input batch (%d)" % (grid.shape[0], input.shape[0]))
    N, C = input.shape[0], input.shape[1]
    if dim == 2:
        D, (H, W) = 1, input.shape[2:]
    else:
        D, H, W = input.shape[2:]
    P = 1
    for s in grid.shape[1:-1]:
        P *= s
    return dim, N, C, D, H, W, P


def _grid_view(grid, name="grid"):
    """grid must be contiguous, or expanded along the cell axis (stride 0)."""
    if not grid.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if grid.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    if grid.is_contiguous():
        n_stride = grid[0].numel() if grid.shape[0] > 0 else 0
        return grid, n_stride
    if grid.shape[0] > 0 and grid.stride(0) == 0 and grid[0].is_contiguous():
        return grid, 0
    raise RuntimeError("%s must be contiguous" % name)


def _as_stream(t, P, name):
    """View a [N,C,*spatial] tensor as a strided [N,C,P] stream without copying when
    its point axis is dense (PIXEL hands us gOut expanded over N); else copy."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    ok = True
    expect = 1
    for d in range(t.dim() - 1, 1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            ok = False
            break
        expect *= t.shape[d]
    if not ok or t.stride(0) < 0 or t.stride(1) < 0:
        t = t.contiguous()
    return t, _lib.Stream3(t.data_ptr(), t.stride(0), t.stride(1))


def _null_stream():
    return _lib.Stream3(None, 0, 0)


class Staged:
    """Channel-last copy [N, T, C] of a grid-shaped field, valid for the field as it
    was when staged (identified by storage pointer and version counter)."""
    __slots__ = ("ptr", "version", "shape", "cl")

    def __init__(self, t, cl):
        self.ptr, self.version, self.shape, self.cl = t.data_ptr(), t._version, tuple(t.shape), cl

    def matches(self, t):
        return (t.data_ptr() == self.ptr and t._version == self.version
                and tuple(t.shape) == self.shape)


def _cur_stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """cheap device guard (cosine_sampler_2d.cpp:53): only switches when needed"""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def uses_channel_last(C):
    return C > 0 and C % 4 == 0


def to_channel_last(field):
    """[N,C,*S] -> [N,T,C] copy through cs_to_channel_last."""
    N, C = field.shape[:2]
    T = field[0, 0].numel() if N > 0 and C > 0 else 0
    cl = torch.empty((N, T, C), dtype=field.dtype, device=field.device)
    with _on_device(field.device), _timed("AUX[to_channel_last]", 8 * N * T * C, field.device):
        rc = _lib.load().cs_to_channel_last(field.data_ptr(), cl.data_ptr(), N, C, T,
                                            _cur_stream(field.device))
    _lib.check(rc, "cs_to_channel_last")
    return cl


def from_channel_last(cl, shape, out=None, accumulate=False):
    """[N,T,C] -> [N,C,*S]."""
    N, C = shape[:2]
    T = cl.shape[1]
    if out is None:
        out = torch.empty(shape, dtype=cl.dtype, device=cl.device)
        accumulate = False
    with _on_device(cl.device), _timed("AUX[from_channel_last]", 8 * N * T * C, cl.device):
        rc = _lib.load().cs_from_channel_last(cl.data_ptr(), out.data_ptr(), N, C, T,
                                              1 if accumulate else 0, _cur_stream(cl.device))
    _lib.check(rc, "cs_from_channel_last")
    return out


def stage(field):
    """Stage a field for vector gathers, or None when its channel count rules that out."""
    if not uses_channel_last(field.shape[1]):
        return None
    return Staged(field, to_channel_last(field))


def _field(input, staged):
    """-> (pointer tensor, layout) for the gather side."""
    if staged is not None and staged.matches(input):
        return staged.cl, _lib.LAYOUT_CHANNEL_LAST
    if uses_channel_last(input.shape[1]):
        return to_channel_last(input), _lib.LAYOUT_CHANNEL_LAST
    return input, _lib.LAYOUT_CHANNEL_FIRST


def _problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, layout, grid_sn):
    pb = _lib.Problem()
    pb.dim, pb.N, pb.C, pb.D, pb.H, pb.W, pb.P = dim, N, C, D, H, W, P
    pb.padding_mode = int(padding_mode)
    pb.align_corners = 1 if align_corners else 0
    pb.kernel = int(kernel)
    pb.multicell = 1 if multicell else 0
    pb.index_mode = _INDEX_MODE
    pb.field_layout = layout
    pb.grid_stride_n = grid_sn
    pb.lanes = _LANES
    pb.small_cell = _SMALL_CELL
    pb.grad_order = _GRAD_ORDER
    return pb


def _new_accumulator(input, layout):
    N, C = input.shape[:2]
    with _timed("AUX[zero accumulator]", 4 * input.numel(), input.device):
        if layout == _lib.LAYOUT_CHANNEL_LAST:
            T = input[0, 0].numel() if N > 0 and C > 0 else 0
            return torch.zeros((N, T, C), dtype=input.dtype, device=input.device)
        return torch.zeros_like(input, memory_format=torch.contiguous_format)


def _finish_accumulator(acc, input, layout):
    if layout == _lib.LAYOUT_CHANNEL_LAST:
        return from_channel_last(acc, tuple(input.shape))
    return acc


# ---------------------------------------------------------------------------
# the four entry points
# ---------------------------------------------------------------------------
def forward(input, grid, offset, padding_mode, align_corners, kernel, multicell, staged=None):
    """`_cosine_Xd.forward` (cpp2d:47-62): returns out [N,C,*grid.shape[1:-1]]."""
    if isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16):
        from . import ops_f64
        return ops_f64.forward(input, grid, offset, padding_mode, align_corners, kernel, multicell)
    _check(input, "input")
    grid, grid_sn = _grid_view(grid)
    _check(offset, "offset")
    dim, N, C, D, H, W, P = _geometry(input, grid)
    field, layout = _field(input, staged)
    out = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device)
    pb = _problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, layout, grid_sn)
    nbytes = algorithmic_bytes(dim, N, C, P, D * H * W, streams=1, per_point=1, fields=1)
    with _on_device(input.device), _timed("F%dd" % dim, nbytes, input.device):
        rc = _lib.load().cs_forward(pb, field.data_ptr(), grid.data_ptr(), offset.data_ptr(),
                                    out.data_ptr(), _cur_stream(input.device))
    _lib.check(rc, "cs_forward")
    return out


def backward(gOut, input, grid, offset, padding_mode, align_corners, input_requires_grad, kernel,
             multicell, staged=None, want_grid=True):
    """`_cosine_Xd.backward` (cpp2d:64-85): returns (gInput or None, gGrid).
    want_grid=False additionally elides gGrid (returns None for it)."""
    if isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16):
        from . import ops_f64
        return ops_f64.backward(gOut, input, grid, offset, padding_mode, align_corners, input_requires_grad, kernel,
                                multicell, want_grid=want_grid)
    _check(input, "input")
    grid, grid_sn = _grid_view(grid)
    _check(offset, "offset")
    dim, N, C, D, H, W, P = _geometry(input, grid)
    gOut, gs = _as_stream(gOut, P, "grad_output")
    if want_grid:
        field, layout = _field(input, staged)
    else:
        field = None
        layout = _lib.LAYOUT_CHANNEL_LAST if uses_channel_last(C) else _lib.LAYOUT_CHANNEL_FIRST
    acc = _new_accumulator(input, layout) if input_requires_grad else None
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    pb = _problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, layout, grid_sn)
    label = "B%dd[%s%s]" % (dim, "I" if acc is not None else "", "G" if want_grid else "")
    bkw = dict(streams=1, per_point=1 + (1 if want_grid else 0),
               fields=(1 if want_grid else 0) + (1 if acc is not None else 0))
    nbytes = algorithmic_bytes(dim, N, C, P, D * H * W, **bkw)
    moved = algorithmic_bytes(dim, N, C, P, D * H * W, expanded_streams=1 if gs.stride_n == 0 and N > 1 else 0, **bkw)
    with _on_device(input.device), _timed(label, nbytes, input.device, moved):
        rc = _lib.load().cs_backward(pb, gs, field.data_ptr() if field is not None else None,
                                     grid.data_ptr(), offset.data_ptr(),
                                     acc.data_ptr() if acc is not None else None,
                                     gGrid.data_ptr() if gGrid is not None else None,
                                     _cur_stream(input.device))
    _lib.check(rc, "cs_backward")
    gInput = _finish_accumulator(acc, input, layout) if acc is not None else None
    return gInput, gGrid


def backward_backward(gOutInput, gOutGrid, input, grid, gOut, offset, padding_mode, align_corners,
                      input_requires_grad, kernel, multicell, staged=None, want=(True, True, True)):
    """`_cosine_Xd.backward_backward` (cpp2d:87-106): returns (gInput, gGrid, ggOut).
    gOutInput is read only when input_requires_grad (mod2d:87); `want` elides outputs."""
    if isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16):
        from . import ops_f64
        return ops_f64.backward_backward(gOutInput, gOutGrid, input, grid, gOut, offset, padding_mode, align_corners,
                                         input_requires_grad, kernel, multicell, want=want)
    _check(input, "input")
    grid, grid_sn = _grid_view(grid)
    _check(offset, "offset")
    _check(gOutGrid, "grad_out_grid")
    dim, N, C, D, H, W, P = _geometry(input, grid)
    want_input, want_grid, want_ggout = want
    gOut, gs = _as_stream(gOut, P, "grad_output")
    need_field = want_grid or want_ggout
    if need_field:
        field, layout = _field(input, staged)
    else:
        field = None
        layout = _lib.LAYOUT_CHANNEL_LAST if uses_channel_last(C) else _lib.LAYOUT_CHANNEL_FIRST
    goi = None
    if input_requires_grad and gOutInput is not None and need_field:
        _check(gOutInput, "grad_out_input", contiguous=False)
        goi = gOutInput.contiguous()
        if layout == _lib.LAYOUT_CHANNEL_LAST:
            goi = to_channel_last(goi)
    acc = _new_accumulator(input, layout) if want_input else None
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) \
        if want_ggout else None
    pb = _problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, layout, grid_sn)
    label = "BB%dd[%s%s%s%s]" % (dim, "I" if want_input else "", "G" if want_grid else "",
                                 "O" if want_ggout else "", "+U" if goi is not None else "")
    bkw = dict(streams=1 + (1 if want_ggout else 0), per_point=2 + (1 if want_grid else 0),
               fields=(1 if need_field else 0) + (1 if want_input else 0) + (1 if goi is not None else 0))
    nbytes = algorithmic_bytes(dim, N, C, P, D * H * W, **bkw)
    moved = algorithmic_bytes(dim, N, C, P, D * H * W, expanded_streams=1 if gs.stride_n == 0 and N > 1 else 0, **bkw)
    with _on_device(input.device), _timed(label, nbytes, input.device, moved):
        rc = _lib.load().cs_backward_backward(
            pb, goi.data_ptr() if goi is not None else None, gOutGrid.data_ptr(),
            field.data_ptr() if field is not None else None, grid.data_ptr(), gs, offset.data_ptr(),
            acc.data_ptr() if acc is not None else None,
            gGrid.data_ptr() if gGrid is not None else None,
            ggOut.data_ptr() if ggOut is not None else None, _cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward")
    gInput = _finish_accumulator(acc, input, layout) if acc is not None else None
    return gInput, gGrid, ggOut


def backward_backward_backward(input, grid, gOut, gOutGrid, gOutgGrid, offset, padding_mode,
                               align_corners, input_requires_grad, kernel, multicell, staged=None,
                               want=(True, True), gOutggOut=None):
    """`_cosine_Xd.backward_backward_backward` (cpp2d:108-127): returns (gInput, ggOut).
    input_requires_grad is accepted and ignored, as in the reference kernel (cu2d:736).
    gOutggOut fuses the `b_input` pass of modules_2d.py:109 into the same kernel."""
    if isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16):
        from . import ops_f64
        return ops_f64.backward_backward_backward(input, grid, gOut, gOutGrid, gOutgGrid, offset, padding_mode,
                                                  align_corners, input_requires_grad, kernel, multicell, want=want,
                                                  gOutggOut=gOutggOut)
    _check(input, "input")
    grid, grid_sn = _grid_view(grid)
    _check(offset, "offset")
    _check(gOutGrid, "gOutGrid")
    _check(gOutgGrid, "gOutgGrid")
    dim, N, C, D, H, W, P = _geometry(input, grid)
    want_input, want_ggout = want
    gOut, gs = _as_stream(gOut, P, "gOut")
    if gOutggOut is not None and want_input:
        gOutggOut, gs2 = _as_stream(gOutggOut, P, "gOutggOut")
    else:
        gs2 = _null_stream()
    if want_ggout:
        field, layout = _field(input, staged)
    else:
        field = None
        layout = _lib.LAYOUT_CHANNEL_LAST if uses_channel_last(C) else _lib.LAYOUT_CHANNEL_FIRST
    acc = _new_accumulator(input, layout) if want_input else None
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) \
        if want_ggout else None
    pb = _problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, layout, grid_sn)
    fused = gs2.ptr is not None
    label = "BBB%dd[%s%s%s]" % (dim, "I" if want_input else "", "O" if want_ggout else "",
                                "+X2" if fused else "")
    bkw = dict(streams=1 + (1 if want_ggout else 0) + (1 if fused else 0), per_point=3,
               fields=(1 if want_ggout else 0) + (1 if want_input else 0))
    nbytes = algorithmic_bytes(dim, N, C, P, D * H * W, **bkw)
    nexp = (1 if gs.stride_n == 0 and N > 1 else 0) + (1 if fused and gs2.stride_n == 0 and N > 1 else 0)
    moved = algorithmic_bytes(dim, N, C, P, D * H * W, expanded_streams=nexp, **bkw)
    with _on_device(input.device), _timed(label, nbytes, input.device, moved):
        rc = _lib.load().cs_backward_backward_backward(
            pb, field.data_ptr() if field is not None else None, grid.data_ptr(), gs,
            gOutGrid.data_ptr(), gOutgGrid.data_ptr(), gs2, offset.data_ptr(),
            acc.data_ptr() if acc is not None else None,
            ggOut.data_ptr() if ggOut is not None else None, _cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward_backward")
    gInput = _finish_accumulator(acc, input, layout) if acc is not None else None
    return gInput, ggOut
