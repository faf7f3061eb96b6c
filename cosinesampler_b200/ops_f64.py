"""Double- and half-precision stage operators (csrc/cs_scalar.cuh): the four entry points of `ops.py` on float64
or float16 tensors.  The reference dispatches double (cu2d:905) but cannot run it (its float offset tensor meets
TensorInfo<double>, cu2d:914); here `input`, `grid` and every gradient are float64, `offset` stays the
reference's float32 tensor (modules_2d.py:24-27), and all arithmetic is double.  `ops.forward` etc. route
here on `input.dtype == torch.float64`, so `CosineSampler2d/3d.apply` work on double tensors unchanged.
float16 (the third type of that dispatch): every tensor, coordinates included, is half; values are widened to
float, the arithmetic is the fp32 formulas, results are rounded to half once, gInput accumulates in a float32
workspace (`cs_*_f16`).
A correctness path (one thread per (cell, point)); the fp32 engine is the fast one."""
import torch

from . import _lib
from . import ops as _ops


DTYPES = {torch.float64: "_f64", torch.float16: "_f16"}


def _check64(t, name, contiguous=True, dtype=torch.float64):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s like input, got %s" % (name, dtype, t.dtype))
    if contiguous and not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _grid64(grid, dtype=torch.float64):
    if not grid.is_cuda:
        raise RuntimeError("grid must be a CUDA tensor")
    if grid.dtype != dtype:
        raise RuntimeError("grid must be %s like input, got %s" % (dtype, grid.dtype))
    if grid.is_contiguous():
        return grid, (grid[0].numel() if grid.shape[0] > 0 else 0)
    if grid.shape[0] > 0 and grid.stride(0) == 0 and grid[0].is_contiguous():
        return grid, 0
    raise RuntimeError("grid must be contiguous")


def _stream64(t, name, dtype=torch.float64):
    if not t.is_cuda or t.dtype != dtype:
        raise RuntimeError("%s must be a %s CUDA tensor" % (name, dtype))
    ok, expect = True, 1
    for d in range(t.dim() - 1, 1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            ok = False
            break
        expect *= t.shape[d]
    if not ok or t.stride(0) < 0 or t.stride(1) < 0:
        t = t.contiguous()
    return t, _lib.Stream3(t.data_ptr(), t.stride(0), t.stride(1))


def _setup(input, grid, offset):
    if input.dtype not in DTYPES:
        raise RuntimeError("input must be float64 or float16 on this path, got %s" % input.dtype)
    _check64(input, "input", dtype=input.dtype)
    grid, grid_sn = _grid64(grid, input.dtype)
    _ops._check(offset, "offset")                       # float32, as the reference builds it
    dim, N, C, D, H, W, P = _ops._geometry(input, grid)
    return grid, grid_sn, dim, N, C, D, H, W, P


def _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn):
    return _ops._problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell,
                         _lib.LAYOUT_CHANNEL_FIRST, grid_sn)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _fn(name, dtype):
    return getattr(_lib.load(), name + DTYPES[dtype])


def _acc(input, wanted):
    """-> (gInput, extra C arguments): float64 accumulates in a zeroed gInput; float16 in a float32 workspace that
    the library zeroes and rounds into gInput."""
    if not wanted:
        return None, ((None,) if input.dtype == torch.float16 else ())
    if input.dtype == torch.float16:
        g = torch.empty_like(input, memory_format=torch.contiguous_format)
        ws = torch.empty(input.numel(), dtype=torch.float32, device=input.device)
        return g, (ws,)
    return torch.zeros_like(input, memory_format=torch.contiguous_format), ()


def forward(input, grid, offset, padding_mode, align_corners, kernel, multicell, staged=None):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    out = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device)
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _fn("cs_forward", input.dtype)(pb, input.data_ptr(), grid.data_ptr(), offset.data_ptr(), out.data_ptr(),
                                            _ops._cur_stream(input.device))
    _lib.check(rc, "cs_forward" + DTYPES[input.dtype])
    return out


def backward(gOut, input, grid, offset, padding_mode, align_corners, input_requires_grad, kernel, multicell,
             staged=None, want_grid=True):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    gOut, gs = _stream64(gOut, "grad_output", input.dtype)
    gInput, ws = _acc(input, input_requires_grad)
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _fn("cs_backward", input.dtype)(pb, gs, input.data_ptr(), grid.data_ptr(), offset.data_ptr(),
                                             _ptr(gInput), _ptr(gGrid), *[_ptr(w) for w in ws],
                                             _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward" + DTYPES[input.dtype])
    return gInput, gGrid


def backward_backward(gOutInput, gOutGrid, input, grid, gOut, offset, padding_mode, align_corners,
                      input_requires_grad, kernel, multicell, staged=None, want=(True, True, True)):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    _check64(gOutGrid, "grad_out_grid", dtype=input.dtype)
    want_input, want_grid, want_ggout = want
    gOut, gs = _stream64(gOut, "grad_output", input.dtype)
    goi = None
    if input_requires_grad and gOutInput is not None and (want_grid or want_ggout):
        _check64(gOutInput, "grad_out_input", contiguous=False, dtype=input.dtype)
        goi = gOutInput.contiguous()
    gInput, ws = _acc(input, want_input)
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) if want_ggout else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _fn("cs_backward_backward", input.dtype)(pb, _ptr(goi), gOutGrid.data_ptr(), input.data_ptr(),
                                                      grid.data_ptr(), gs, offset.data_ptr(), _ptr(gInput),
                                                      _ptr(gGrid), _ptr(ggOut), *[_ptr(w) for w in ws],
                                                      _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward" + DTYPES[input.dtype])
    return gInput, gGrid, ggOut


def backward_backward_backward(input, grid, gOut, gOutGrid, gOutgGrid, offset, padding_mode, align_corners,
                               input_requires_grad, kernel, multicell, staged=None, want=(True, True),
                               gOutggOut=None):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    _check64(gOutGrid, "gOutGrid", dtype=input.dtype)
    _check64(gOutgGrid, "gOutgGrid", dtype=input.dtype)
    want_input, want_ggout = want
    gOut, gs = _stream64(gOut, "gOut", input.dtype)
    if gOutggOut is not None and want_input:
        gOutggOut, gs2 = _stream64(gOutggOut, "gOutggOut", input.dtype)
    else:
        gs2 = _ops._null_stream()
    gInput, ws = _acc(input, want_input)
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) if want_ggout else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _fn("cs_backward_backward_backward", input.dtype)(pb, input.data_ptr(), grid.data_ptr(), gs,
                                                               gOutGrid.data_ptr(), gOutgGrid.data_ptr(), gs2,
                                                               offset.data_ptr(), _ptr(gInput), _ptr(ggOut),
                                                               *[_ptr(w) for w in ws], _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward_backward" + DTYPES[input.dtype])
    return gInput, ggOut
