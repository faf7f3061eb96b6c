"""Double-precision stage operators (csrc/cs_scalar.cuh): the four entry points of `ops.py` on float64
tensors.  The reference dispatches double (cu2d:905) but cannot run it (its float offset tensor meets
TensorInfo<double>, cu2d:914); here `input`, `grid` and every gradient are float64, `offset` stays the
reference's float32 tensor (modules_2d.py:24-27), and all arithmetic is double.  `ops.forward` etc. route
here on `input.dtype == torch.float64`, so `CosineSampler2d/3d.apply` work on double tensors unchanged.
A correctness path (one thread per (cell, point)); the fp32 engine is the fast one."""
import torch

from . import _lib
from . import ops as _ops


def _check64(t, name, contiguous=True):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float64:
        raise RuntimeError("%s must be float64 like input, got %s" % (name, t.dtype))
    if contiguous and not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _grid64(grid):
    if not grid.is_cuda:
        raise RuntimeError("grid must be a CUDA tensor")
    if grid.dtype != torch.float64:
        raise RuntimeError("grid must be float64 like input, got %s" % grid.dtype)
    if grid.is_contiguous():
        return grid, (grid[0].numel() if grid.shape[0] > 0 else 0)
    if grid.shape[0] > 0 and grid.stride(0) == 0 and grid[0].is_contiguous():
        return grid, 0
    raise RuntimeError("grid must be contiguous")


def _stream64(t, name):
    if not t.is_cuda or t.dtype != torch.float64:
        raise RuntimeError("%s must be a float64 CUDA tensor" % name)
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
    _check64(input, "input")
    grid, grid_sn = _grid64(grid)
    _ops._check(offset, "offset")                       # float32, as the reference builds it
    dim, N, C, D, H, W, P = _ops._geometry(input, grid)
    return grid, grid_sn, dim, N, C, D, H, W, P


def _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn):
    return _ops._problem(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell,
                         _lib.LAYOUT_CHANNEL_FIRST, grid_sn)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def forward(input, grid, offset, padding_mode, align_corners, kernel, multicell, staged=None):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    out = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device)
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _lib.load().cs_forward_f64(pb, input.data_ptr(), grid.data_ptr(), offset.data_ptr(), out.data_ptr(),
                                        _ops._cur_stream(input.device))
    _lib.check(rc, "cs_forward_f64")
    return out


def backward(gOut, input, grid, offset, padding_mode, align_corners, input_requires_grad, kernel, multicell,
             staged=None, want_grid=True):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    gOut, gs = _stream64(gOut, "grad_output")
    gInput = torch.zeros_like(input, memory_format=torch.contiguous_format) if input_requires_grad else None
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _lib.load().cs_backward_f64(pb, gs, input.data_ptr(), grid.data_ptr(), offset.data_ptr(),
                                         _ptr(gInput), _ptr(gGrid), _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward_f64")
    return gInput, gGrid


def backward_backward(gOutInput, gOutGrid, input, grid, gOut, offset, padding_mode, align_corners,
                      input_requires_grad, kernel, multicell, staged=None, want=(True, True, True)):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    _check64(gOutGrid, "grad_out_grid")
    want_input, want_grid, want_ggout = want
    gOut, gs = _stream64(gOut, "grad_output")
    goi = None
    if input_requires_grad and gOutInput is not None and (want_grid or want_ggout):
        _check64(gOutInput, "grad_out_input", contiguous=False)
        goi = gOutInput.contiguous()
    gInput = torch.zeros_like(input, memory_format=torch.contiguous_format) if want_input else None
    gGrid = torch.empty(tuple(grid.shape), dtype=input.dtype, device=input.device) if want_grid else None
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) if want_ggout else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _lib.load().cs_backward_backward_f64(pb, _ptr(goi), gOutGrid.data_ptr(), input.data_ptr(),
                                                  grid.data_ptr(), gs, offset.data_ptr(), _ptr(gInput), _ptr(gGrid),
                                                  _ptr(ggOut), _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward_f64")
    return gInput, gGrid, ggOut


def backward_backward_backward(input, grid, gOut, gOutGrid, gOutgGrid, offset, padding_mode, align_corners,
                               input_requires_grad, kernel, multicell, staged=None, want=(True, True),
                               gOutggOut=None):
    grid, grid_sn, dim, N, C, D, H, W, P = _setup(input, grid, offset)
    _check64(gOutGrid, "gOutGrid")
    _check64(gOutgGrid, "gOutgGrid")
    want_input, want_ggout = want
    gOut, gs = _stream64(gOut, "gOut")
    if gOutggOut is not None and want_input:
        gOutggOut, gs2 = _stream64(gOutggOut, "gOutggOut")
    else:
        gs2 = _ops._null_stream()
    gInput = torch.zeros_like(input, memory_format=torch.contiguous_format) if want_input else None
    ggOut = torch.empty((N, C) + tuple(grid.shape[1:-1]), dtype=input.dtype, device=input.device) if want_ggout else None
    pb = _pb(dim, N, C, D, H, W, P, padding_mode, align_corners, kernel, multicell, grid_sn)
    with _ops._on_device(input.device):
        rc = _lib.load().cs_backward_backward_backward_f64(pb, input.data_ptr(), grid.data_ptr(), gs,
                                                           gOutGrid.data_ptr(), gOutgGrid.data_ptr(), gs2,
                                                           offset.data_ptr(), _ptr(gInput), _ptr(ggOut),
                                                           _ops._cur_stream(input.device))
    _lib.check(rc, "cs_backward_backward_backward_f64")
    return gInput, ggOut
