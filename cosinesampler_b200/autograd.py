"""The autograd chain of the reference, re-implemented on top of `ops`.

The reference maps autograd levels 0..3 onto its four native entry points with
three chained `torch.autograd.Function`s per dimensionality
(`cosine_sampler_2d/modules_2d.py:20-111`, `cosine_sampler_3d/modules_3d.py:20-100`):

    CosineSampler{2,3}d            forward -> F      backward -> CosineSamplerBackward
    CosineSamplerBackward          forward -> B      backward -> CosineSamplerBackwardBackward
    CosineSamplerBackwardBackward  forward -> BB     backward -> BBB (+ a BB pass for `b_input`)

Class names, `forward` signatures, defaults and `backward` return arities
(6 / 8 / 10) are the reference's.  What differs, on purpose:

  * no host synchronisation: the reference decides `input_requires_grad` with
    `(gOutInput != 0.).any().item()` (mod2d:87,104; mod3d:41,80,94); here
    undefined incoming gradients arrive as `None` (`set_materialize_grads(False)`)
    and `None` is the test.  The value of every output is unchanged: a zero
    gOutInput contributes zero.
  * dead outputs are not computed: before each stage we ask the autograd
    engine which consumers it will actually run in this pass
    (`torch._C._will_engine_execute_node`), so u_x / u_xx evaluations do not
    pay the scatter into a gInput that nobody reads, and `grad(loss, cells)`
    does not produce coordinate gradients.  `ctx.needs_input_grad` alone cannot
    tell (it is fixed at forward time).
  * `offset` is built once per (N, multicell, device) with the reference's exact
    fp32 `linspace` (mod2d:24-27) instead of a CPU tensor + H2D copy per call,
    and the channel-last staging of `input` made at forward time rides along
    on it to every later stage of the same graph.
  * the expanded (stride-0) `gradOut` PIXEL produces with `val.sum(0)` is read
    in place instead of being materialised N-fold (`.contiguous()`, mod2d:42).
  * under `torch.no_grad` semantics of a plain `.backward()` (grad mode off in
    the triple backward) BBB and its `b_input` BB pass (mod2d:106-109) run as
    one fused kernel.
  * reference crashes are not reproduced: the CPU `torch.zeros(1)` handed to the
    CUDA check (mod2d:89 vs cpp2d:91) and the 7-tuple early-out (mod3d:42).
"""
import threading

import torch

from . import _lib, ops

_lib.load()  # the operators have no CPU / PyTorch fallback: importing them without the library raises


def padding_mode_enum(padding_mode):
    """mod2d:4-10: anything that is not 'zeros' or 'border' means reflection."""
    if padding_mode == "zeros":
        return _lib.PAD_ZEROS
    if padding_mode == "border":
        return _lib.PAD_BORDER
    return _lib.PAD_REFLECTION


def _kernel_enum(kernel, linear_name):
    """mod2d:12-18 / mod3d:12-18: unknown names map to None, which the native
    layer rejects with a TypeError."""
    if kernel == "cosine":
        return _lib.KERNEL_COSINE
    if kernel == linear_name:
        return _lib.KERNEL_LINEAR
    if kernel == "smooth-step":
        return _lib.KERNEL_SMOOTHSTEP
    return None


def _require_kernel(code, kernel):
    if code is None:
        raise TypeError("unsupported interpolation kernel %r" % (kernel,))
    return code


_offset_cache = {}


def cell_offsets(n_cells, multicell, device):
    """offset[n] = linspace(0, 1 - 1/N, N)[n] in fp32, or zeros (mod2d:24-27).
    Evaluated on the CPU exactly as the reference does, moved once, cached."""
    key = (int(n_cells), bool(multicell), str(device))
    t = _offset_cache.get(key)
    if t is None:
        if multicell and n_cells > 0:
            t = torch.linspace(0, 1 - (1 / n_cells), n_cells)
        else:
            t = torch.zeros(n_cells)
        t = t.to(device)
        _offset_cache[key] = t
    return t


# per-thread hint from a `backward` to the `forward` it is about to `apply`:
# which outputs the engine actually needs in this pass
_tls = threading.local()


def _push_want(want):
    _tls.want = want


def _apply_with_want(fn, want, *args):
    """fn.apply(*args) with the per-thread output hint set for exactly that call."""
    _push_want(want)
    try:
        return fn.apply(*args)
    finally:
        _tls.want = None


def _pop_want(default):
    w = getattr(_tls, "want", None)
    _tls.want = None
    return default if w is None else w


# Dead-output elision rests on a private torch API.  It is probed once at import: when it is missing every
# backward computes every output again (correct, but the u_x / u_xx evaluations pay their dead gInput scatters:
# a large slowdown), and that is said once instead of silently.  tests/test_gpu_chain.py pins the launch counts.
_HAVE_EXEC_INFO = hasattr(torch._C, "_will_engine_execute_node")
_warned_exec_info = False


def _warn_no_exec_info(why):
    global _warned_exec_info
    if not _warned_exec_info:
        _warned_exec_info = True
        import warnings
        warnings.warn("cosinesampler_b200: cannot ask the autograd engine which outputs it needs (%s); "
                      "every backward stage computes all its outputs (slower, same results)" % why)


if not _HAVE_EXEC_INFO:
    _warn_no_exec_info("torch._C._will_engine_execute_node is missing in torch %s" % torch.__version__)


def _engine_wants(ctx, idx):
    """Will the engine consume the gradient we return for input `idx` in this pass?
    `idx` indexes the forward's inputs: the tensor arguments come FIRST in every forward signature of this
    module (input, grid, gOut, ...), so `ctx.next_functions[idx]` is the edge of that tensor."""
    if not ctx.needs_input_grad[idx]:
        return False
    if not _HAVE_EXEC_INFO:
        return True
    try:
        fn = ctx.next_functions[idx][0]
    except Exception as exc:
        _warn_no_exec_info("ctx.next_functions: %s" % exc)
        return True
    if fn is None:
        return False
    try:
        return bool(torch._C._will_engine_execute_node(fn))
    except RuntimeError:
        # the documented case: a leaf captured by autograd.grad, or a pass without exec info
        # (plain .backward()): the call raises and the gradient is needed
        return True
    except Exception as exc:
        _warn_no_exec_info("%s: %s" % (type(exc).__name__, exc))
        return True


def _staged_of(offset, input):
    if isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16):
        return None                      # the double-precision path reads the reference layout directly
    ops._check(input, "input")
    st = getattr(offset, "_cs_staged", None)
    if st is not None and st.matches(input):
        return st
    st = ops.stage(input)
    try:
        offset._cs_staged = st
    except Exception:
        pass
    return st


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a + b


def make_functions(dim):
    """Build the three autograd Functions for 2D (dim=2) or 3D (dim=3)."""
    linear_name = "bilinear" if dim == 2 else "trilinear"
    sampler_name = "CosineSampler%dd" % dim

    def kernel_enum(kernel):
        return _kernel_enum(kernel, linear_name)

    class CosineSamplerBackwardBackward(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, grid, gOut, gOutInput, gOutGrid, offset, padding_mode="zeros",
                    align_corners=True, kernel="cosine", multicell=True):
            ctx.align_corners = align_corners
            ctx.padding_mode = padding_mode
            ctx.offset = offset
            ctx.kernel = kernel
            ctx.multicell = multicell
            ctx.set_materialize_grads(False)
            want = _pop_want((True, True, True))
            if gOutGrid is None:
                gOutGrid = torch.zeros(grid.shape, dtype=grid.dtype, device=grid.device)
            gOutGrid = gOutGrid.contiguous()
            input_requires_grad = gOutInput is not None
            gInput, gGrid, ggOut = ops.backward_backward(
                gOutInput, gOutGrid, input, grid, gOut, offset,
                padding_mode_enum(padding_mode), align_corners, input_requires_grad,
                _require_kernel(kernel_enum(kernel), kernel), multicell,
                staged=_staged_of(offset, input), want=want)
            ctx.save_for_backward(input, grid, gOut, gOutGrid)
            return gInput, gGrid, ggOut

        @staticmethod
        def backward(ctx, gOutgInput, gOutgGrid, gOutggOut):
            input, grid, gOut, gOutGrid = ctx.saved_tensors
            want_input = _engine_wants(ctx, 0)
            want_ggout = _engine_wants(ctx, 2)
            pm = padding_mode_enum(ctx.padding_mode)
            kn = _require_kernel(kernel_enum(ctx.kernel), ctx.kernel)
            staged = _staged_of(ctx.offset, input)
            gInput = ggOut = None
            if torch.is_grad_enabled():
                # create_graph=True: keep the reference's graph structure (mod2d:106-111):
                # a raw BBB call plus a differentiable BB node for b_input.
                if gOutgGrid is not None and (want_input or want_ggout):
                    gInput, ggOut = ops.backward_backward_backward(
                        input, grid, gOut, gOutGrid, gOutgGrid.contiguous(), ctx.offset, pm,
                        ctx.align_corners, gOutgInput is not None, kn, ctx.multicell,
                        staged=staged, want=(want_input, want_ggout))
                if gOutggOut is not None and want_input:
                    b_input, _, _ = _apply_with_want(
                        CosineSamplerBackwardBackward, (True, False, False),
                        input, grid, gOutggOut, None, gOutGrid, ctx.offset, ctx.padding_mode,
                        ctx.align_corners, ctx.kernel, ctx.multicell)
                    gInput = _add(gInput, b_input)
            else:
                if gOutgGrid is not None and (want_input or want_ggout):
                    gInput, ggOut = ops.backward_backward_backward(
                        input, grid, gOut, gOutGrid, gOutgGrid.contiguous(), ctx.offset, pm,
                        ctx.align_corners, gOutgInput is not None, kn, ctx.multicell,
                        staged=staged, want=(want_input, want_ggout),
                        gOutggOut=gOutggOut if want_input else None)
                elif gOutggOut is not None and want_input:
                    gInput, _, _ = ops.backward_backward(
                        None, gOutGrid, input, grid, gOutggOut, ctx.offset, pm, ctx.align_corners,
                        False, kn, ctx.multicell, staged=staged, want=(True, False, False))
            return gInput, None, ggOut, None, None, None, None, None, None, None

    class CosineSamplerBackward(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, grid, gOut, offset, padding_mode="zeros", align_corners=True,
                    kernel="cosine", multicell=True):
            ctx.align_corners = align_corners
            ctx.padding_mode = padding_mode
            ctx.offset = offset
            ctx.kernel = kernel
            ctx.multicell = multicell
            ctx.set_materialize_grads(False)
            want_input, want_grid = _pop_want((bool(input.requires_grad), True))
            gInput, gGrid = ops.backward(
                gOut, input, grid, offset, padding_mode_enum(padding_mode), align_corners,
                want_input, _require_kernel(kernel_enum(kernel), kernel), multicell,
                staged=_staged_of(offset, input) if want_grid else None, want_grid=want_grid)
            ctx.save_for_backward(input, grid, gOut)
            return gInput, gGrid

        @staticmethod
        def backward(ctx, gOutInput, gOutGrid):
            input, grid, gOut = ctx.saved_tensors
            if gOutInput is None and gOutGrid is None:
                return None, None, None, None, None, None, None, None
            want = (_engine_wants(ctx, 0), _engine_wants(ctx, 1), _engine_wants(ctx, 2))
            if not any(want):
                return None, None, None, None, None, None, None, None
            if gOutInput is not None:
                gOutInput = gOutInput.contiguous()
            if gOutGrid is not None:
                gOutGrid = gOutGrid.contiguous()
            gInput, gGrid, ggOut = _apply_with_want(
                CosineSamplerBackwardBackward, want,
                input, grid, gOut, gOutInput, gOutGrid, ctx.offset, ctx.padding_mode,
                ctx.align_corners, ctx.kernel, ctx.multicell)
            return gInput, gGrid, ggOut, None, None, None, None, None

    class CosineSampler(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, grid, padding_mode="zeros", align_corners=True, kernel="cosine",
                    multicell=True):
            if not (isinstance(input, torch.Tensor) and input.dtype in (torch.float64, torch.float16)):
                ops._check(input, "input")
                ops._grid_view(grid)
            offset = cell_offsets(input.shape[0], multicell, input.device)
            # a private view object per call: it carries the staging of `input`
            # to every later stage of this graph
            offset = offset.view(-1)
            staged = _staged_of(offset, input)
            ctx.offset = offset
            ctx.padding_mode = padding_mode
            ctx.align_corners = align_corners
            ctx.kernel = kernel
            ctx.multicell = multicell
            ctx.set_materialize_grads(False)
            output = ops.forward(input, grid, offset, padding_mode_enum(padding_mode), align_corners,
                                 _require_kernel(kernel_enum(kernel), kernel), multicell, staged=staged)
            ctx.save_for_backward(input, grid)
            return output

        @staticmethod
        def backward(ctx, gradOut):
            if gradOut is None:
                return None, None, None, None, None, None
            input, grid = ctx.saved_tensors
            want = (_engine_wants(ctx, 0), _engine_wants(ctx, 1))
            if not any(want):
                return None, None, None, None, None, None
            d_input, d_grid = _apply_with_want(
                CosineSamplerBackward, want,
                input, grid, gradOut, ctx.offset, ctx.padding_mode, ctx.align_corners, ctx.kernel,
                ctx.multicell)
            return d_input, d_grid, None, None, None, None

    CosineSampler.__name__ = CosineSampler.__qualname__ = sampler_name
    for cls in (CosineSampler, CosineSamplerBackward, CosineSamplerBackwardBackward):
        cls.__module__ = "cosinesampler_b200.modules_%dd" % dim
    return CosineSampler, CosineSamplerBackward, CosineSamplerBackwardBackward, kernel_enum
