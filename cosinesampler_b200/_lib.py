"""ctypes binding of libcosine_sampler_b200.so (include/cosine_sampler_b200.h).

There is no fallback: if the shared library is missing this module raises, and
so does every operator built on it."""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# COSINE_SAMPLER_LIB selects an experimental build variant of the same library (tools/)
LIB_PATH = os.environ.get("COSINE_SAMPLER_LIB") or os.path.join(_PKG_DIR, "libcosine_sampler_b200.so")

# every symbol include/cosine_sampler_b200.h declares
EXPORTS = (
    "cs_version", "cs_last_error", "cs_launch_count",
    "cs_forward", "cs_backward", "cs_backward_backward", "cs_backward_backward_backward",
    "cs_to_channel_last", "cs_from_channel_last",
    "cs_forward_f64", "cs_backward_f64", "cs_backward_backward_f64", "cs_backward_backward_backward_f64",
    "cs_forward_f16", "cs_backward_f16", "cs_backward_backward_f16", "cs_backward_backward_backward_f16",
    "cs_jet_forward", "cs_jet_backward", "cs_pde_head_step", "cs_peer_allreduce_from_channel_last",
    "cs_peer_allreduce",
    "cs_bin_workspace_bytes", "cs_bin_points", "cs_head_premix", "cs_head_postmix", "cs_pde_fused_step",
)

PAD_ZEROS, PAD_BORDER, PAD_REFLECTION = 0, 1, 2
KERNEL_COSINE, KERNEL_LINEAR, KERNEL_SMOOTHSTEP = 0, 1, 2
LAYOUT_CHANNEL_FIRST, LAYOUT_CHANNEL_LAST = 0, 1
INDEX_SEPARATE, INDEX_FUSED = 0, 1


class Problem(ctypes.Structure):
    """struct cs_problem"""
    _fields_ = [
        ("dim", ctypes.c_int32), ("N", ctypes.c_int32), ("C", ctypes.c_int32),
        ("D", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("P", ctypes.c_int64),
        ("padding_mode", ctypes.c_int32), ("align_corners", ctypes.c_int32),
        ("kernel", ctypes.c_int32), ("multicell", ctypes.c_int32),
        ("index_mode", ctypes.c_int32), ("field_layout", ctypes.c_int32),
        ("grid_stride_n", ctypes.c_int64),
        ("lanes", ctypes.c_int32), ("small_cell", ctypes.c_int32),
        ("grad_order", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class PdeResidual(ctypes.Structure):
    """struct cs_pde_residual"""
    _fields_ = [("c_u", ctypes.c_float), ("c_u3", ctypes.c_float),
                ("c1", ctypes.c_float * 3), ("c2", ctypes.c_float * 3)]


class Stream3(ctypes.Structure):
    """struct cs_stream: strided [N, C, P] view, P contiguous"""
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_n", ctypes.c_int64), ("stride_c", ctypes.c_int64)]


_lib = None


def load():
    """Load the library once; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "cosinesampler_b200: %s is missing. Build it with "
            "`python -m cosinesampler_b200._build` (needs nvcc); there is no CPU or "
            "PyTorch fallback for this operator." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    pp = ctypes.POINTER(Problem)
    lib.cs_version.restype = ctypes.c_int
    lib.cs_version.argtypes = []
    lib.cs_last_error.restype = ctypes.c_char_p
    lib.cs_last_error.argtypes = []
    lib.cs_launch_count.restype = ctypes.c_uint64
    lib.cs_launch_count.argtypes = []
    lib.cs_forward.restype = ctypes.c_int
    lib.cs_forward.argtypes = [pp, vp, vp, vp, vp, vp]
    lib.cs_backward.restype = ctypes.c_int
    lib.cs_backward.argtypes = [pp, Stream3, vp, vp, vp, vp, vp, vp]
    lib.cs_backward_backward.restype = ctypes.c_int
    lib.cs_backward_backward.argtypes = [pp, vp, vp, vp, vp, Stream3, vp, vp, vp, vp, vp]
    lib.cs_backward_backward_backward.restype = ctypes.c_int
    lib.cs_backward_backward_backward.argtypes = [pp, vp, vp, Stream3, vp, vp, Stream3, vp, vp, vp, vp]
    lib.cs_forward_f64.restype = ctypes.c_int
    lib.cs_forward_f64.argtypes = [pp, vp, vp, vp, vp, vp]
    lib.cs_backward_f64.restype = ctypes.c_int
    lib.cs_backward_f64.argtypes = [pp, Stream3, vp, vp, vp, vp, vp, vp]
    lib.cs_backward_backward_f64.restype = ctypes.c_int
    lib.cs_backward_backward_f64.argtypes = [pp, vp, vp, vp, vp, Stream3, vp, vp, vp, vp, vp]
    lib.cs_backward_backward_backward_f64.restype = ctypes.c_int
    lib.cs_backward_backward_backward_f64.argtypes = [pp, vp, vp, Stream3, vp, vp, Stream3, vp, vp, vp, vp]
    lib.cs_forward_f16.restype = ctypes.c_int
    lib.cs_forward_f16.argtypes = [pp, vp, vp, vp, vp, vp]
    lib.cs_backward_f16.restype = ctypes.c_int
    lib.cs_backward_f16.argtypes = [pp, Stream3, vp, vp, vp, vp, vp, vp, vp]
    lib.cs_backward_backward_f16.restype = ctypes.c_int
    lib.cs_backward_backward_f16.argtypes = [pp, vp, vp, vp, vp, Stream3, vp, vp, vp, vp, vp, vp]
    lib.cs_backward_backward_backward_f16.restype = ctypes.c_int
    lib.cs_backward_backward_backward_f16.argtypes = [pp, vp, vp, Stream3, vp, vp, Stream3, vp, vp, vp, vp, vp]
    lib.cs_jet_forward.restype = ctypes.c_int
    lib.cs_jet_forward.argtypes = [pp, i32, vp, vp, vp, vp, vp]
    lib.cs_jet_backward.restype = ctypes.c_int
    lib.cs_jet_backward.argtypes = [pp, i32, vp, vp, vp, vp, vp]
    lib.cs_pde_head_step.restype = ctypes.c_int
    lib.cs_pde_head_step.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, ctypes.POINTER(PdeResidual),
                                     ctypes.c_float, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.cs_peer_allreduce_from_channel_last.restype = ctypes.c_int
    lib.cs_peer_allreduce_from_channel_last.argtypes = [i32, i32, vp, vp, i32, i32, i64, vp, vp, i32, vp]
    lib.cs_peer_allreduce.restype = ctypes.c_int
    lib.cs_peer_allreduce.argtypes = [i32, i32, vp, vp, i64, vp, vp, vp, vp, i32, vp]
    lib.cs_bin_workspace_bytes.restype = ctypes.c_int
    lib.cs_bin_workspace_bytes.argtypes = [pp, ctypes.POINTER(ctypes.c_int64)]
    lib.cs_bin_points.restype = ctypes.c_int
    lib.cs_bin_points.argtypes = [pp, vp, vp, vp, vp, vp, i64, vp]
    lib.cs_head_premix.restype = ctypes.c_int
    lib.cs_head_premix.argtypes = [i32, i32, i64, i32, vp, vp, vp, vp]
    lib.cs_head_postmix.restype = ctypes.c_int
    lib.cs_head_postmix.argtypes = [i32, i32, i64, i32, vp, i32, vp, vp, vp, i32, vp, vp]
    lib.cs_pde_fused_step.restype = ctypes.c_int
    lib.cs_pde_fused_step.argtypes = [pp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(PdeResidual), ctypes.c_float,
                                      vp, vp, vp, vp, vp, i32, vp]
    lib.cs_to_channel_last.restype = ctypes.c_int
    lib.cs_to_channel_last.argtypes = [vp, vp, i32, i32, i64, vp]
    lib.cs_from_channel_last.restype = ctypes.c_int
    lib.cs_from_channel_last.argtypes = [vp, vp, i32, i32, i64, i32, vp]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cs_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().cs_launch_count())
