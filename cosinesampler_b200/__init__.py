"""cosinesampler_b200: B200-native (sm_100a) implementation of the CosineSampler
hot path -- the 2D/3D grid-sample operator with cosine / linear / smoothstep
kernels, multicell offsets and its autograd chain up to the triple backward.

Public surface (the reference's): `CosineSampler2d`, `CosineSampler3d`
(`torch.autograd.Function`s; use `.apply(input, grid, padding_mode, align_corners,
kernel, multicell)`).  Importing this package needs the in-tree shared library
`libcosine_sampler_b200.so`; there is no CPU or pure-PyTorch fallback.
"""
from . import _lib

_lib.load()  # fail loudly, at import time, when the native library is missing

from .modules_2d import CosineSampler2d  # noqa: E402
from .modules_3d import CosineSampler3d  # noqa: E402
from .ops import set_index_mode, get_index_mode, set_lanes, set_small_cell, set_grad_order  # noqa: E402

__all__ = ["CosineSampler2d", "CosineSampler3d", "set_index_mode", "get_index_mode", "set_lanes", "set_small_cell", "set_grad_order"]
