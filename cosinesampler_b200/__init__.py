"""cosinesampler_b200: B200-native (sm_100a) implementation of the CosineSampler
hot path -- the 2D/3D grid-sample operator with cosine / linear / smoothstep
kernels, multicell offsets and its autograd chain up to the triple backward.

Public surface (the reference's): `CosineSampler2d`, `CosineSampler3d`
(`torch.autograd.Function`s; use `.apply(input, grid, padding_mode, align_corners,
kernel, multicell)`).  Every operator needs the in-tree shared library
`libcosine_sampler_b200.so`; there is no CPU or pure-PyTorch fallback: the first
access to an operator loads the library and raises when it is missing.

The package itself imports lazily (PEP 562): `cosinesampler_b200.chain` / `.dp` are
torch-only scaffolding that the CPU legs of bench.py use without mapping the native
library.
"""
import importlib

_OPERATORS = {
    "CosineSampler2d": ("modules_2d", "CosineSampler2d"),
    "CosineSampler3d": ("modules_3d", "CosineSampler3d"),
    "set_index_mode": ("ops", "set_index_mode"),
    "get_index_mode": ("ops", "get_index_mode"),
    "set_lanes": ("ops", "set_lanes"),
    "set_small_cell": ("ops", "set_small_cell"),
    "set_grad_order": ("ops", "set_grad_order"),
}

__all__ = sorted(_OPERATORS)


def __getattr__(name):
    if name in _OPERATORS:
        from . import _lib
        _lib.load()  # fail loudly when the native library is missing
        mod, attr = _OPERATORS[name]
        value = getattr(importlib.import_module("." + mod, __name__), attr)
        globals()[name] = value
        return value
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
