"""The C-ABI library loads, exports every symbol include/cosine_sampler_b200.h declares,
and the ctypes mirror of its structs matches the C layout.  No compute calls (no GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cosine_sampler_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_four_entry_points():
    names = _declared_functions()
    for n in ("cs_forward", "cs_backward", "cs_backward_backward", "cs_backward_backward_backward",
              "cs_to_channel_last", "cs_from_channel_last", "cs_last_error", "cs_version",
              "cs_launch_count"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert lib.cs_version() == 200
    assert lib.cs_launch_count() == 0 or lib.cs_launch_count() > 0   # callable without a GPU


def test_struct_layout_matches_c(tmp_path):
    from cosinesampler_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "cosine_sampler_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(cs_problem), offsetof(cs_problem, P),
         offsetof(cs_problem, padding_mode), offsetof(cs_problem, field_layout),
         offsetof(cs_problem, grid_stride_n), offsetof(cs_problem, grad_order), sizeof(cs_stream));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P = _lib.Problem
    want = [ctypes.sizeof(P), P.P.offset, P.padding_mode.offset, P.field_layout.offset,
            P.grid_stride_n.offset, P.grad_order.offset, ctypes.sizeof(_lib.Stream3)]
    assert got == want


def test_bad_arguments_are_rejected_without_a_gpu():
    """argument validation happens before any CUDA call"""
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    pb = _lib.Problem()
    pb.dim = 4
    rc = lib.cs_forward(ctypes.byref(pb), None, None, None, None, None)
    assert rc == -1
    assert b"dim" in lib.cs_last_error()
    pb.dim, pb.N, pb.C, pb.D, pb.H, pb.W, pb.P = 2, 1, 4, 1, 8, 8, 16
    rc = lib.cs_forward(ctypes.byref(pb), None, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.cs_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "cs_forward")


def _prototypes():
    """name -> number of parameters, parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|uint64_t|const char \*)\s*(cs_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_ctypes_signatures_have_the_arity_of_the_header():
    """every binding in _lib.py passes exactly as many arguments as the C prototype takes"""
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    protos = _prototypes()
    assert set(protos) == set(_lib.EXPORTS)
    for name, n in protos.items():
        argtypes = getattr(lib, name).argtypes
        assert argtypes is not None and len(argtypes) == n, (name, n, argtypes)


def test_new_entry_points_validate_arguments_without_a_gpu():
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    pb = _lib.Problem()
    pb.dim, pb.N, pb.C, pb.D, pb.H, pb.W, pb.P = 2, 1, 6, 1, 8, 8, 16
    pb.field_layout = _lib.LAYOUT_CHANNEL_LAST
    assert lib.cs_jet_forward(ctypes.byref(pb), 2, None, None, None, None, None) == -2
    assert b"C in" in lib.cs_last_error()
    pb.C = 8
    assert lib.cs_jet_forward(ctypes.byref(pb), 4, None, None, None, None, None) == -1
    assert b"order" in lib.cs_last_error()
    pb.field_layout = _lib.LAYOUT_CHANNEL_FIRST
    assert lib.cs_jet_backward(ctypes.byref(pb), 2, None, None, None, None, None) == -2
    res = _lib.PdeResidual()
    assert lib.cs_pde_head_step(2, 6, 16, None, None, None, None, None, ctypes.byref(res), 1.0,
                                None, None, None, None, None, None, None, None) == -2
    assert lib.cs_pde_head_step(2, 8, 16, None, None, None, None, None, ctypes.byref(res), 1.0,
                                None, None, None, None, None, None, None, None) == -1
    assert lib.cs_peer_allreduce_from_channel_last(9, 0, None, None, 1, 8, 64, None, None, 0, None) == -2
    assert lib.cs_peer_allreduce_from_channel_last(2, 0, None, None, 1, 8, 64, None, None, 0, None) == -1


def test_one_pass_entry_points_validate_arguments_without_a_gpu():
    """cs_bin_* / cs_head_*mix / cs_pde_fused_step reject bad problems before any CUDA call, and the binning
    workspace size is a pure host computation."""
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    pb = _lib.Problem()
    pb.dim, pb.N, pb.C, pb.D, pb.H, pb.W, pb.P = 2, 4, 16, 1, 256, 256, 2 ** 22
    pb.align_corners, pb.multicell = 1, 1
    n = ctypes.c_int64(0)
    assert lib.cs_bin_workspace_bytes(ctypes.byref(pb), ctypes.byref(n)) == 0
    # 65 points per texel: 16 sub-bins per texel -> 2^20 bins, a key and a rank word per point
    assert n.value >= 4 * (2 ** 20 + 2 * 2 ** 22)
    pb.P = 2 ** 31
    assert lib.cs_bin_workspace_bytes(ctypes.byref(pb), ctypes.byref(n)) == -2
    pb.P = 64
    assert lib.cs_bin_points(ctypes.byref(pb), None, None, None, None, None, 0, None) == -1
    assert b"NULL" in lib.cs_last_error()
    assert lib.cs_head_premix(4, 16, 64, 12, None, None, None, None) == -2
    assert lib.cs_head_premix(4, 16, 64, 128, None, None, None, None) == -2
    assert b"hidden width" in lib.cs_last_error()
    assert lib.cs_head_premix(4, 65, 64, 16, None, None, None, None) == -2
    assert lib.cs_head_premix(4, 16, 64, 16, None, None, None, None) == -1
    assert lib.cs_head_postmix(4, 16, 64, 16, None, 0, None, None, None, 0, None, None) == -1
    res = _lib.PdeResidual()
    pb.field_layout = _lib.LAYOUT_CHANNEL_FIRST
    args = (None, None, None, None, None, None, ctypes.byref(res), 1.0, None, None, None, None, None, 1, None)
    assert lib.cs_pde_fused_step(ctypes.byref(pb), *args) == -2            # needs the mixed, channel-last cells
    pb.field_layout = _lib.LAYOUT_CHANNEL_LAST
    pb.C = 12
    assert lib.cs_pde_fused_step(ctypes.byref(pb), *args) == -2 and b"hidden width" in lib.cs_last_error()
    pb.C = 16
    assert lib.cs_pde_fused_step(ctypes.byref(pb), *args) == -1 and b"NULL" in lib.cs_last_error()
    pb.P = 0
    assert lib.cs_pde_fused_step(ctypes.byref(pb), *args) == 0             # empty problem: nothing to launch


def test_scalar_precision_entry_points_validate_arguments_without_a_gpu():
    """cs_*_f64 / cs_*_f16 reject bad problems and NULL tensors before any CUDA call; the half path insists on its
    float workspace when gInput is wanted."""
    from cosinesampler_b200 import _lib
    lib = _lib.load()
    pb = _lib.Problem()
    pb.dim, pb.N, pb.C, pb.D, pb.H, pb.W, pb.P = 2, 2, 3, 1, 8, 8, 16
    pb.align_corners, pb.multicell = 1, 1
    pb.field_layout = _lib.LAYOUT_CHANNEL_FIRST
    one = ctypes.c_void_p(256)                      # a non-NULL pointer that is never dereferenced on these paths
    null_stream = _lib.Stream3(None, 0, 0)
    for sfx in ("_f64", "_f16"):
        fwd = getattr(lib, "cs_forward" + sfx)
        assert fwd(ctypes.byref(pb), None, one, one, None, None) == -1 and b"NULL" in lib.cs_last_error()
        assert fwd(ctypes.byref(pb), None, None, None, None, None) == -1
        pb.field_layout = _lib.LAYOUT_CHANNEL_LAST
        assert fwd(ctypes.byref(pb), one, one, one, one, None) == -2 and b"channel-first" in lib.cs_last_error()
        pb.field_layout = _lib.LAYOUT_CHANNEL_FIRST
        pb.P = 0
        assert fwd(ctypes.byref(pb), None, None, None, None, None) == 0          # empty problem
        pb.P = 16
    assert lib.cs_backward_f64(ctypes.byref(pb), null_stream, one, one, one, one, one, None) == -1
    assert lib.cs_backward_f16(ctypes.byref(pb), null_stream, one, one, one, one, one, None, None) == -1
    gs = _lib.Stream3(256, 0, 0)
    assert lib.cs_backward_f16(ctypes.byref(pb), gs, one, one, one, one, one, None, None) == -1
    assert b"workspace" in lib.cs_last_error()
    assert lib.cs_backward_f16(ctypes.byref(pb), gs, one, one, one, None, None, None, None) == 0     # nothing wanted
