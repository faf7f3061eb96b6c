"""Boundary proof with the reference's OWN Python layer.

`/root/reference/cosine_sampler_{2,3}d/modules_{2,3}d.py` are imported UNMODIFIED (from where they lie; nothing
is copied) with a stand-in for the pybind extension `_cosine_{2,3}d` that

  * binds every call the reference makes to the signature of `cosinesampler_b200.ops.{forward, backward,
    backward_backward, backward_backward_backward}` -- the host mirror of the C ABI -- so a call that does not
    fit the mirrored pybind argument order (cosine_sampler_2d.cpp:130-135) raises, and
  * computes the result with the stage oracle on the CPU (oracle/stage_oracle.py).

The chain of test/test_2d.py / test/test_3d.py is then driven through the reference's three autograd Functions
and compared with nested autograd over the oracle sampler: the reference's call pattern, flags and return
arities land on our native surface unchanged.  Runs wherever /root/reference exists (the build container);
the GPU box has no reference tree, which is why this proof runs on the CPU."""
import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, grid_sample_3d, make_head
from util import assert_close_scaled, safe_coords

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


def _ops_signatures():
    """inspect.Signature of the four entry points of cosinesampler_b200/ops.py, read from the source without
    importing the module's native library."""
    import ast
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cosinesampler_b200", "ops.py")
    tree = ast.parse(open(path).read())
    sigs = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("forward", "backward", "backward_backward",
                                                               "backward_backward_backward"):
            src = "def %s(%s): pass" % (node.name, ast.unparse(node.args))
            ns = {}
            exec(src, ns)
            sigs[node.name] = inspect.signature(ns[node.name])
    assert len(sigs) == 4
    return sigs


class _StandIn(types.ModuleType):
    """`_cosine_Xd`: same four names as the pybind module; records the calls."""

    def __init__(self, name, sigs):
        super().__init__(name)
        self.sigs, self.calls = sigs, []

    def _bind(self, fn, args):
        b = self.sigs[fn].bind(*args)           # TypeError when the reference's call does not fit ops.<fn>
        self.calls.append(fn)
        return b.arguments

    def forward(self, *args):
        a = self._bind("forward", args)
        assert isinstance(a["padding_mode"], int) and isinstance(a["kernel"], int)
        return so.forward(a["input"], a["grid"], a["offset"], a["padding_mode"], a["align_corners"], a["kernel"],
                          a["multicell"], index_mode=2, compute_dtype=a["input"].dtype)

    def backward(self, *args):
        a = self._bind("backward", args)
        gI, gG = so.backward(a["gOut"], a["input"], a["grid"], a["offset"], a["padding_mode"], a["align_corners"],
                             a["input_requires_grad"], a["kernel"], a["multicell"], index_mode=2, compute_dtype=a["input"].dtype)
        return gI, gG

    def backward_backward(self, *args):
        a = self._bind("backward_backward", args)
        return so.backward_backward(a["gOutInput"], a["gOutGrid"], a["input"], a["grid"], a["gOut"], a["offset"],
                                    a["padding_mode"], a["align_corners"], a["input_requires_grad"], a["kernel"],
                                    a["multicell"], index_mode=2, compute_dtype=a["input"].dtype)

    def backward_backward_backward(self, *args):
        a = self._bind("backward_backward_backward", args)
        return so.backward_backward_backward(a["input"], a["grid"], a["gOut"], a["gOutGrid"], a["gOutgGrid"], a["offset"],
                                             a["padding_mode"], a["align_corners"], a["input_requires_grad"],
                                             a["kernel"], a["multicell"], index_mode=2, compute_dtype=a["input"].dtype)


def _load_reference_module(dim, standin, monkeypatch):
    pkg = "cosine_sampler_%dd" % dim
    fake = types.ModuleType(pkg)
    fake.__path__ = []
    setattr(fake, "_cosine_%dd" % dim, standin)
    monkeypatch.setitem(sys.modules, pkg, fake)
    monkeypatch.setitem(sys.modules, pkg + "._cosine_%dd" % dim, standin)
    spec = importlib.util.spec_from_file_location("reference_modules_%dd" % dim,
                                                  os.path.join(REF, pkg, "modules_%dd.py" % dim))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_python_layer_drives_our_native_surface(dim, monkeypatch):
    # the reference moves its offset tensor `.to('cuda')` (modules_2d.py:25): a no-op on this CPU-only proof
    real_to = torch.Tensor.to

    def to(self, *args, **kw):
        if args and isinstance(args[0], str) and args[0].startswith("cuda"):
            return self
        return real_to(self, *args, **kw)
    monkeypatch.setattr(torch.Tensor, "to", to)
    standin = _StandIn("_cosine_%dd" % dim, _ops_signatures())
    mod = _load_reference_module(dim, standin, monkeypatch)
    S = getattr(mod, "CosineSampler%dd" % dim)

    gen = torch.Generator().manual_seed(5 + dim)
    shape = (4, 4, 12, 12) if dim == 2 else (3, 4, 8, 8, 8)
    cells0 = torch.rand(shape, generator=gen, dtype=torch.float64)
    coords0 = safe_coords(600, dim, shape[2:][::-1], shape[0], True, gen)
    head = make_head(shape[1], seed=2, dtype=torch.float64)
    residual = "t2d" if dim == 2 else "laplace"
    oracle_fn = grid_sample_2d if dim == 2 else grid_sample_3d

    def run(sampler):
        cells = cells0.clone().requires_grad_(True)
        coords = [coords0[:, a:a + 1].clone().requires_grad_(True) for a in range(dim)]
        return derivative_chain(sampler, cells, coords, head, residual=residual)

    # fp64 cells keep the reference's value tests (`(gOutInput != 0.).any().item()`, mod2d:87,104) meaningful
    with torch.autograd.set_detect_anomaly(False):
        got = run(lambda c, g: S.apply(c, g, "zeros", True, "cosine", True))
    want = run(lambda c, g: oracle_fn(c, g, step="cosine", offset=True))
    skip = {"u_xy"}
    for k in want:
        if k in skip:
            continue
        assert_close_scaled(got[k], want[k].reshape(got[k].shape), "%dD %s through the reference's Python layer" % (dim, k),
                            rtol=1e-9, atol_scale=1e-9)
    # the reference's call pattern: 1 forward, first backwards, double backwards, triple backward + BB
    n = {fn: standin.calls.count(fn) for fn in ("forward", "backward", "backward_backward", "backward_backward_backward")}
    assert n["forward"] == 1 and n["backward"] >= dim and n["backward_backward"] >= dim
    assert n["backward_backward_backward"] >= 1
