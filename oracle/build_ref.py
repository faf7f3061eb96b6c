"""ORACLE tooling: compile the reference's own CUDA op for sm_100 into oracle/_ref/.

The sources are compiled where they lie under /root/reference (nothing is copied into this
repo); outputs go only to oracle/_ref/ (git-ignored, but shipped to the GPU box):

    oracle/_ref/_cosine_2d.so      pybind module of cosine_sampler_2d/csrc/*.{cpp,cu}
    oracle/_ref/_cosine_3d.so      pybind module of cosine_sampler_3d/csrc/*.{cpp,cu}

The reference's own build system (setup.py / CUDAExtension) is not run; this is the short
recipe it boils down to: nvcc for the .cu, g++ for the .cpp, link against libtorch.  Flags
follow setup.py:33-41 (-O3 --use_fast_math -std=c++17).  Takes ~6 minutes per module (the
AT_DISPATCH macro instantiates 3 dtypes x 2 index types x 4 kernels); the two modules build in
parallel.  On the GPU box the prebuilt files are used as they are; tests skip when absent.

Usage: python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"
MODULES = {
    "_cosine_2d": ("cosine_sampler_2d/csrc/cosine_sampler_2d.cpp",
                   "cosine_sampler_2d/csrc/cosine_sampler_2d_kernel.cu"),
    "_cosine_3d": ("cosine_sampler_3d/csrc/cosine_sampler_3d.cpp",
                   "cosine_sampler_3d/csrc/cosine_sampler_3d_kernel.cu"),
}


def available():
    return all(os.path.exists(os.path.join(OUT, m + ".so")) for m in MODULES)


def _build_one(name, force):
    import torch
    from torch.utils import cpp_extension as ce
    so_path = os.path.join(OUT, name + ".so")
    if os.path.exists(so_path) and not force:
        return so_path
    cpp, cu = (os.path.join(REF, p) for p in MODULES[name])
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    defs = ["-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    obj_cu = os.path.join(OUT, name + "_kernel.o")
    obj_cpp = os.path.join(OUT, name + "_bind.o")
    nvcc = ["nvcc", "-c", cu, "-o", obj_cu, "-O3", "--use_fast_math", "-std=c++17",
            "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC",
            "--expt-relaxed-constexpr", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
            "-D__CUDA_NO_HALF2_OPERATORS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__"] + defs + inc
    gxx = ["g++", "-c", cpp, "-o", obj_cpp, "-O3", "-std=c++17", "-fPIC"] + defs + inc
    for cmd in (gxx, nvcc):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed (%s):\n%s" % (" ".join(cmd[:3]), r.stderr[-4000:]))
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", obj_cpp, obj_cu, "-o", so_path, "-L" + libdir, "-L/usr/local/cuda/lib64",
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
            "-Wl,-rpath," + libdir]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference link failed:\n%s" % r.stderr[-4000:])
    for o in (obj_cu, obj_cpp):
        os.remove(o)
    return so_path


def build(force=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(lambda n: _build_one(n, force), MODULES))
    return True


def load(name):
    """import oracle/_ref/<name>.so (needs `import torch` first); None if not built."""
    import importlib.util
    import torch  # noqa: F401
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("built" if ok else "no /root/reference here; nothing built", OUT)
