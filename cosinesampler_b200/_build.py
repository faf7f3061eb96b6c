"""Build recipe for libcosine_sampler_b200.so (plain nvcc, sm_100a only, in-tree).

The three translation units compile in parallel; the result is a C-ABI shared
library with no libtorch dependency (include/cosine_sampler_b200.h)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libcosine_sampler_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = ["cs_api.cu", "cs_stage_2d.cu", "cs_stage_3d.cu"]
HEADERS = ["cs_engine.cuh", "cs_launch.cuh", os.path.join("..", "..", "include", "cosine_sampler_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build %s" % LIB_NAME)
    return exe


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    logs = {}

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, logs[src]))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s%s" % (r.stdout, r.stderr))
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        for src in SOURCES:
            f.write("==== %s ====\n%s\n" % (src, logs[src]))
    if verbose:
        print("built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
