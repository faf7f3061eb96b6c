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
# (source, object name, extra defines): the stage engine is compiled once per variant
VARIANTS = [(d, 4, l) for d in (2, 3) for l in (0, 1, 2, 3)] + [(2, 1, 0), (3, 1, 0)]
UNITS = [("cs_api.cu", "cs_api.o", [])] + [
    ("cs_stage_inst.cu", "cs_stage_d%d_v%d_l%d.o" % v,
     ["-DCS_DIM=%d" % v[0], "-DCS_VEC=%d" % v[1], "-DCS_LSHIFT=%d" % v[2]]) for v in VARIANTS]
UNITS += [("cs_jet_inst.cu", "cs_jet_d%d_l%d.o" % (d, l), ["-DCS_DIM=%d" % d, "-DCS_LSHIFT=%d" % l])
          for d in (2, 3) for l in (0, 1, 2, 3)]
UNITS += [("cs_head_inst.cu", "cs_head.o", [])]
UNITS += [("cs_fused_inst.cu", "cs_fused_d%d_l%d.o" % (d, l), ["-DCS_DIM=%d" % d, "-DCS_LSHIFT=%d" % l])
          for d in (2, 3) for l in (0, 1, 2, 3, 4)]
SOURCES = ["cs_api.cu", "cs_stage_inst.cu", "cs_jet_inst.cu", "cs_head_inst.cu", "cs_fused_inst.cu"]
HEADERS = ["cs_engine.cuh", "cs_launch.cuh", "cs_jet.cuh", "cs_head.cuh", "cs_head_mma.cuh", "cs_fused.cuh", "cs_scalar.cuh", os.path.join("..", "..", "include", "cosine_sampler_b200.h")]
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


def build(force=False, verbose=False, defines=(), out_path=None):
    """Compile the library if it is missing or older than its sources.
    `defines` / `out_path` build an experimental variant next to the default library."""
    if out_path is None:
        out_path = LIB_PATH
    if not force and not defines and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build" if out_path == LIB_PATH else
                          "build_" + os.path.basename(out_path).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    logs = {}
    extra_defines = tuple(defines)

    def compile_one(unit):
        src, objname, defines = unit
        obj = os.path.join(objdir, objname)
        cmd = [nvcc] + NVCC_FLAGS + list(extra_defines) + defines + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[objname] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (objname, logs[objname]))
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, UNITS))
    cmd = [nvcc, "-shared", "-o", out_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s%s" % (r.stdout, r.stderr))
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        for _, objname, _ in UNITS:
            f.write("==== %s ====\n%s\n" % (objname, logs[objname]))
    if verbose:
        print("built", out_path)
    return out_path


if __name__ == "__main__":
    # python -m cosinesampler_b200._build [--force] [--variant NAME -DFOO=1 ...]
    argv = sys.argv[1:]
    if "--variant" in argv:
        name = argv[argv.index("--variant") + 1]
        defs = [a for a in argv if a.startswith("-D")]
        build(force=True, verbose=True, defines=defs,
              out_path=os.path.join(PKG_DIR, "libcosine_sampler_b200_%s.so" % name))
    else:
        build(force="--force" in argv, verbose=True)
