"""Build libkmpc.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no torch extension
machinery: the library is a plain C-ABI shared object (include/kmpc.h) loaded with ctypes.
Translation units are compiled in parallel into csrc/_obj/*.o (rebuilt only when stale) and
linked with nvcc -shared."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(PKG_DIR, "libkmpc.so")
SOURCES = ["abi.cu", "stages.cu", "lift.cu", "tc_lift.cu", "edmd.cu", "predict.cu", "closed_loop.cu", "fused.cu"]
NVCC_FLAGS = ["-Xcompiler", "-fPIC", "-O3", "-lineinfo", "-std=c++17",
              "-gencode", "arch=compute_100a,code=sm_100a"]


def _flags():
    """NVCC_FLAGS plus developer extras from KMPC_NVCC_EXTRA (build time only, e.g. -DKMPC_PROFILING)."""
    extra = os.environ.get("KMPC_NVCC_EXTRA", "").split()
    return NVCC_FLAGS + extra


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG_DIR, "..", "include", "kmpc.h"))
    return hs


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    deps = [os.path.join(CSRC, s) for s in SOURCES] + _headers()
    return _newest(deps) > os.path.getmtime(LIB_PATH)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libkmpc.so (skipped when up to date).  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = _newest(_headers())

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if (not force and os.path.exists(o)
                and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_time)):
            return o
        cmd = [nvcc, "-c"] + _flags() + ["-o", o, s]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
