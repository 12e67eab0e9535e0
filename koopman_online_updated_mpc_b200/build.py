"""Build libkmpc.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no torch extension
machinery: the library is a plain C-ABI shared object (include/kmpc.h) loaded with ctypes."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libkmpc.so")
SOURCES = ["abi.cu", "stages.cu", "lift.cu", "edmd.cu", "closed_loop.cu"]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-O3", "-lineinfo", "-std=c++17",
              "-gencode", "arch=compute_100a,code=sm_100a"]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG_DIR, "..", "include", "kmpc.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libkmpc.so (skipped when up to date).  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
