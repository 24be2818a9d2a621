"""In-tree nvcc build of libstabletriton_b200.so (sm_100a only) and the standalone selftest binary.

The library is plain C ABI (include/stabletriton_b200.h): no torch, no pybind -- Python reaches it
through ctypes (stabletriton_b200/_cabi.py).  Objects are rebuilt only when a source or header is
newer than the object, so `build()` is cheap to call from tests / bench / smoke.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(CSRC, "libstabletriton_b200.so")
SELFTEST_PATH = os.path.join(CSRC, "selftest")

LIB_SOURCES = ["common.cu", "gemm.cu", "norms.cu", "elementwise.cu", "attention.cu", "peer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libstabletriton_b200.so cannot be built")
    return nvcc


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "stabletriton_b200.h"))
    return hs


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    src_path = os.path.join(CSRC, src)
    if _stale(obj, [src_path] + _headers()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", src_path, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return obj


def build(verbose: bool = False, selftest: bool = True) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns the .so path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(LIB_SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), LIB_SOURCES))
    if _stale(LIB_PATH, objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    if selftest:
        st_src = os.path.join(CSRC, "selftest.cu")
        if _stale(SELFTEST_PATH, [st_src, LIB_PATH] + _headers()):
            cmd = [_nvcc(), *NVCC_FLAGS, st_src, "-o", SELFTEST_PATH, "-L" + CSRC, "-lstabletriton_b200",
                   "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
