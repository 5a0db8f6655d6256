"""Build libmagma_b200.so in-tree (magma_b200/lib/) with nvcc for sm_100a.

    python -m magma_b200.build [--force] [--verbose]

Objects are rebuilt only when their source (or a header) is newer. nvcc cross-compiles without a
GPU, so this also runs on the CPU-only build box; the resulting .so travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmagma_b200.so")
LIB_INTERPOSE = os.path.join(LIBDIR, "libmagma_b200_interpose.so")

SOURCES = ["runtime.cu", "aux.cu", "lu_small.cu", "lu_small_sq.cu", "lu_mid.cu", "lu_fused.cu", "lu_blocked.cu", "getrs.cu", "getri.cu", "lu_scz.cu", "compat.cu", "rbt.cu", "blas3.cu", "api.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "magma_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile_one(src: str, force: bool, verbose: bool, interpose: bool = False) -> str:
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ + ("_interpose" if interpose else ""), src.replace(".cu", ".o"))
    newest = max(os.path.getmtime(s), _headers_mtime())
    if not force and os.path.exists(o) and os.path.getmtime(o) >= newest:
        return o
    cmd = ["nvcc", *NVCC_FLAGS, *(["-DMB200_INTERPOSE"] if interpose else []), "-c", s, "-o", o]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return o


def build(force: bool = False, verbose: bool = False, interpose: bool = False) -> str:
    """interpose=True builds lib/libmagma_b200_interpose.so (-DMB200_INTERPOSE): only the batched-LU entry points,
    to be loaded on top of a real libmagma whose queues, allocator and xerbla it uses (mode 2)."""
    lib = LIB_INTERPOSE if interpose else LIB
    os.makedirs(OBJ + ("_interpose" if interpose else ""), exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, force, verbose, interpose), SOURCES))
    if force or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        # -Bsymbolic: calls between this library's own entry points (gesv -> getrf -> getrs ...) bind inside it even when
        # another libmagma with the same names was loaded first (interpose mode)
        cmd = ["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib + ".tmp", *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, interpose="--interpose" in sys.argv))
