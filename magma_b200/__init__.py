"""magma_b200 -- B200-native batched FP64 LU (getrf / getrs / gesv / vbatched getrf) behind
MAGMA's own C API. The product is `lib/libmagma_b200.so` (CUDA, sm_100a; see include/magma_b200.h);
this package is the thin host-side mirror used by the tests and the benchmark:

    _lib      ctypes binding of every exported symbol (fails loudly if the library is missing)
    batched   the reference's call signatures over raw device pointers + torch-tensor helpers
    mgpu      one-process-per-GPU sharding of a batch by matrix index (no collective)
    build     nvcc recipe

There is no CPU fallback anywhere in this package.
"""
from . import _lib  # noqa: F401
from .batched import *  # noqa: F401,F403
