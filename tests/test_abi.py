"""CPU tests of the drop-in boundary: the library builds, loads without a GPU, exports every
symbol include/magma_b200.h declares, and reproduces the reference's argument-error behaviour
(return -i after xerbla, quick returns) without touching the device."""
import ctypes as C
import re
import subprocess

import numpy as np

from magma_b200 import _lib, batched


def test_every_declared_symbol_is_exported(lib):
    declared = _lib.header_symbols()
    assert len(declared) > 60
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (\w+)", out.stdout))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in magma_b200.h but not exported: {missing}"
    # and the python binding table covers them all
    unbound = [s for s in declared if s not in _lib.SIGNATURES]
    assert not unbound, f"no ctypes signature for: {unbound}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "magma_b200.h"\nint main(void){ magma_int_t a=0,b=0,c=0; (void)a;(void)b;(void)c; return 0; }\n')
    import os
    inc = os.path.join(os.path.dirname(os.path.dirname(_lib.HERE + "/")), "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)


def test_no_vendor_blas_dependency(lib):
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in out and "cusolver" not in out


def test_sass_is_sm100a_and_uses_fp64_pipe(lib):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    # FP64 pipe, warp-reduce unit (redux.sync -> CREDUX), 128-bit shared-memory broadcast
    assert "DFMA" in sass and "REDUX" in sass and "LDS.128" in sass


def test_version(lib):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib.magma_version(C.addressof(a), C.addressof(b), C.addressof(c))
    assert (a.value, b.value, c.value) == (2, 10, 0)


def test_argument_errors_getrf(lib, capfd):
    f = batched.magma_dgetrf_batched
    assert f(-1, 4, 0, 4, 0, 0, 1, 0) == -1
    assert f(4, -1, 0, 4, 0, 0, 1, 0) == -2
    assert f(4, 4, 0, 3, 0, 0, 1, 0) == -4
    err = capfd.readouterr().err
    assert "On entry to magma_dgetrf_batched, parameter 1 had an illegal value (info = -1)" in err
    assert "parameter 4 had an illegal value" in err
    # quick returns never dereference anything (src/zgetrf_batched.cpp:108-109)
    assert f(0, 4, 0, 1, 0, 0, 10, 0) == 0
    assert f(4, 0, 0, 4, 0, 0, 10, 0) == 0


def test_argument_errors_getrs_gesv(lib, capfd):
    g = batched.magma_dgetrs_batched
    assert g(999, 4, 1, 0, 4, 0, 0, 4, 1, 0) == -1
    assert g(111, -1, 1, 0, 4, 0, 0, 4, 1, 0) == -2
    assert g(111, 4, -1, 0, 4, 0, 0, 4, 1, 0) == -3
    assert g(111, 4, 1, 0, 3, 0, 0, 4, 1, 0) == -5
    assert g(111, 4, 1, 0, 4, 0, 0, 3, 1, 0) == -8
    assert g(111, 0, 1, 0, 1, 0, 0, 1, 1, 0) == 0
    s = batched.magma_dgesv_batched
    assert s(-1, 1, 0, 4, 0, 0, 4, 0, 1, 0) == -1
    assert s(4, -1, 0, 4, 0, 0, 4, 0, 1, 0) == -2
    assert s(4, 1, 0, 3, 0, 0, 4, 0, 1, 0) == -4
    assert s(4, 1, 0, 4, 0, 0, 3, 0, 1, 0) == -6
    assert s(4, 0, 0, 4, 0, 0, 4, 0, 1, 0) == 0
    err = capfd.readouterr().err
    assert "magma_dgetrs_batched, parameter 8" in err and "magma_dgesv_batched, parameter 6" in err


def test_strerror_and_xerbla_wording(lib, capfd):
    assert lib.magma_strerror(0) == b"success"
    assert lib.magma_strerror(-113) == b"cannot allocate memory on GPU device"
    lib.magma_xerbla(b"foo", 113)
    assert "Error in foo, cannot allocate memory on GPU device (info = -113)" in capfd.readouterr().err


def test_offset_helpers_and_tuning(lib):
    base = 0x10000
    assert lib.magma_doffset_1d(base, 2, 3) == base + 8 * 4
    assert lib.magma_doffset_2d(base, 10, 2, 3) == base + 8 * (1 + 2 * 10)
    assert lib.magma_ioffset_2d(base, 10, 2, 3) == base + 4 * (1 + 2 * 10)
    assert batched.magma_get_dgetrf_batched_nbparam(512) == (32, 32)
    assert batched.magma_get_dgetrf_batched_nbparam(1000) == (16, 16)
    assert lib.magma_get_dgetrf_batched_ntcol(16, 16) == 8
    assert lib.magma_get_dtrsm_batched_stop_nb(141, 100, 100) == 32
    # the getters and the drivers read ONE table (csrc/common.cuh): monotone panel widths, tier boundaries
    widths = [batched.magma_get_dgetrf_batched_nbparam(r)[0] for r in (1, 512, 513, 1024, 1025, 2048, 4096, 4097, 8192)]
    assert widths == [32, 32, 16, 16, 8, 8, 4, 2, 2]
    assert all(a >= b for a, b in zip(widths, widths[1:]))
    assert [lib.magma_b200_get_dgetrf_batched_crossover(k) for k in (0, 1, 2)] == [32, 44, 512]
    assert lib.magma_get_dgetrf_batched_ntcol(32, 32) == 4 and lib.magma_get_dgetrf_batched_ntcol(33, 33) == 1
    assert lib.magma_get_dgetrf_batched_ntcol(8, 8) == 16


def test_drop_in_headers_compile(tmp_path):
    """`#include "magma_v2.h"` (and the batched sub-headers) keep working for a caller that switches libraries."""
    import os
    import subprocess
    inc = os.path.join(os.path.dirname(_lib.HERE), "include")
    for hdr in ("magma_v2.h", "magma_batched.h", "magma_dbatched.h"):
        src = tmp_path / (hdr.replace(".", "_") + ".c")
        src.write_text(f'#include "{hdr}"\nint f(void) {{ return (int)MagmaNoTrans + (int)sizeof(magma_queue_t); }}\n')
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_init_fails_loudly_without_gpu(lib, capfd):
    import torch
    if torch.cuda.is_available():
        return
    assert lib.magma_init() == -101  # MAGMA_ERR_NOT_INITIALIZED: no silent CPU path
    assert "no CUDA device" in capfd.readouterr().err


def test_product_does_not_import_oracle():
    import os
    root = os.path.dirname(_lib.HERE + "/")
    for dirpath, _, files in os.walk(_lib.HERE):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "lu_oracle" not in txt.replace(
                    "oracle/lu_oracle.c", ""), f
