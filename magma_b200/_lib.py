"""ctypes binding of libmagma_b200.so. Importing this module never touches oracle/."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB200_LIB") or os.path.join(HERE, "lib", "libmagma_b200.so")  # MB200_LIB: A/B runs of two builds
HEADER = os.path.join(os.path.dirname(HERE), "include", "magma_b200.h")

_lib = None

i32, i64, dbl, vp, sz = C.c_int, C.c_int64, C.c_double, C.c_void_p, C.c_size_t
cstr = C.c_char_p

# name -> (restype, argtypes). Pointers are passed as integers (device or host addresses).
SIGNATURES = {
    "magma_init": (i32, []),
    "magma_finalize": (i32, []),
    "magma_version": (None, [vp, vp, vp]),
    "magma_print_environment": (None, []),
    "magma_num_gpus": (i32, []),
    "magma_getdevices": (None, [vp, i32, vp]),
    "magma_getdevice": (None, [vp]),
    "magma_setdevice": (None, [i32]),
    "magma_getdevice_arch": (i32, []),
    "magma_getdevice_multiprocessor_count": (i32, []),
    "magma_mem_size": (sz, [vp]),
    "magma_queue_create_internal": (None, [i32, vp, cstr, cstr, i32]),
    "magma_queue_create_from_cuda_internal": (None, [i32, vp, vp, vp, vp, cstr, cstr, i32]),
    "magma_queue_destroy_internal": (None, [vp, cstr, cstr, i32]),
    "magma_queue_sync_internal": (None, [vp, cstr, cstr, i32]),
    "magma_queue_get_device": (i32, [vp]),
    "magma_queue_get_cuda_stream": (vp, [vp]),
    "magma_malloc": (i32, [vp, sz]),
    "magma_malloc_cpu": (i32, [vp, sz]),
    "magma_malloc_pinned": (i32, [vp, sz]),
    "magma_free_internal": (i32, [vp, cstr, cstr, i32]),
    "magma_free_cpu": (i32, [vp]),
    "magma_free_pinned_internal": (i32, [vp, cstr, cstr, i32]),
    "magma_memset": (i32, [vp, i32, sz]),
    "magma_memset_async": (i32, [vp, i32, sz, vp]),
    "magma_setvector_internal": (None, [i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_getvector_internal": (None, [i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_setvector_async_internal": (None, [i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_getvector_async_internal": (None, [i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_setmatrix_internal": (None, [i32, i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_getmatrix_internal": (None, [i32, i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_setmatrix_async_internal": (None, [i32, i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_getmatrix_async_internal": (None, [i32, i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_copymatrix_internal": (None, [i32, i32, i32, vp, i32, vp, i32, vp, cstr, cstr, i32]),
    "magma_xerbla": (None, [cstr, i32]),
    "magma_strerror": (cstr, [i32]),
    "magma_wtime": (dbl, []),
    "magma_sync_wtime": (dbl, [vp]),
    "magma_dset_pointer": (None, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "magma_iset_pointer": (None, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "magma_ddisplace_pointers": (None, [vp, vp, i32, i32, i32, i32, vp]),
    "magma_idisplace_pointers": (None, [vp, vp, i32, i32, i32, i32, vp]),
    "magma_doffset_1d": (vp, [vp, i32, i32]),
    "magma_ioffset_1d": (vp, [vp, i32, i32]),
    "magma_doffset_2d": (vp, [vp, i32, i32, i32]),
    "magma_ioffset_2d": (vp, [vp, i32, i32, i32]),
    "magma_dgetrf_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp]),
    "magma_dgetrs_batched": (i32, [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]),
    "magma_dgesv_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_dgetrf_vbatched": (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    "magma_dgetrf_vbatched_max_nocheck_work": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
    "magma_dgetrf_vbatched_max_nocheck": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]),
    "magma_dgetrf_batched_smallsq_noshfl": (i32, [i32, vp, i32, vp, vp, i32, vp]),
    "magma_dgetri_outofplace_batched": (i32, [i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_dgetrf_nopiv_batched": (i32, [i32, i32, vp, i32, vp, i32, vp]),
    "magma_dgetrs_nopiv_batched": (i32, [i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]),
    "magma_dgesv_nopiv_batched": (i32, [i32, i32, vp, i32, vp, i32, vp, i32, vp]),
    "magma_dgesv_batched_small": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_dlaswp_rowserial_batched": (None, [i32, vp, i32, i32, i32, vp, i32, vp]),
    "magmablas_dtrsm_batched": (None, [i32, i32, i32, i32, i32, i32, dbl, vp, i32, vp, i32, i32, vp]),
    "magma_dgemm_batched_core": (None, [i32, i32, i32, i32, i32, dbl, vp, i32, i32, i32, vp, i32, i32, i32, dbl,
                                        vp, i32, i32, i32, i32, vp]),
    "magma_get_dgetrf_batched_nbparam": (None, [i32, vp, vp]),
    "magma_get_dgetrf_vbatched_nbparam": (None, [i32, i32, vp, vp]),
    "magma_get_dgetrf_batched_ntcol": (i32, [i32, i32]),
    "magma_get_dtrsm_batched_stop_nb": (i32, [i32, i32, i32]),
    "magma_b200_dgetrf_batched_mgpu": (i32, [i32, i32, i32, vp, i32, vp, vp, vp, vp]),
    "magma_b200_dgesv_batched_mgpu": (i32, [i32, i32, i32, vp, i32, vp, vp, i32, vp, vp, vp]),
    "magma_b200_dgetrf_batched_host": (i32, [i32, i32, vp, i32, vp, vp, i32, vp]),
    "magma_b200_dgesv_batched_host": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_b200_dlarnv_uniform": (None, [vp, i64, vp, vp]),
    "magma_b200_fp64_peak_tflops": (dbl, [i32, vp]),
    "magma_b200_hbm_copy_gbs": (dbl, [sz, vp]),
    "magma_b200_launch_count": (i64, []),
    "magma_b200_set_tier": (None, [i32]),
    "magma_b200_set_small_rows": (None, [i32]),
    "magma_b200_set_mid_max": (None, [i32]),
    "magma_b200_set_fused_max": (None, [i32]),
    "magma_b200_set_chain_panel": (None, [i32]),
    "magma_b200_set_fused_tail": (None, [i32]),
    "magma_b200_set_tall_panel": (None, [i32]),
    "magma_b200_set_split": (None, [i32]),
    "magma_b200_set_getri_fused": (None, [i32]),
    "magma_sgetrf_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp]),
    "magma_sgetrs_batched": (i32, [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]),
    "magma_sgesv_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_sgetrf_vbatched": (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    "magma_cgetrf_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp]),
    "magma_cgetrs_batched": (i32, [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]),
    "magma_cgesv_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_cgetrf_vbatched": (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    "magma_zgetrf_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp]),
    "magma_zgetrs_batched": (i32, [i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]),
    "magma_zgesv_batched": (i32, [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]),
    "magma_zgetrf_vbatched": (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    "magma_b200_get_dgetrf_batched_crossover": (i32, [i32]),
    "magma_b200_rcp_selftest": (i64, [i64, vp]),
    "magma_dgemm_batched": (None, [i32, i32, i32, i32, i32, dbl, vp, i32, vp, i32, dbl, vp, i32, i32, vp]),
    "magma_dgetrf_batched_strided": (i32, [i32, i32, vp, i32, i32, vp, i32, vp, i32, vp]),
    "magma_dgetrs_batched_strided": (i32, [i32, i32, i32, vp, i32, i32, vp, i32, vp, i32, i32, i32, vp]),
    "magma_dgesv_batched_strided": (i32, [i32, i32, vp, i32, i32, vp, i32, vp, i32, i32, vp, i32, vp]),
    "magma_dgesv_rbt_batched": (i32, [i32, i32, vp, i32, vp, i32, vp, i32, vp]),
    "magma_dgerbt_batched": (i32, [i32, i32, i32, vp, i32, vp, i32, vp, vp, vp, i32, vp]),
    "magmablas_dprbt_batched": (None, [i32, vp, i32, vp, vp, i32, vp]),
    "magmablas_dprbt_mv_batched": (None, [i32, i32, vp, vp, i32, i32, vp]),
    "magmablas_dprbt_mtv_batched": (None, [i32, i32, vp, vp, i32, i32, vp]),
    "magma_dgetf2_fused_batched": (i32, [i32, i32, vp, i32, i32, i32, vp, vp, i32, vp]),
    "magma_dgetf2_batched": (i32, [i32, i32, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp]),
    "magma_dgetrf_recpanel_batched": (i32, [i32, i32, i32, vp, i32, i32, i32, vp, vp, vp, i32, i32, vp]),
    "magma_dlaswp_rowparallel_batched": (None, [i32, vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "magmablas_dtrsv_batched": (None, [i32, i32, i32, i32, vp, i32, vp, i32, i32, vp]),
    "magmaf_dgetrf_batched_": (None, [vp] * 9),
    "magmaf_dgetrs_batched_": (None, [cstr] + [vp] * 10),
    "magmaf_dgesv_batched_": (None, [vp] * 11),
    "magmaf_dgetrf_vbatched_": (None, [vp] * 9),
}


def header_symbols() -> list[str]:
    """Every function name include/magma_b200.h declares (non-static, non-macro)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    txt = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", txt)
    txt = re.sub(r"static\s+inline[^{]*\{[^}]*\}", "", txt)
    names = re.findall(r"\b(magma\w*|magmaf_\w*|magmablas_\w*)\s*\(", txt)
    seen, out = set(), []
    for nme in names:
        if nme not in seen:
            seen.add(nme)
            out.append(nme)
    return out


def load():
    """Load the shared library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m magma_b200.build` "
            "(there is no CPU fallback for the batched LU path)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
