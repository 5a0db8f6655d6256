"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checker for the batched FP64 LU path: `lu_oracle.c` (the restated arithmetic) and
`lapack_loop.c` (the reference testers' host-LAPACK OpenMP loop, dlopen'ing the OpenBLAS that
scipy bundles). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; nothing under magma_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

MagmaNoTrans, MagmaTrans, MagmaConjTrans = 111, 112, 113


def _compile(src: str, out: str, extra=(), deps=()):
    os.makedirs(_BUILD, exist_ok=True)
    srcp = os.path.join(_HERE, src)
    outp = os.path.join(_BUILD, out)
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in (src, *deps))
    if os.path.exists(outp) and os.path.getmtime(outp) >= newest:
        return outp
    cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
           "-o", outp + ".tmp", srcp, "-lm", *extra]
    subprocess.run(cmd, check=True)
    os.replace(outp + ".tmp", outp)
    return outp


def build():
    """Compile both checker libraries (idempotent). Returns their paths."""
    return (_compile("lu_oracle.c", "liblu_oracle.so"),
            _compile("lapack_loop.c", "liblapack_loop.so", extra=("-ldl",)),
            _compile("lu_oracle_scz.c", "liblu_oracle_scz.so", deps=("lu_oracle_tmpl.h",)))


_lib = None
_lap = None

_i, _l, _d = C.c_int, C.c_long, C.c_double
_pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_pl = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build()[0])
        L.oracle_dlarnv_uniform01.argtypes = [_pi, _l, _pd]
        L.oracle_dgetf2.argtypes = [_i, _i, _pd, _i, _pi]
        L.oracle_dgetf2.restype = _i
        L.oracle_dgetrs.argtypes = [_i, _i, _i, _pd, _i, _pi, _pd, _i]
        L.oracle_dgesv.argtypes = [_i, _i, _pd, _i, _pi, _pd, _i]
        L.oracle_dgesv.restype = _i
        L.oracle_dgetrf_batched.argtypes = [_i, _i, _pd, _i, _l, _pi, _l, _pi, _l]
        L.oracle_dgetrf_nopiv_batched.argtypes = [_i, _i, _pd, _i, _l, _pi, _l]
        L.oracle_dgetrs_batched.argtypes = [_i, _i, _i, _pd, _i, _l, _pi, _l, _pd, _i, _l, _l]
        L.oracle_dgesv_batched.argtypes = [_i, _i, _pd, _i, _l, _pi, _l, _pd, _i, _l, _pi, _l]
        L.oracle_dgetrf_vbatched.argtypes = [_pi, _pi, _pd, _pi, _pl, _pi, _pl, _pi, _l]
        L.oracle_lu_backward_error.argtypes = [_i, _i, _pd, _i, _pd, _i, _pi]
        L.oracle_lu_backward_error.restype = _d
        L.oracle_solve_residual.argtypes = [_i, _i, _i, _pd, _i, _pd, _i, _pd, _i]
        L.oracle_solve_residual.restype = _d
        L.oracle_lu_backward_error_batched.argtypes = [_i, _i, _pd, _i, _l, _pd, _i, _l, _pi, _l, _l]
        L.oracle_lu_backward_error_batched.restype = _d
        L.oracle_solve_residual_batched.argtypes = [_i, _i, _i, _pd, _i, _l, _pd, _i, _l, _pd, _i,
                                                    _l, _l]
        L.oracle_solve_residual_batched.restype = _d
        L.oracle_dgemm_lu_update.argtypes = [_i, _i, _i, _pd, _i, _pd, _i, _pd, _i]
        L.oracle_num_threads.restype = _i
        _lib = L
    return _lib


def find_host_lapack() -> str:
    """Path of the OpenBLAS scipy bundles (the host LAPACK available in this image)."""
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                   "libscipy_openblas*.so"))
    if not cands:
        raise RuntimeError("host LAPACK (scipy's OpenBLAS) not found")
    return os.path.realpath(cands[0])


def lapack():
    global _lap
    if _lap is None:
        L = C.CDLL(build()[1])
        L.lapack_loop_init.argtypes = [C.c_char_p]
        L.lapack_loop_init.restype = _i
        L.lapack_loop_describe.restype = C.c_char_p
        L.lapack_loop_threads.restype = _i
        L.lapack_loop_set_threads.argtypes = [_i]
        L.lapack_loop_set_threads.restype = None
        rc = L.lapack_loop_init(find_host_lapack().encode())
        if rc != 0:
            raise RuntimeError("lapack_loop_init: " + L.lapack_loop_describe().decode())
        L.lapack_dlarnv.argtypes = [_i, _pi, _l, _pd]
        L.lapack_dgetrf_loop.argtypes = [_i, _i, _pd, _i, _l, _pi, _l, _pi, _l]
        L.lapack_dgetrf_loop.restype = _d
        L.lapack_dgetrs_loop.argtypes = [_i, _i, _i, _pd, _i, _l, _pi, _l, _pd, _i, _l, _l]
        L.lapack_dgetrs_loop.restype = _d
        L.lapack_dgesv_loop.argtypes = [_i, _i, _pd, _i, _l, _pi, _l, _pd, _i, _l, _pi, _l]
        L.lapack_dgesv_loop.restype = _d
        L.lapack_dgetrf_vloop.argtypes = [_pi, _pi, _pd, _pi, _pl, _pi, _pl, _pi, _l]
        L.lapack_dgetrf_vloop.restype = _d
        _lap = L
    return _lap


# --------------------------------------------------------------------------------------------
# numpy-level helpers. Batches are stored as arrays of shape (batch, ncols, ld): element
# [b, j, i] is A_b(i, j), i.e. each matrix is column-major with leading dimension ld.
# --------------------------------------------------------------------------------------------

def dlarnv(n: int, iseed=None) -> tuple[np.ndarray, np.ndarray]:
    """The testers' input stream: dlarnv(idist=1, ISEED={0,0,0,1}). Returns (x, next_seed)."""
    seed = np.array([0, 0, 0, 1] if iseed is None else iseed, dtype=np.int32)
    x = np.empty(n, dtype=np.float64)
    lib().oracle_dlarnv_uniform01(seed, n, x)
    return x, seed


def random_batch(batch: int, m: int, n: int, ld: int | None = None, iseed=None):
    """batch matrices m x n drawn back-to-back (lda=m) from the dlarnv stream, then laid out
    with leading dimension ld (padding rows are zero). Returns (A[batch,n,ld], next_seed)."""
    ld = m if ld is None else ld
    x, seed = dlarnv(batch * m * n, iseed)
    A = np.zeros((batch, n, ld), dtype=np.float64)
    A[:, :, :m] = x.reshape(batch, n, m)
    return A, seed


def getrf_batched(A: np.ndarray, m: int):
    """In-place LU of A[batch, n, ld] (m rows used). Returns (ipiv[batch, min(m,n)], info)."""
    batch, n, ld = A.shape
    mn = min(m, n)
    ipiv = np.zeros((batch, max(mn, 1)), dtype=np.int32)
    info = np.zeros(batch, dtype=np.int32)
    lib().oracle_dgetrf_batched(m, n, A.reshape(-1), ld, n * ld, ipiv.reshape(-1), max(mn, 1),
                                info, batch)
    return ipiv[:, :mn], info


def getrs_batched(trans: int, LU: np.ndarray, ipiv: np.ndarray, B: np.ndarray, n: int):
    batch, _, lda = LU.shape
    _, nrhs, ldb = B.shape
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    lib().oracle_dgetrs_batched(trans, n, nrhs, LU.reshape(-1), lda, LU.shape[1] * lda,
                                ip.reshape(-1), ip.shape[1], B.reshape(-1), ldb, nrhs * ldb, batch)


def gesv_batched(A: np.ndarray, B: np.ndarray, n: int):
    batch, _, lda = A.shape
    _, nrhs, ldb = B.shape
    ipiv = np.zeros((batch, n), dtype=np.int32)
    info = np.zeros(batch, dtype=np.int32)
    lib().oracle_dgesv_batched(n, nrhs, A.reshape(-1), lda, A.shape[1] * lda, ipiv.reshape(-1), n,
                               B.reshape(-1), ldb, nrhs * ldb, info, batch)
    return ipiv, info


def getrf_nopiv_batched(A: np.ndarray, m: int) -> np.ndarray:
    """In-place LU without pivoting of A[batch, n, ld] (m rows used); returns info (src/zgetrf_nopiv_batched.cpp:75)."""
    batch, n, ld = A.shape
    info = np.zeros(batch, dtype=np.int32)
    lib().oracle_dgetrf_nopiv_batched(m, n, A.reshape(-1), ld, n * ld, info, batch)
    return info


def getrs_nopiv_batched(trans: int, LU: np.ndarray, B: np.ndarray, n: int):
    """Solve from the no-pivoting factors: oracle_dgetrs with the identity interchanges (src/zgetrs_nopiv_batched.cpp)."""
    batch = LU.shape[0]
    ident = np.ascontiguousarray(np.broadcast_to(np.arange(1, n + 1, dtype=np.int32), (batch, n)))
    getrs_batched(trans, LU, ident, B, n)


def getri_outofplace_batched(LU: np.ndarray, ipiv: np.ndarray, n: int) -> np.ndarray:
    """inv(A) per matrix from the factors (src/zgetri_outofplace_batched.cpp:114-135: identity, unit-lower solve,
    upper solve, column interchanges in reverse). Column j of that product is U^-1 L^-1 e_pi(j), which is the solve
    A X = I with the interchanges applied to the identity first: oracle_dgetrs on the identity, same arithmetic.
    LU is (batch, n, ld) in the stored (column-major) layout of getrf_batched; the result is (batch, n, n)."""
    batch = LU.shape[0]
    eye = np.ascontiguousarray(np.broadcast_to(np.eye(n), (batch, n, n)))
    getrs_batched(111, LU, ipiv, eye, n)  # in place
    return eye


def lu_backward_error(A0: np.ndarray, LU: np.ndarray, ipiv: np.ndarray, m: int) -> float:
    """max over the batch of ||P A0 - L U||_F / (||A0||_F n)."""
    batch, n, ld0 = A0.shape
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    return lib().oracle_lu_backward_error_batched(
        m, n, np.ascontiguousarray(A0).reshape(-1), ld0, n * ld0,
        np.ascontiguousarray(LU).reshape(-1), LU.shape[2], n * LU.shape[2],
        ip.reshape(-1), ip.shape[1], batch)


def solve_residual(trans: int, A0: np.ndarray, X: np.ndarray, B0: np.ndarray, n: int) -> float:
    batch, _, lda = A0.shape
    _, nrhs, ldx = X.shape
    return lib().oracle_solve_residual_batched(
        trans, n, nrhs, np.ascontiguousarray(A0).reshape(-1), lda, A0.shape[1] * lda,
        np.ascontiguousarray(X).reshape(-1), ldx, nrhs * ldx,
        np.ascontiguousarray(B0).reshape(-1), B0.shape[2], nrhs * B0.shape[2], batch)


EPS = float(np.finfo(np.float64).eps) / 2  # LAPACK dlamch('E') = 2^-53, as the testers use
TOL = 30 * EPS                              # testing/magma_util.cpp:192


def flops_getrf(m: float, n: float) -> float:
    """FLOPS_DGETRF, testing/flops.h:79-84,274 (fmuls + fadds)."""
    if m < n:
        mul = 0.5 * m * (m * (n - (1.0 / 3.0) * m - 1.0) + n) + (2.0 / 3.0) * m
        add = 0.5 * m * (m * (n - (1.0 / 3.0) * m) - n) + (1.0 / 6.0) * m
    else:
        mul = 0.5 * n * (n * (m - (1.0 / 3.0) * n - 1.0) + m) + (2.0 / 3.0) * n
        add = 0.5 * n * (n * (m - (1.0 / 3.0) * n) - m) + (1.0 / 6.0) * n
    return mul + add


def flops_getrs(n: float, nrhs: float) -> float:
    """FLOPS_DGETRS, testing/flops.h:90-91,284."""
    return nrhs * n * n + nrhs * n * (n - 1)


# --------------------------------------------------------------------------------------------
# Random butterfly transformation: numpy restatement of magmablas/zgerbt_kernels.cu:21-170 and of the
# level order of magmablas/zgerbt_func_batched.cu:57-210 (z -> d). Every product and sum is rounded separately
# (numpy never contracts into FMA), in the reference's operation order. u, v: arrays of 2n butterfly scalars.
# --------------------------------------------------------------------------------------------
def _rbt_block(A, Ai, Aj, Am, An, u, v):
    """One butterfly level on the Am x An block at (Ai, Aj) of A[batch, col, row] (in place)."""
    r1, c1 = (Am + 1) // 2, (An + 1) // 2
    r2, c2 = Am - r1, An - c1
    z = lambda cols, rows: np.zeros((A.shape[0], cols, rows))  # noqa: E731
    a00 = A[:, Aj:Aj + c1, Ai:Ai + r1].copy()
    a01, a10, a11 = z(c1, r1), z(c1, r1), z(c1, r1)
    a01[:, :c2, :] = A[:, Aj + c1:Aj + c1 + c2, Ai:Ai + r1]
    a10[:, :, :r2] = A[:, Aj:Aj + c1, Ai + r1:Ai + r1 + r2]
    a11[:, :c2, :r2] = A[:, Aj + c1:Aj + c1 + c2, Ai + r1:Ai + r1 + r2]
    u1, u2 = u[:r1], np.concatenate([u[r1:r1 + r2], np.zeros(r1 - r2)])
    v1, v2 = v[:c1], np.concatenate([v[c1:c1 + c2], np.zeros(c1 - c2)])
    b1, b2, b3, b4 = a00 + a01, a10 + a11, a00 - a01, a10 - a11
    uv = lambda uu, vv: (uu[None, None, :] * vv[None, :, None])  # noqa: E731   (u * v) first, as the kernel does
    A[:, Aj:Aj + c1, Ai:Ai + r1] = uv(u1, v1) * (b1 + b2)
    A[:, Aj + c1:Aj + c1 + c2, Ai:Ai + r1] = (uv(u1, v2) * (b3 + b4))[:, :c2, :]
    A[:, Aj:Aj + c1, Ai + r1:Ai + r1 + r2] = (uv(u2, v1) * (b1 - b2))[:, :, :r2]
    A[:, Aj + c1:Aj + c1 + c2, Ai + r1:Ai + r1 + r2] = (uv(u2, v2) * (b3 - b4))[:, :c2, :r2]


def prbt(A: np.ndarray, n: int, u: np.ndarray, v: np.ndarray):
    """A <- U^T A V in place: inner level on the four quadrants (entries [n, 2n)), then the outer level ([0, n))."""
    n1 = (n + 1) // 2
    n2 = n - n1
    ui, vi = u[n:], v[n:]
    _rbt_block(A, 0, 0, n1, n1, ui[0:], vi[0:])
    _rbt_block(A, 0, n1, n1, n2, ui[0:], vi[n1:])
    _rbt_block(A, n1, 0, n2, n1, ui[n1:], vi[0:])
    _rbt_block(A, n1, n1, n2, n2, ui[n1:], vi[n1:])
    _rbt_block(A, 0, 0, n, n, u[0:], v[0:])


def _rbt_vec(B, off, n, u, transpose):
    if n < (2 if transpose else 1):
        return
    n1 = (n + 1) // 2
    n2 = n - n1
    top = B[:, :, off:off + n1].copy()
    bot = np.zeros_like(top)
    bot[:, :, :n2] = B[:, :, off + n1:off + n]
    u0 = u[:n1]
    u1 = np.concatenate([u[n1:n], np.zeros(n1 - n2)])
    if transpose:
        a1, a2 = top + bot, top - bot
        B[:, :, off:off + n1] = u0 * a1
        B[:, :, off + n1:off + n] = (u1 * a2)[:, :, :n2]
    else:
        a1, a2 = u0 * top, u1 * bot
        B[:, :, off:off + n1] = a1 + a2
        B[:, :, off + n1:off + n] = (a1 - a2)[:, :, :n2]


def prbt_mtv(B: np.ndarray, n: int, u: np.ndarray):
    """B <- U^T B in place (B[batch, nrhs, ld]): the two halves, then the outer level."""
    n1 = (n + 1) // 2
    _rbt_vec(B, 0, n1, u[n:], True)
    _rbt_vec(B, n1, n - n1, u[n + n1:], True)
    _rbt_vec(B, 0, n, u, True)


def prbt_mv(B: np.ndarray, n: int, v: np.ndarray):
    """B <- V B in place: the outer level, then the two halves."""
    n1 = (n + 1) // 2
    _rbt_vec(B, 0, n, v, False)
    _rbt_vec(B, 0, n1, v[n:], False)
    _rbt_vec(B, n1, n - n1, v[n + n1:], False)


def gesv_rbt_batched(A: np.ndarray, B: np.ndarray, n: int, u: np.ndarray, v: np.ndarray):
    """In place: butterflies, LU without pivoting, solves, X <- V Y (src/zgesv_rbt_batched.cpp:81-166). Returns info."""
    prbt(A, n, u, v)
    prbt_mtv(B, n, u)
    info = getrf_nopiv_batched(A, n)
    getrs_nopiv_batched(MagmaNoTrans, A, B, n)
    prbt_mv(B, n, v)
    return info


def gemm_lu_update(A: np.ndarray, B: np.ndarray, Cm: np.ndarray):
    """C <- C - A B per element as the chain fma(-a(i,k), b(k,j), c), k increasing (lu_oracle.c). Arrays are single
    matrices in the stored layout [col, row] (column-major), updated in place: A[k, m], B[n, k], C[n, m]."""
    k, m = A.shape
    n = B.shape[0]
    lib().oracle_dgemm_lu_update(m, n, k, A.reshape(-1), m, B.reshape(-1), B.shape[1], Cm.reshape(-1), Cm.shape[1])


# ---- s / c / z: lu_oracle_scz.c ---------------------------------------------------------------
# p is one of "s", "c", "z" ("q" = the same template in double, used only to pin the template against lu_oracle.c).
PREC_DTYPE = {"s": np.float32, "c": np.complex64, "z": np.complex128, "q": np.float64}
_scz = None


def lib_scz():
    global _scz
    if _scz is None:
        L = C.CDLL(build()[2])
        vp = C.c_void_p
        for p in PREC_DTYPE:
            getattr(L, f"oracle_{p}getrf_batched").argtypes = [_i, _i, vp, _i, _l, _pi, _l, _pi, _l]
            getattr(L, f"oracle_{p}getrs_batched").argtypes = [_i, _i, _i, vp, _i, _l, _pi, _l, vp, _i, _l, _l]
            getattr(L, f"oracle_{p}gesv_batched").argtypes = [_i, _i, vp, _i, _l, _pi, _l, vp, _i, _l, _pi, _l]
        _scz = L
    return _scz


def _chk(p, *arrs):
    for a in arrs:
        assert a.dtype == PREC_DTYPE[p] and a.flags.c_contiguous, (a.dtype, p)


def random_batch_prec(p: str, batch: int, m: int, n: int, ld: int | None = None, seed: int = 0) -> np.ndarray:
    """batch matrices m x n, uniform in (0,1) (both parts for c / z), stored (batch, n, ld) column-major like random_batch."""
    ld = m if ld is None else ld
    rng = np.random.default_rng(seed)
    A = np.zeros((batch, n, ld), dtype=PREC_DTYPE[p])
    x = rng.random((batch, n, m))
    if p in ("c", "z"):
        x = x + 1j * rng.random((batch, n, m))
    A[:, :, :m] = x.astype(PREC_DTYPE[p])
    return A


def getrf_batched_prec(p: str, A: np.ndarray, m: int):
    """In-place LU of A[batch, n, ld] in precision p; returns (ipiv, info) like getrf_batched."""
    _chk(p, A)
    batch, n, ld = A.shape
    mn = min(m, n)
    ipiv = np.zeros((batch, max(mn, 1)), dtype=np.int32)
    info = np.zeros(batch, dtype=np.int32)
    getattr(lib_scz(), f"oracle_{p}getrf_batched")(m, n, A.ctypes.data, ld, n * ld, ipiv.reshape(-1), max(mn, 1), info, batch)
    return ipiv[:, :mn], info


def getrs_batched_prec(p: str, trans: int, LU: np.ndarray, ipiv: np.ndarray, B: np.ndarray, n: int):
    _chk(p, LU, B)
    batch, _, lda = LU.shape
    _, nrhs, ldb = B.shape
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    getattr(lib_scz(), f"oracle_{p}getrs_batched")(trans, n, nrhs, LU.ctypes.data, lda, LU.shape[1] * lda, ip.reshape(-1),
                                                   ip.shape[1], B.ctypes.data, ldb, nrhs * ldb, batch)


def gesv_batched_prec(p: str, A: np.ndarray, B: np.ndarray, n: int):
    _chk(p, A, B)
    batch, _, lda = A.shape
    _, nrhs, ldb = B.shape
    ipiv = np.zeros((batch, n), dtype=np.int32)
    info = np.zeros(batch, dtype=np.int32)
    getattr(lib_scz(), f"oracle_{p}gesv_batched")(n, nrhs, A.ctypes.data, lda, A.shape[1] * lda, ipiv.reshape(-1), n,
                                                  B.ctypes.data, ldb, nrhs * ldb, info, batch)
    return ipiv, info
