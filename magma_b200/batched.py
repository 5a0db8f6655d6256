"""Host-side mirror of the reference's batched-LU interface (include/magma_zbatched.h:463-470,
829-870,1008-1015 and include/magma_zvbatched.h:29-119, z -> d): same names, same argument order
and meaning, same return codes. Every call goes through the C ABI of libmagma_b200.so; torch is
only the owner of device memory and streams.

Array arguments may be torch CUDA tensors (their data_ptr() is passed) or raw integer addresses.
Batches are stored the way the reference testers store them: one contiguous allocation, matrices
back to back, column-major with leading dimension ld. As a tensor that is shape (batch, ncols, ld):
element [b, j, i] is A_b(i, j).
"""
from __future__ import annotations

import ctypes as C

from . import _lib

MagmaNoTrans, MagmaTrans, MagmaConjTrans = 111, 112, 113
MagmaUpper, MagmaLower = 121, 122
MagmaNonUnit, MagmaUnit = 131, 132
MagmaLeft, MagmaRight = 141, 142
MAGMA_SUCCESS = 0
MAGMA_ERR_DEVICE_ALLOC = -113

__all__ = [
    "MagmaNoTrans", "MagmaTrans", "MagmaConjTrans", "MagmaUpper", "MagmaLower", "MagmaNonUnit",
    "MagmaUnit", "MagmaLeft", "MagmaRight", "Queue", "ptr", "magma_init", "magma_finalize",
    "magma_dgetrf_batched", "magma_dgetrs_batched", "magma_dgesv_batched", "magma_dgetrf_vbatched",
    "magma_dgetrf_vbatched_max_nocheck_work", "magma_dgetrf_batched_smallsq_noshfl", "magma_dgetri_outofplace_batched",
    "magma_dgetrf_nopiv_batched", "magma_dgetrs_nopiv_batched", "magma_dgesv_nopiv_batched",
    "magma_dgesv_batched_small", "magma_dset_pointer", "magma_iset_pointer", "magma_ddisplace_pointers",
    "magma_dlaswp_rowserial_batched", "magmablas_dtrsm_batched", "magma_dgemm_batched_core",
    "magma_get_dgetrf_batched_nbparam", "dlarnv_uniform", "set_tier", "set_small_rows", "set_mid_max", "set_fused_max", "launch_count",
    "fp64_peak_tflops", "hbm_copy_gbs", "DeviceBatch", "dgetrf_batched_host", "dgesv_batched_host",
]


def ptr(x) -> int:
    """Address of a tensor / numpy array / raw int (None -> 0)."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


def magma_init() -> int:
    return _lib.load().magma_init()


def magma_finalize() -> int:
    return _lib.load().magma_finalize()


class Queue:
    """magma_queue_t. `Queue(device)` owns a new stream (magma_queue_create); `Queue.from_stream`
    wraps an existing cudaStream_t, e.g. torch's current stream (magma_queue_create_from_cuda)."""

    def __init__(self, device: int = 0, _handle: int | None = None):
        self._L = _lib.load()
        if _handle is None:
            q = C.c_void_p()
            self._L.magma_queue_create_internal(device, C.addressof(q), b"Queue", b"batched.py", 0)
            _handle = q.value
        self.handle = _handle
        self.device = device

    @classmethod
    def from_stream(cls, device: int, cuda_stream: int) -> "Queue":
        L = _lib.load()
        q = C.c_void_p()
        L.magma_queue_create_from_cuda_internal(device, cuda_stream, None, None, C.addressof(q), b"Queue",
                                                b"batched.py", 0)
        return cls(device, q.value)

    @classmethod
    def from_torch(cls, device: int | None = None) -> "Queue":
        import torch
        dev = torch.cuda.current_device() if device is None else device
        return cls.from_stream(dev, torch.cuda.current_stream(dev).cuda_stream)

    @property
    def cuda_stream(self) -> int:
        return self._L.magma_queue_get_cuda_stream(self.handle) or 0

    def sync(self):
        self._L.magma_queue_sync_internal(self.handle, b"sync", b"batched.py", 0)

    def destroy(self):
        if self.handle:
            self._L.magma_queue_destroy_internal(self.handle, b"destroy", b"batched.py", 0)
            self.handle = None


def _q(queue) -> int:
    return queue.handle if isinstance(queue, Queue) else queue


# ---- the reference's entry points ---------------------------------------------------------------

def magma_dgetrf_batched(m, n, dA_array, ldda, ipiv_array, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgetrf_batched(m, n, ptr(dA_array), ldda, ptr(ipiv_array), ptr(info_array),
                                            batchCount, _q(queue))


def magma_dgetrs_batched(trans, n, nrhs, dA_array, ldda, dipiv_array, dB_array, lddb, batchCount, queue) -> int:
    return _lib.load().magma_dgetrs_batched(trans, n, nrhs, ptr(dA_array), ldda, ptr(dipiv_array),
                                            ptr(dB_array), lddb, batchCount, _q(queue))


def magma_dgesv_batched(n, nrhs, dA_array, ldda, dipiv_array, dB_array, lddb, dinfo_array, batchCount,
                        queue) -> int:
    return _lib.load().magma_dgesv_batched(n, nrhs, ptr(dA_array), ldda, ptr(dipiv_array), ptr(dB_array), lddb,
                                           ptr(dinfo_array), batchCount, _q(queue))


def magma_dgetrf_vbatched(m, n, dA_array, ldda, ipiv_array, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgetrf_vbatched(ptr(m), ptr(n), ptr(dA_array), ptr(ldda), ptr(ipiv_array),
                                             ptr(info_array), batchCount, _q(queue))


def magma_dgetrf_vbatched_max_nocheck_work(m, n, max_m, max_n, max_minmn, max_mxn, dA_array, ldda, dipiv_array,
                                           info_array, work, lwork, batchCount, queue) -> int:
    """lwork is a 1-element numpy int32 array (in/out), as `magma_int_t*` in the reference."""
    return _lib.load().magma_dgetrf_vbatched_max_nocheck_work(
        ptr(m), ptr(n), max_m, max_n, max_minmn, max_mxn, ptr(dA_array), ptr(ldda), ptr(dipiv_array),
        ptr(info_array), ptr(work), ptr(lwork), batchCount, _q(queue))


def magma_dgetrf_batched_smallsq_noshfl(n, dA_array, ldda, ipiv_array, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgetrf_batched_smallsq_noshfl(n, ptr(dA_array), ldda, ptr(ipiv_array),
                                                           ptr(info_array), batchCount, _q(queue))


def magma_dgetri_outofplace_batched(n, dA_array, ldda, dipiv_array, dinvA_array, lddia, info_array, batchCount,
                                    queue) -> int:
    """inv(A) from the factors, out of place (src/zgetri_outofplace_batched.cpp:81)."""
    return _lib.load().magma_dgetri_outofplace_batched(n, ptr(dA_array), ldda, ptr(dipiv_array), ptr(dinvA_array),
                                                       lddia, ptr(info_array), batchCount, _q(queue))


def magma_dgetrf_nopiv_batched(m, n, dA_array, ldda, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgetrf_nopiv_batched(m, n, ptr(dA_array), ldda, ptr(info_array), batchCount, _q(queue))


def magma_dgetrs_nopiv_batched(trans, n, nrhs, dA_array, ldda, dB_array, lddb, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgetrs_nopiv_batched(trans, n, nrhs, ptr(dA_array), ldda, ptr(dB_array), lddb,
                                                  ptr(info_array), batchCount, _q(queue))


def magma_dgesv_nopiv_batched(n, nrhs, dA_array, ldda, dB_array, lddb, info_array, batchCount, queue) -> int:
    return _lib.load().magma_dgesv_nopiv_batched(n, nrhs, ptr(dA_array), ldda, ptr(dB_array), lddb, ptr(info_array),
                                                 batchCount, _q(queue))


def magma_dgesv_batched_small(n, nrhs, dA_array, ldda, dipiv_array, dB_array, lddb, dinfo_array, batchCount,
                              queue) -> int:
    return _lib.load().magma_dgesv_batched_small(n, nrhs, ptr(dA_array), ldda, ptr(dipiv_array), ptr(dB_array),
                                                 lddb, ptr(dinfo_array), batchCount, _q(queue))


def magma_dset_pointer(output_array, input, lda, row, column, batch_offset, batchCount, queue):
    _lib.load().magma_dset_pointer(ptr(output_array), ptr(input), lda, row, column, batch_offset, batchCount,
                                   _q(queue))


def magma_iset_pointer(output_array, input, lda, row, column, batchSize, batchCount, queue):
    _lib.load().magma_iset_pointer(ptr(output_array), ptr(input), lda, row, column, batchSize, batchCount,
                                   _q(queue))


def magma_ddisplace_pointers(output_array, input_array, lda, row, column, batchCount, queue):
    _lib.load().magma_ddisplace_pointers(ptr(output_array), ptr(input_array), lda, row, column, batchCount,
                                         _q(queue))


def magma_dlaswp_rowserial_batched(n, dA_array, lda, k1, k2, ipiv_array, batchCount, queue):
    _lib.load().magma_dlaswp_rowserial_batched(n, ptr(dA_array), lda, k1, k2, ptr(ipiv_array), batchCount,
                                               _q(queue))


def magmablas_dtrsm_batched(side, uplo, transA, diag, m, n, alpha, dA_array, ldda, dB_array, lddb, batchCount,
                            queue):
    _lib.load().magmablas_dtrsm_batched(side, uplo, transA, diag, m, n, alpha, ptr(dA_array), ldda,
                                        ptr(dB_array), lddb, batchCount, _q(queue))


def magma_dgemm_batched_core(transA, transB, m, n, k, alpha, dA_array, Ai, Aj, ldda, dB_array, Bi, Bj, lddb,
                             beta, dC_array, Ci, Cj, lddc, batchCount, queue):
    _lib.load().magma_dgemm_batched_core(transA, transB, m, n, k, alpha, ptr(dA_array), Ai, Aj, ldda,
                                         ptr(dB_array), Bi, Bj, lddb, beta, ptr(dC_array), Ci, Cj, lddc,
                                         batchCount, _q(queue))


def magma_get_dgetrf_batched_nbparam(n: int) -> tuple[int, int]:
    nb, recnb = C.c_int(), C.c_int()
    _lib.load().magma_get_dgetrf_batched_nbparam(n, C.addressof(nb), C.addressof(recnb))
    return nb.value, recnb.value


# ---- additions ------------------------------------------------------------------------------------

def dlarnv_uniform(iseed, n: int, dx, queue):
    """Fill dx[0:n] on the device with LAPACK's dlarnv(1, iseed) stream; iseed (numpy int32[4]) is
    advanced in place."""
    _lib.load().magma_b200_dlarnv_uniform(ptr(iseed), n, ptr(dx), _q(queue))


def set_tier(tier: int):
    _lib.load().magma_b200_set_tier(tier)


def set_small_rows(rows: int):
    _lib.load().magma_b200_set_small_rows(rows)


def rcp_selftest(n: int, queue) -> int:
    return _lib.load().magma_b200_rcp_selftest(n, _q(queue))


def set_chain_panel(on: int):
    _lib.load().magma_b200_set_chain_panel(on)


def set_split(parts: int):
    _lib.load().magma_b200_set_split(parts)


def set_tall_panel(on: int):
    _lib.load().magma_b200_set_tall_panel(on)


def set_fused_tail(level: int):
    _lib.load().magma_b200_set_fused_tail(level)


def set_getri_fused(on: int):
    _lib.load().magma_b200_set_getri_fused(on)


def set_fused_max(n: int):
    _lib.load().magma_b200_set_fused_max(n)


def set_mid_max(n: int):
    _lib.load().magma_b200_set_mid_max(n)


def launch_count() -> int:
    return _lib.load().magma_b200_launch_count()


def fp64_peak_tflops(kind: int, queue) -> float:
    return _lib.load().magma_b200_fp64_peak_tflops(kind, _q(queue))


def hbm_copy_gbs(nbytes: int, queue) -> float:
    return _lib.load().magma_b200_hbm_copy_gbs(nbytes, _q(queue))


def dgetrf_batched_host(m, n, hA, lda, hipiv, hinfo, batchCount, queue) -> int:
    return _lib.load().magma_b200_dgetrf_batched_host(m, n, ptr(hA), lda, ptr(hipiv), ptr(hinfo), batchCount,
                                                      _q(queue))


def dgesv_batched_host(n, nrhs, hA, lda, hipiv, hB, ldb, hinfo, batchCount, queue) -> int:
    return _lib.load().magma_b200_dgesv_batched_host(n, nrhs, ptr(hA), lda, ptr(hipiv), ptr(hB), ldb,
                                                     ptr(hinfo), batchCount, _q(queue))


class DeviceBatch:
    """Device storage for a fixed-size batch laid out like the reference testers do
    (testing/testing_zgetrf_batched.cpp:161-202): A[batch, n, ldda], ipiv[batch, min(m,n)],
    info[batch], plus the device pointer arrays built with magma_dset_pointer / magma_iset_pointer."""

    def __init__(self, batch: int, m: int, n: int, ldda: int | None = None, nrhs: int = 0,
                 lddb: int | None = None, device: int = 0, queue: Queue | None = None):
        import torch
        self.torch = torch
        self.batch, self.m, self.n, self.nrhs = batch, m, n, nrhs
        self.ldda = m if ldda is None else ldda
        self.lddb = (max(m, n) if lddb is None else lddb) if nrhs else 0
        self.mn = min(m, n)
        dev = torch.device("cuda", device)
        self.queue = queue or Queue.from_torch(device)
        self.A = torch.zeros((batch, n, self.ldda), dtype=torch.float64, device=dev)
        self.ipiv = torch.zeros((batch, max(self.mn, 1)), dtype=torch.int32, device=dev)
        self.info = torch.full((batch,), -999, dtype=torch.int32, device=dev)
        self.dA_array = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.dipiv_array = torch.zeros(batch, dtype=torch.int64, device=dev)
        magma_dset_pointer(self.dA_array, self.A, self.ldda, 0, 0, n * self.ldda, batch, self.queue)
        magma_iset_pointer(self.dipiv_array, self.ipiv, 1, 0, 0, max(self.mn, 1), batch, self.queue)
        if nrhs:
            self.B = torch.zeros((batch, nrhs, self.lddb), dtype=torch.float64, device=dev)
            self.dB_array = torch.zeros(batch, dtype=torch.int64, device=dev)
            magma_dset_pointer(self.dB_array, self.B, self.lddb, 0, 0, nrhs * self.lddb, batch, self.queue)

    def upload(self, A_np, B_np=None):
        t = self.torch
        self.A.copy_(t.from_numpy(A_np))
        if B_np is not None:
            self.B.copy_(t.from_numpy(B_np))

    def getrf(self) -> int:
        return magma_dgetrf_batched(self.m, self.n, self.dA_array, self.ldda, self.dipiv_array, self.info,
                                    self.batch, self.queue)

    def getrs(self, trans=MagmaNoTrans) -> int:
        return magma_dgetrs_batched(trans, self.n, self.nrhs, self.dA_array, self.ldda, self.dipiv_array,
                                    self.dB_array, self.lddb, self.batch, self.queue)

    def gesv(self) -> int:
        return magma_dgesv_batched(self.n, self.nrhs, self.dA_array, self.ldda, self.dipiv_array, self.dB_array,
                                   self.lddb, self.info, self.batch, self.queue)

    def download(self):
        """(LU, ipiv[:, :min(m,n)], info[, X]) as numpy arrays, after syncing the queue."""
        self.queue.sync()
        self.torch.cuda.synchronize()
        out = [self.A.cpu().numpy(), self.ipiv.cpu().numpy()[:, :self.mn], self.info.cpu().numpy()]
        if self.nrhs:
            out.append(self.B.cpu().numpy())
        return tuple(out)
