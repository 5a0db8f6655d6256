"""s / c / z precisions of the four entry points through the C ABI, against oracle/lu_oracle_scz.c
(bit-exact: the kernels use the oracle's operation sequences) -- the structure of the reference's generated testers
testing_{s,c,z}getrf_batched.cpp / testing_{s,c,z}gesv_batched.cpp."""
import numpy as np
import pytest

import oracle
from magma_b200 import _lib
from magma_b200 import batched as mb

pytestmark = pytest.mark.gpu

TDT = {"s": "float32", "c": "complex64", "z": "complex128"}


def _dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def _ptrs(t, stride_elems, batch):
    """device pointer array to matrix b of a contiguous (batch, ...) tensor"""
    import torch
    return torch.tensor([t.data_ptr() + b * stride_elems * t.element_size() for b in range(batch)], dtype=torch.int64, device="cuda")


def _ipiv(batch, k):
    import torch
    ip = torch.zeros((batch, max(k, 1)), dtype=torch.int32, device="cuda")
    return ip, _ptrs(ip, max(k, 1), batch)


def _same(a, b):
    return np.array_equal(a.view(np.uint8), b.view(np.uint8)) or np.array_equal(a, b)


@pytest.mark.parametrize("p", ["s", "c", "z"])
@pytest.mark.parametrize("m,n,ld", [(1, 1, 1), (5, 7, 5), (7, 5, 9), (8, 8, 8), (16, 16, 16), (24, 17, 24), (32, 32, 32), (31, 32, 33),
                                    (33, 33, 33), (40, 50, 40), (50, 40, 52), (64, 64, 64), (100, 100, 100), (130, 70, 130),
                                    (70, 130, 72), (200, 200, 200), (300, 40, 300)])
def test_getrf_batched_prec(gpu_queue, p, m, n, ld):
    import torch
    batch = 6 if m * n <= 10000 else 3
    L = _lib.load()
    A0 = oracle.random_batch_prec(p, batch, m, n, ld=ld, seed=m * 131 + n)
    A0[:, :, m:] = 7.0  # padding must survive
    dA = _dev(A0)
    ip, ipp = _ipiv(batch, min(m, n))
    info = torch.full((batch,), -9, dtype=torch.int32, device="cuda")
    ap = _ptrs(dA, n * ld, batch)  # pointer arrays stay alive until the call has run
    rc = getattr(L, f"magma_{p}getrf_batched")(m, n, ap.data_ptr(), ld, ipp.data_ptr(), info.data_ptr(), batch, gpu_queue.handle)
    assert rc == 0
    gpu_queue.sync()
    ref = A0.copy()
    ipr, infr = oracle.getrf_batched_prec(p, ref, m)
    LU = dA.cpu().numpy()
    assert np.array_equal(ip.cpu().numpy()[:, :min(m, n)], ipr)
    assert np.array_equal(info.cpu().numpy(), infr)
    assert _same(LU, ref), f"max diff {np.abs(LU - ref).max()}"


@pytest.mark.parametrize("p", ["s", "c", "z"])
def test_getrf_prec_singular_and_ties(gpu_queue, p):
    import torch
    n = 24
    L = _lib.load()
    dt = oracle.PREC_DTYPE[p]
    rng = np.random.default_rng(3)
    mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)), rng.integers(-2, 3, size=(n, n)).astype(float)]
    Z = rng.random((n, n)); Z[:, 9] = 0.0; mats.append(Z)
    A0 = np.stack(mats).astype(dt)
    if p != "s":
        A0 = (A0 * (1 + 1j)).astype(dt)  # 1 / (1 + i) and the products with it are exact: the ties stay exact ties
    batch = len(mats)
    for big in (False, True):  # register kernel, then the CTA kernel (48 x 48 with the same blocks on the diagonal)
        if big:
            B0 = np.zeros((batch, 2 * n, 2 * n), dtype=dt)
            B0[:, :n, :n] = A0; B0[:, n:, n:] = A0
            X0, nn = B0, 2 * n
        else:
            X0, nn = A0, n
        dA = _dev(X0)
        ip, ipp = _ipiv(batch, nn)
        info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        ap = _ptrs(dA, nn * nn, batch)
        assert getattr(L, f"magma_{p}getrf_batched")(nn, nn, ap.data_ptr(), nn, ipp.data_ptr(), info.data_ptr(), batch,
                                                     gpu_queue.handle) == 0
        gpu_queue.sync()
        ref = X0.copy()
        ipr, infr = oracle.getrf_batched_prec(p, ref, nn)
        assert np.array_equal(ip.cpu().numpy(), ipr)
        assert np.array_equal(info.cpu().numpy(), infr) and infr[0] == 1
        assert _same(dA.cpu().numpy(), ref)


@pytest.mark.parametrize("p", ["s", "c", "z"])
@pytest.mark.parametrize("trans", [111, 112, 113])
@pytest.mark.parametrize("n,nrhs", [(1, 1), (9, 3), (32, 1), (32, 5), (50, 4), (128, 2), (300, 1)])
def test_getrs_batched_prec(gpu_queue, p, trans, n, nrhs):
    batch = 4
    L = _lib.load()
    A0 = oracle.random_batch_prec(p, batch, n, n, seed=n)
    B0 = oracle.random_batch_prec(p, batch, n, nrhs, seed=n + 1)
    LU = A0.copy()
    ipiv, _ = oracle.getrf_batched_prec(p, LU, n)
    dLU, dB = _dev(LU), _dev(B0)
    dip = _dev(ipiv.astype(np.int32))
    ap, pp, bp = _ptrs(dLU, n * n, batch), _ptrs(dip, n, batch), _ptrs(dB, nrhs * n, batch)
    rc = getattr(L, f"magma_{p}getrs_batched")(trans, n, nrhs, ap.data_ptr(), n, pp.data_ptr(), bp.data_ptr(), n, batch,
                                               gpu_queue.handle)
    assert rc == 0
    gpu_queue.sync()
    Xr = B0.copy()
    oracle.getrs_batched_prec(p, trans, LU, ipiv, Xr, n)
    X = dB.cpu().numpy()
    assert _same(X, Xr), f"max diff {np.abs(X - Xr).max()}"
    # the reference tester's residual, in this precision
    eps = np.finfo(oracle.PREC_DTYPE[p]).eps
    for b in range(batch):
        M = A0[b].T
        op = {111: M, 112: M.T, 113: M.conj().T}[trans]
        r = np.linalg.norm(op @ X[b].T - B0[b].T, 1) / (n * np.linalg.norm(M, 1) * np.linalg.norm(X[b].T, 1))
        assert r < 30 * eps


@pytest.mark.parametrize("p", ["s", "c", "z"])
@pytest.mark.parametrize("n,nrhs", [(4, 1), (16, 1), (32, 2), (48, 3), (100, 1)])
def test_gesv_batched_prec(gpu_queue, p, n, nrhs):
    import torch
    batch = 5
    L = _lib.load()
    A0 = oracle.random_batch_prec(p, batch, n, n, seed=2 * n)
    B0 = oracle.random_batch_prec(p, batch, n, nrhs, seed=2 * n + 1)
    dA, dB = _dev(A0), _dev(B0)
    ip, ipp = _ipiv(batch, n)
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ap, bp = _ptrs(dA, n * n, batch), _ptrs(dB, nrhs * n, batch)
    rc = getattr(L, f"magma_{p}gesv_batched")(n, nrhs, ap.data_ptr(), n, ipp.data_ptr(), bp.data_ptr(), n, info.data_ptr(), batch,
                                              gpu_queue.handle)
    assert rc == 0
    gpu_queue.sync()
    Ar, Br = A0.copy(), B0.copy()
    ipr, infr = oracle.gesv_batched_prec(p, Ar, Br, n)
    assert np.array_equal(ip.cpu().numpy(), ipr) and np.array_equal(info.cpu().numpy(), infr)
    assert _same(dA.cpu().numpy(), Ar) and _same(dB.cpu().numpy(), Br)


@pytest.mark.parametrize("p", ["s", "c", "z"])
@pytest.mark.parametrize("small", [True, False])
def test_getrf_vbatched_prec(gpu_queue, p, small):
    """Variable sizes (device arrays m, n, ldda): every matrix against the oracle; empty matrices allowed."""
    import torch
    L = _lib.load()
    rng = np.random.default_rng(11)
    batch = 23
    hi = 32 if small else 90
    ms = rng.integers(0, hi + 1, size=batch).astype(np.int32)
    ns = rng.integers(0, hi + 1, size=batch).astype(np.int32)
    ms[0], ns[0] = hi, hi
    lds = np.maximum(ms, 1).astype(np.int32) + rng.integers(0, 3, size=batch).astype(np.int32)
    dt = oracle.PREC_DTYPE[p]
    mats = [oracle.random_batch_prec(p, 1, int(lds[b]), int(ns[b]), seed=100 + b)[0] if ns[b] > 0 else np.zeros((0, lds[b]), dtype=dt)
            for b in range(batch)]
    offs = np.cumsum([0] + [x.size for x in mats])
    flat = np.concatenate([x.reshape(-1) for x in mats] + [np.zeros(1, dtype=dt)])
    dflat = _dev(flat)
    esz = dflat.element_size()
    aptr = torch.tensor([dflat.data_ptr() + int(offs[b]) * esz for b in range(batch)], dtype=torch.int64, device="cuda")
    kmax = max(1, int(np.minimum(ms, ns).max()))
    ip, ipp = _ipiv(batch, kmax)
    info = torch.full((batch,), -3, dtype=torch.int32, device="cuda")
    dm, dn, dl = _dev(ms), _dev(ns), _dev(lds)
    rc = getattr(L, f"magma_{p}getrf_vbatched")(dm.data_ptr(), dn.data_ptr(), aptr.data_ptr(), dl.data_ptr(), ipp.data_ptr(),
                                                info.data_ptr(), batch, gpu_queue.handle)
    assert rc == 0
    gpu_queue.sync()
    out = dflat.cpu().numpy()
    iph, infh = ip.cpu().numpy(), info.cpu().numpy()
    for b in range(batch):
        m, n, ld = int(ms[b]), int(ns[b]), int(lds[b])
        if m == 0 or n == 0:
            assert infh[b] == 0
            continue
        ref = mats[b].copy().reshape(1, n, ld)
        ipr, infr = oracle.getrf_batched_prec(p, ref, m)
        got = out[offs[b]:offs[b + 1]].reshape(1, n, ld)
        assert np.array_equal(iph[b, :min(m, n)], ipr[0]) and infh[b] == infr[0], (b, m, n)
        assert _same(got, ref), (b, m, n)


def test_prec_argument_errors(gpu_queue):
    L = _lib.load()
    q = gpu_queue.handle
    for p in "scz":
        assert getattr(L, f"magma_{p}getrf_batched")(-1, 4, None, 4, None, None, 1, q) == -1
        assert getattr(L, f"magma_{p}getrf_batched")(4, -1, None, 4, None, None, 1, q) == -2
        assert getattr(L, f"magma_{p}getrf_batched")(4, 4, None, 3, None, None, 1, q) == -4
        assert getattr(L, f"magma_{p}getrf_batched")(0, 4, None, 1, None, None, 1, q) == 0
        assert getattr(L, f"magma_{p}getrs_batched")(7, 4, 1, None, 4, None, None, 4, 1, q) == -1
        assert getattr(L, f"magma_{p}getrs_batched")(111, 4, 1, None, 3, None, None, 4, 1, q) == -5
        assert getattr(L, f"magma_{p}getrs_batched")(111, 4, 1, None, 4, None, None, 3, 1, q) == -8
        assert getattr(L, f"magma_{p}gesv_batched")(4, 1, None, 3, None, None, 4, None, 1, q) == -4
        assert getattr(L, f"magma_{p}gesv_batched")(4, 1, None, 4, None, None, 3, None, 1, q) == -6
        assert getattr(L, f"magma_{p}getrf_vbatched")(None, None, None, None, None, None, -1, q) == -7
