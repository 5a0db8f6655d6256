"""GPU tests added in round 2 (run with -m gpu on the B200 box), all through the C ABI:
  * the single-launch shared-memory tier (lu_fused.cu), forced on over its whole range,
  * the panel-level / BLAS-level source-compatibility entry points (compat.cu),
  * the single-process multi-GPU entry points on two devices,
  * shapes the round-1 suite never reached (tall panels up to 9000 rows, n = 1024, batchCount > 2^24),
  * the compiled reference GPU path (oracle/_ref, run in a child process) as a second witness,
  * a C program compiled and LINKED against libmagma_b200.so that walks the tester sequence.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from magma_b200 import _lib
from magma_b200 import batched as mb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from test_gpu_parity import check_against_oracle, run_getrf  # noqa: E402


# ---- single-launch tier --------------------------------------------------------------------------------------

@pytest.fixture
def fused_tier():
    mb.set_fused_max(128)
    yield
    mb.set_fused_max(0)


def test_rcp_selftest(gpu_queue):
    """The chain's inline reciprocal equals IEEE 1.0/x bit for bit on 50M pseudo-random inputs."""
    assert mb.rcp_selftest(50_000_000, gpu_queue) == 0


@pytest.mark.parametrize("m,n,batch", [(128, 128, 9), (64, 64, 9), (40, 40, 11), (33, 33, 7), (65, 65, 7), (96, 96, 5),
                                       (100, 100, 5), (127, 127, 4), (128, 64, 5), (64, 128, 5), (128, 72, 4), (72, 128, 4),
                                       (97, 113, 4), (113, 97, 4), (48, 120, 4), (120, 48, 4), (128, 8, 4), (8, 128, 4),
                                       (128, 1, 3), (1, 128, 3), (70, 66, 4), (66, 70, 4), (128, 104, 3), (35, 90, 3)])
def test_fused_tier(gpu_queue, fused_tier, m, n, batch):
    A0, _ = oracle.random_batch(batch, m, n)
    check_against_oracle(gpu_queue, A0, m)


@pytest.mark.parametrize("m,n,ldda", [(128, 128, 130), (128, 128, 129), (100, 100, 101), (64, 64, 72), (127, 127, 128),
                                      (90, 90, 90)])
def test_fused_tier_ldda(gpu_queue, fused_tier, m, n, ldda):
    """Even/aligned shapes take the TMA bulk path, odd leading dimensions and odd row counts the plain-load path."""
    A0, _ = oracle.random_batch(4, m, n)
    check_against_oracle(gpu_queue, A0, m, ldda=ldda)


def test_fused_tier_singular_and_ties(gpu_queue, fused_tier):
    rng = np.random.default_rng(5)
    for n in (40, 72, 128):
        mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)),
                rng.integers(-3, 4, size=(n, n)).astype(float)]
        Z = rng.random((n, n))
        Z[:, n // 3] = 0.0
        mats.append(Z)
        Z2 = rng.random((n, n))
        Z2[n // 2:, :] = Z2[:n - n // 2, :]
        mats.append(Z2)
        Z3 = rng.random((n, n)) * 1e-305  # pivots below 2^-999: the reciprocal's out-of-range branch (1/x still finite)
        mats.append(Z3)
        Z4 = rng.random((n, n)) * 1e300
        mats.append(Z4)
        check_against_oracle(gpu_queue, np.stack(mats), n)


def test_fused_tier_full_size_sample(gpu_queue, fused_tier):
    """C3 shape at a batch that fills the GPU several times over: bit-exact sample + residual on all."""
    n, batch = 128, 3000
    A0, _ = oracle.random_batch(batch, n, n)
    check_against_oracle(gpu_queue, A0, n)


@pytest.mark.parametrize("chain", [0, 1, 3, 4])
@pytest.mark.parametrize("m,n,batch", [(128, 128, 7), (100, 100, 5), (97, 97, 5), (128, 40, 5), (120, 200, 4), (127, 96, 4), (226, 130, 3),
                                       (64, 64, 6), (65, 90, 4), (50, 45, 5), (33, 60, 4)])
def test_chain_panel_switch(gpu_queue, chain, m, n, batch):
    """Panels of at most 128 rows: the single-warp chain kernel (default: 33..128 rows) and the one-thread-per-row kernel
    give the same bits at every switch level."""
    mb.set_chain_panel(chain)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_chain_panel(3)


def test_chain_panel_structured(gpu_queue):
    rng = np.random.default_rng(9)
    n = 128
    mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)), rng.integers(-3, 4, size=(n, n)).astype(float)]
    Z = rng.random((n, n)); Z[:, 5] = 0.0; Z[:, 100] = 0.0; mats.append(Z)
    Z2 = rng.random((n, n)); Z2[n // 2:, :] = Z2[:n - n // 2, :]; mats.append(Z2)
    check_against_oracle(gpu_queue, np.stack(mats), n)


# ---- source-compatibility entry points ---------------------------------------------------------------------------

def _panel_case(q, fn_name, m, n, ai, batch=5, M=150, N=70):
    """Factor the m x n block at (ai, ai) of M x N matrices through one of the panel entry points."""
    import torch
    L = _lib.load()
    A0, _ = oracle.random_batch(batch, M, N)
    db = mb.DeviceBatch(batch, M, N, queue=q)
    db.upload(A0)
    db.info.fill_(0)
    dummy = torch.zeros(batch, dtype=torch.int64, device="cuda")
    if fn_name == "magma_dgetf2_fused_batched":
        rc = L.magma_dgetf2_fused_batched(m, n, mb.ptr(db.dA_array), ai, ai, M, mb.ptr(db.dipiv_array), mb.ptr(db.info), batch,
                                          q.handle)
    elif fn_name == "magma_dgetf2_batched":
        rc = L.magma_dgetf2_batched(m, n, mb.ptr(db.dA_array), ai, ai, M, mb.ptr(db.dipiv_array), mb.ptr(dummy),
                                    mb.ptr(db.info), ai, batch, q.handle)
    else:
        rc = L.magma_dgetrf_recpanel_batched(m, n, 32, mb.ptr(db.dA_array), ai, ai, M, mb.ptr(db.dipiv_array),
                                             mb.ptr(dummy), mb.ptr(db.info), ai, batch, q.handle)
    assert rc == 0
    q.sync()
    LU = db.A.cpu().numpy()
    ipiv = db.ipiv.cpu().numpy()
    # oracle on the block alone
    blk = np.ascontiguousarray(A0[:, ai:ai + n, ai:ai + m]).copy()
    ipr, infr = oracle.getrf_batched(blk, m)
    assert np.array_equal(LU[:, ai:ai + n, ai:ai + m], blk), "panel factors differ"
    assert np.array_equal(ipiv[:, ai:ai + min(m, n)], ipr), "panel pivots must be relative to row ai"
    # nothing outside the block is touched
    mask = np.ones_like(A0, dtype=bool)
    mask[:, ai:ai + n, ai:ai + m] = False
    assert np.array_equal(LU[mask], A0[mask])
    assert np.array_equal(db.info.cpu().numpy(), np.where(infr != 0, infr + ai, 0))


@pytest.mark.parametrize("fn", ["magma_dgetf2_fused_batched", "magma_dgetf2_batched", "magma_dgetrf_recpanel_batched"])
@pytest.mark.parametrize("m,n,ai", [(150, 32, 0), (118, 32, 32), (86, 6, 64), (40, 8, 20)])
def test_panel_entry_points(gpu_queue, fn, m, n, ai):
    _panel_case(gpu_queue, fn, m, n, ai)


def test_recpanel_wide_panel(gpu_queue):
    _panel_case(gpu_queue, "magma_dgetrf_recpanel_batched", 120, 64, 6)


def test_getf2_fused_rejects_wide(gpu_queue):
    L = _lib.load()
    assert L.magma_dgetf2_fused_batched(64, 33, 0, 0, 0, 64, 0, 0, 1, gpu_queue.handle) == -2


def test_panel_info_merge(gpu_queue):
    """A zero column inside the block records gbstep + step + 1 unless an earlier panel already recorded one."""
    import torch
    L = _lib.load()
    M = 64
    A0, _ = oracle.random_batch(3, M, M)
    A0[:, 20, :] = 0.0  # column 20 is exactly zero
    db = mb.DeviceBatch(3, M, M, queue=gpu_queue)
    db.upload(A0)
    db.info.copy_(torch.tensor([0, 7, 0], dtype=torch.int32))
    assert L.magma_dgetf2_fused_batched(M - 16, 16, mb.ptr(db.dA_array), 16, 16, M, mb.ptr(db.dipiv_array), mb.ptr(db.info), 3,
                                        gpu_queue.handle) == 0
    gpu_queue.sync()
    assert db.info.cpu().tolist() == [21, 7, 21]


def test_laswp_rowparallel(gpu_queue):
    """new[r] = old[pivinfo[r]-1] for the top rows (to the output) and for the rows they came from (in place)."""
    import torch
    L = _lib.load()
    batch, m, n, h = 4, 50, 19, 8
    rng = np.random.default_rng(3)
    A0 = rng.random((batch, n, m))
    piv = np.zeros((batch, m), dtype=np.int32)
    for b in range(batch):
        # permutation as setup_pivinfo builds it from h sequential interchanges
        p = np.arange(m)
        for i in range(h):
            j = rng.integers(i, m)
            p[[i, j]] = p[[j, i]]
        piv[b] = p + 1
    dA = torch.from_numpy(A0).cuda()
    dP = torch.from_numpy(piv).cuda()
    pA = torch.tensor([dA.data_ptr() + b * n * m * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
    pP = torch.tensor([dP.data_ptr() + b * m * 4 for b in range(batch)], dtype=torch.int64, device="cuda")
    L.magma_dlaswp_rowparallel_batched(n, mb.ptr(pA), 0, 0, m, mb.ptr(pA), 0, 0, m, 0, h, mb.ptr(pP), batch, gpu_queue.handle)
    gpu_queue.sync()
    out = dA.cpu().numpy()
    for b in range(batch):
        p = piv[b] - 1
        exp = A0[b].copy()
        touched = set(range(h)) | set(p[:h].tolist())
        for r in touched:
            exp[:, r] = A0[b][:, p[r]]
        assert np.array_equal(out[b], exp)


@pytest.mark.parametrize("uplo", [mb.MagmaLower, mb.MagmaUpper])
@pytest.mark.parametrize("trans", [mb.MagmaNoTrans, mb.MagmaTrans])
@pytest.mark.parametrize("diag", [mb.MagmaUnit, mb.MagmaNonUnit])
@pytest.mark.parametrize("n,incb", [(1, 1), (31, 1), (32, 2), (100, 1), (257, 3)])
def test_trsv_batched(gpu_queue, uplo, trans, diag, n, incb):
    import torch
    L = _lib.load()
    batch = 3
    rng = np.random.default_rng(n)
    A = rng.random((batch, n, n)) + n * np.eye(n)  # [b, col, row], well conditioned
    x0 = rng.random((batch, n * incb))
    dA = torch.from_numpy(A).cuda()
    dx = torch.from_numpy(x0).cuda()
    pA = torch.tensor([dA.data_ptr() + b * n * n * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
    px = torch.tensor([dx.data_ptr() + b * n * incb * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
    L.magmablas_dtrsv_batched(uplo, trans, diag, n, mb.ptr(pA), n, mb.ptr(px), incb, batch, gpu_queue.handle)
    gpu_queue.sync()
    x = dx.cpu().numpy()
    for b in range(batch):
        M = A[b].T  # M[row, col]
        T = np.tril(M) if uplo == mb.MagmaLower else np.triu(M)
        if diag == mb.MagmaUnit:
            np.fill_diagonal(T, 1.0)
        if trans != mb.MagmaNoTrans:
            T = T.T
        ref = np.linalg.solve(T, x0[b, ::incb])
        assert np.allclose(x[b, ::incb], ref, rtol=1e-11, atol=1e-13)
        if incb > 1:  # the gaps are untouched
            keep = np.ones(n * incb, dtype=bool)
            keep[::incb] = False
            assert np.array_equal(x[b][keep], x0[b][keep])


# ---- standalone BLAS-3 and strided front ends -----------------------------------------------------------------------

def _ptrs(t, stride_elems, batch, esize=8):
    import torch
    return torch.tensor([t.data_ptr() + esize * stride_elems * b for b in range(batch)], dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("ta", [mb.MagmaNoTrans, mb.MagmaTrans])
@pytest.mark.parametrize("tb", [mb.MagmaNoTrans, mb.MagmaTrans])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (65, 63, 17), (100, 130, 45), (200, 96, 128)])
def test_dgemm_batched_all_transposes(gpu_queue, ta, tb, m, n, k):
    import torch
    L = _lib.load()
    batch = 3
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    ar, ac = (m, k) if ta == mb.MagmaNoTrans else (k, m)
    br, bc = (k, n) if tb == mb.MagmaNoTrans else (n, k)
    A = rng.random((batch, ac, ar + 2)) - 0.5      # stored [col, row], lda = ar + 2
    B = rng.random((batch, bc, br + 1)) - 0.5
    Cm = rng.random((batch, n, m + 3)) - 0.5
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, Cm))
    pa, pb, pc = _ptrs(dA, A[0].size, batch), _ptrs(dB, B[0].size, batch), _ptrs(dC, Cm[0].size, batch)  # kept alive
    for alpha, beta in ((1.5, -0.5), (2.0, 0.0), (-1.0, 1.0)):
        dC.copy_(torch.from_numpy(Cm))
        L.magma_dgemm_batched(ta, tb, m, n, k, alpha, mb.ptr(pa), ar + 2, mb.ptr(pb), br + 1, beta, mb.ptr(pc), m + 3,
                              batch, gpu_queue.handle)
        gpu_queue.sync()
        got = dC.cpu().numpy()
        for b in range(batch):
            Am = A[b].T[:ar, :ac]
            Bm = B[b].T[:br, :bc]
            opA = Am if ta == mb.MagmaNoTrans else Am.T
            opB = Bm if tb == mb.MagmaNoTrans else Bm.T
            ref = alpha * (opA @ opB) + beta * Cm[b].T[:m, :n]
            assert np.allclose(got[b].T[:m, :n], ref, rtol=1e-12, atol=1e-12)
            assert np.array_equal(got[b].T[m:, :], Cm[b].T[m:, :]), "wrote into the ldc padding"


@pytest.mark.parametrize("m,n,k", [(64, 64, 32), (130, 70, 64), (33, 200, 100)])
def test_dgemm_batched_lu_update_is_canonical(gpu_queue, m, n, k):
    """alpha = -1, beta = 1 (the LU trailing update): bit-identical to the fma chain with k increasing."""
    import torch
    L = _lib.load()
    batch = 2
    rng = np.random.default_rng(k)
    A = rng.random((batch, k, m)); B = rng.random((batch, n, k)); Cm = rng.random((batch, n, m))
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, Cm))
    pa, pb, pc = _ptrs(dA, A[0].size, batch), _ptrs(dB, B[0].size, batch), _ptrs(dC, Cm[0].size, batch)  # kept alive
    L.magma_dgemm_batched(mb.MagmaNoTrans, mb.MagmaNoTrans, m, n, k, -1.0, mb.ptr(pa), m, mb.ptr(pb), k, 1.0, mb.ptr(pc), m,
                          batch, gpu_queue.handle)
    gpu_queue.sync()
    got = dC.cpu().numpy()
    for b in range(batch):
        ref = Cm[b].copy()
        oracle.gemm_lu_update(np.ascontiguousarray(A[b]), np.ascontiguousarray(B[b]), ref)
        assert np.array_equal(got[b], ref)


@pytest.mark.parametrize("uplo", [mb.MagmaLower, mb.MagmaUpper])
@pytest.mark.parametrize("trans", [mb.MagmaNoTrans, mb.MagmaTrans])
@pytest.mark.parametrize("diag", [mb.MagmaUnit, mb.MagmaNonUnit])
@pytest.mark.parametrize("m,n", [(1, 1), (5, 31), (130, 33), (40, 100)])
def test_dtrsm_batched_right(gpu_queue, uplo, trans, diag, m, n):
    import torch
    batch = 3
    rng = np.random.default_rng(m + 7 * n)
    T = rng.random((batch, n, n)) + n * np.eye(n)
    B = rng.random((batch, n, m + 1)) - 0.5       # m x n stored [col, row], ldb = m + 1
    dT, dB = torch.from_numpy(T).cuda(), torch.from_numpy(B).cuda()
    pt, pb = _ptrs(dT, n * n, batch), _ptrs(dB, B[0].size, batch)  # kept alive until the sync
    mb.magmablas_dtrsm_batched(mb.MagmaRight, uplo, trans, diag, m, n, 1.5, pt, n, pb, m + 1, batch, gpu_queue)
    gpu_queue.sync()
    X = dB.cpu().numpy()
    for b in range(batch):
        Tm = T[b].T
        Tm = np.tril(Tm) if uplo == mb.MagmaLower else np.triu(Tm)
        if diag == mb.MagmaUnit:
            Tm = Tm - np.diag(np.diag(Tm)) + np.eye(n)
        op = Tm if trans == mb.MagmaNoTrans else Tm.T
        ref = 1.5 * B[b].T[:m, :] @ np.linalg.inv(op)
        assert np.allclose(X[b].T[:m, :], ref, rtol=1e-10, atol=1e-12)
        assert np.array_equal(X[b].T[m:, :], B[b].T[m:, :])


@pytest.mark.parametrize("n,nrhs,batch", [(16, 1, 300), (40, 2, 30), (130, 3, 5)])
def test_strided_front_ends(gpu_queue, n, nrhs, batch):
    """dA + b*stride forms: same results as the pointer-array entry points (and hence as the oracle)."""
    import torch
    L = _lib.load()
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    Ar, Br = A0.copy(), B0.copy()
    ipr, infr = oracle.gesv_batched(Ar, Br, n)
    dA, dB = torch.from_numpy(A0).cuda(), torch.from_numpy(B0).cuda()
    ip = torch.zeros((batch, n), dtype=torch.int32, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    assert L.magma_dgesv_batched_strided(n, nrhs, mb.ptr(dA), n, n * n, mb.ptr(ip), n, mb.ptr(dB), n, n * nrhs, mb.ptr(info),
                                         batch, gpu_queue.handle) == 0
    gpu_queue.sync()
    assert np.array_equal(ip.cpu().numpy(), ipr) and np.array_equal(dA.cpu().numpy(), Ar)
    assert np.array_equal(dB.cpu().numpy(), Br)
    # getrf + getrs, strided
    dA.copy_(torch.from_numpy(A0)); dB.copy_(torch.from_numpy(B0))
    assert L.magma_dgetrf_batched_strided(n, n, mb.ptr(dA), n, n * n, mb.ptr(ip), n, mb.ptr(info), batch, gpu_queue.handle) == 0
    assert L.magma_dgetrs_batched_strided(mb.MagmaNoTrans, n, nrhs, mb.ptr(dA), n, n * n, mb.ptr(ip), n, mb.ptr(dB), n,
                                          n * nrhs, batch, gpu_queue.handle) == 0
    gpu_queue.sync()
    assert np.array_equal(dA.cpu().numpy(), Ar) and np.array_equal(ip.cpu().numpy(), ipr)
    assert np.array_equal(dB.cpu().numpy(), Br)


# ---- random butterfly transformation ---------------------------------------------------------------------------------

@pytest.mark.parametrize("n,nrhs,batch", [(1, 1, 3), (2, 2, 3), (7, 1, 5), (16, 3, 9), (33, 2, 5), (64, 4, 4), (100, 1, 3), (257, 2, 2)])
def test_rbt_pieces_match_oracle(gpu_queue, n, nrhs, batch):
    """magma_dgerbt_batched with the caller's butterflies (gen = MagmaFalse): A and B bit-identical to the restatement;
    magmablas_dprbt_mv_batched likewise."""
    import torch
    L = _lib.load()
    rng = np.random.default_rng(n)
    u = np.exp((rng.random(2 * n) - 0.5) / 10)
    v = np.exp((rng.random(2 * n) - 0.5) / 10)
    A0 = rng.random((batch, n, n)) - 0.5
    B0 = rng.random((batch, nrhs, n)) - 0.5
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    db.upload(A0, B0)
    info = C.c_int(99)
    rc = L.magma_dgerbt_batched(0, n, nrhs, mb.ptr(db.dA_array), n, mb.ptr(db.dB_array), n, u.ctypes.data, v.ctypes.data,
                                C.addressof(info), batch, gpu_queue.handle)
    assert rc == 0 and info.value == 0
    gpu_queue.sync()
    Ar, Br = A0.copy(), B0.copy()
    oracle.prbt(Ar, n, u, v)
    oracle.prbt_mtv(Br, n, u)
    assert np.array_equal(db.A.cpu().numpy(), Ar)
    assert np.array_equal(db.B.cpu().numpy(), Br)
    dv = torch.from_numpy(v).cuda()
    L.magmablas_dprbt_mv_batched(n, nrhs, mb.ptr(dv), mb.ptr(db.dB_array), n, batch, gpu_queue.handle)
    gpu_queue.sync()
    oracle.prbt_mv(Br, n, v)
    assert np.array_equal(db.B.cpu().numpy(), Br)


@pytest.mark.parametrize("n,nrhs,batch", [(8, 1, 11), (32, 2, 7), (48, 3, 5), (100, 1, 4), (200, 2, 3), (512, 1, 2)])
def test_gesv_rbt(gpu_queue, n, nrhs, batch):
    """magma_dgesv_rbt_batched draws U, V with rand() like the reference: seed the C library, redraw the same numbers
    here, and the solution must equal the restatement bit for bit; the residual check of the testers holds loosely
    (no pivoting: 1e-9 instead of 30 eps)."""
    L = _lib.load()
    libc = C.CDLL(None)
    libc.rand.restype = C.c_int
    RAND_MAX = 2147483647
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    db.upload(A0, B0)
    libc.srand(4242)
    rc = L.magma_dgesv_rbt_batched(n, nrhs, mb.ptr(db.dA_array), n, mb.ptr(db.dB_array), n, mb.ptr(db.info), batch,
                                   gpu_queue.handle)
    assert rc == 0
    gpu_queue.sync()
    libc.srand(4242)
    u, v = np.zeros(2 * n), np.zeros(2 * n)
    import math
    for i in range(2 * n):
        u[i] = math.exp((((libc.rand() * 1.0) / RAND_MAX) - 0.5) / 10)
        v[i] = math.exp((((libc.rand() * 1.0) / RAND_MAX) - 0.5) / 10)
    Ar, Br = A0.copy(), B0.copy()
    infr = oracle.gesv_rbt_batched(Ar, Br, n, u, v)
    X = db.B.cpu().numpy()
    assert np.array_equal(db.info.cpu().numpy(), infr)
    assert np.array_equal(db.A.cpu().numpy(), Ar)
    assert np.array_equal(X, Br)
    assert oracle.solve_residual(oracle.MagmaNoTrans, A0, X, B0, n) < 1e-9


def test_gesv_rbt_argument_errors(gpu_queue, capfd):
    L = _lib.load()
    assert L.magma_dgesv_rbt_batched(-1, 1, 0, 1, 0, 1, 0, 1, gpu_queue.handle) == -1
    assert L.magma_dgesv_rbt_batched(4, 1, 0, 3, 0, 4, 0, 1, gpu_queue.handle) == -4
    assert L.magma_dgesv_rbt_batched(4, 1, 0, 4, 0, 3, 0, 1, gpu_queue.handle) == -6
    assert L.magma_dgesv_rbt_batched(0, 1, 0, 1, 0, 1, 0, 1, gpu_queue.handle) == 0
    assert "magma_dgesv_rbt_batched" in capfd.readouterr().err


# ---- single-process multi-GPU -------------------------------------------------------------------------------------

def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.parametrize("n", [16, 128, 200])
def test_mgpu_getrf_two_devices(n):
    """magma_b200_dgetrf_batched_mgpu over per-device queues: kernels that opt in to > 48 KB of shared memory must
    do so on EVERY device (the attribute is per device), and each shard must equal the oracle bit for bit."""
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    import torch
    from magma_b200 import mgpu
    L = _lib.load()
    assert mb.magma_init() == 0
    batch = 37
    A0, _ = oracle.random_batch(batch, n, n)
    ref = A0.copy()
    ipr, infr = oracle.getrf_batched(ref, n)
    ngpu = 2
    queues, dbs, cnts = [], [], []
    for g in range(ngpu):
        torch.cuda.set_device(g)
        q = mb.Queue(g)
        lo, hi = mgpu.shard_range(batch, ngpu, g)
        db = mb.DeviceBatch(hi - lo, n, n, device=g, queue=q)
        db.upload(A0[lo:hi])
        queues.append(q); dbs.append(db); cnts.append(hi - lo)
    torch.cuda.set_device(0)
    PP = C.c_void_p * ngpu
    dA = PP(*[db.dA_array.data_ptr() for db in dbs])
    dP = PP(*[db.dipiv_array.data_ptr() for db in dbs])
    dI = PP(*[db.info.data_ptr() for db in dbs])
    cn = (C.c_int * ngpu)(*cnts)
    qs = PP(*[q.handle for q in queues])
    rc = L.magma_b200_dgetrf_batched_mgpu(ngpu, n, n, C.addressof(dA), n, C.addressof(dP), C.addressof(dI), C.addressof(cn),
                                          C.addressof(qs))
    assert rc == 0
    for g in range(ngpu):
        torch.cuda.set_device(g)
        queues[g].sync()
        torch.cuda.synchronize(g)
        lo, hi = mgpu.shard_range(batch, ngpu, g)
        assert np.array_equal(dbs[g].ipiv.cpu().numpy()[:, :n], ipr[lo:hi]), f"device {g}"
        assert np.array_equal(dbs[g].A.cpu().numpy(), ref[lo:hi]), f"device {g}"
        assert np.array_equal(dbs[g].info.cpu().numpy(), infr[lo:hi])
    torch.cuda.set_device(0)


def test_mgpu_gesv_two_devices():
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    import torch
    from magma_b200 import mgpu
    L = _lib.load()
    n, nrhs, batch = 96, 3, 21
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    Ar, Br = A0.copy(), B0.copy()
    ipr, infr = oracle.gesv_batched(Ar, Br, n)
    ngpu = 2
    queues, dbs, cnts = [], [], []
    for g in range(ngpu):
        torch.cuda.set_device(g)
        q = mb.Queue(g)
        lo, hi = mgpu.shard_range(batch, ngpu, g)
        db = mb.DeviceBatch(hi - lo, n, n, nrhs=nrhs, device=g, queue=q)
        db.upload(A0[lo:hi], B0[lo:hi])
        queues.append(q); dbs.append(db); cnts.append(hi - lo)
    torch.cuda.set_device(0)
    PP = C.c_void_p * ngpu
    arr = lambda f: PP(*[f(db) for db in dbs])  # noqa: E731
    dA, dP, dB, dI = arr(lambda d: d.dA_array.data_ptr()), arr(lambda d: d.dipiv_array.data_ptr()), \
        arr(lambda d: d.dB_array.data_ptr()), arr(lambda d: d.info.data_ptr())
    cn = (C.c_int * ngpu)(*cnts)
    qs = PP(*[q.handle for q in queues])
    rc = L.magma_b200_dgesv_batched_mgpu(ngpu, n, nrhs, C.addressof(dA), n, C.addressof(dP), C.addressof(dB), n,
                                         C.addressof(dI), C.addressof(cn), C.addressof(qs))
    assert rc == 0
    for g in range(ngpu):
        torch.cuda.set_device(g)
        queues[g].sync()
        torch.cuda.synchronize(g)
        lo, hi = mgpu.shard_range(batch, ngpu, g)
        assert np.array_equal(dbs[g].B.cpu().numpy(), Br[lo:hi]), f"device {g}"
    torch.cuda.set_device(0)


# ---- ranges round 1 never reached ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("m,n,batch", [(1536, 16, 2), (1536, 40, 2), (3000, 24, 2), (5000, 16, 1), (5000, 40, 1), (9000, 16, 1),
                                       (9000, 33, 1), (2049, 8, 2)])
def test_tall_panels(gpu_queue, m, n, batch):
    """R = 4 / 8 / 16 register panels (1025..8192 rows) and the global-memory panel above 8192 rows."""
    A0, _ = oracle.random_batch(batch, m, n)
    check_against_oracle(gpu_queue, A0, m)


def test_square_1024(gpu_queue):
    """Right-looking driver over many steps (more than 512 rows)."""
    A0, _ = oracle.random_batch(2, 1024, 1024)
    check_against_oracle(gpu_queue, A0, 1024)


def test_wide_600x1500(gpu_queue):
    A0, _ = oracle.random_batch(2, 600, 1500)
    check_against_oracle(gpu_queue, A0, 600)


def test_batchcount_above_2_24(gpu_queue):
    """One call with more matrices than a launch chunk (2^24): the chunk loop. n = 2: 16.8M x 32 bytes."""
    import torch
    n = 2
    batch = (1 << 24) + 1000
    db = mb.DeviceBatch(batch, n, n, queue=gpu_queue)
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, batch * n * n, db.A, gpu_queue)
    gpu_queue.sync()
    A0 = db.A.clone()
    assert db.getrf() == 0
    gpu_queue.sync()
    # 2 x 2 LU in closed form on the device, every matrix: pivot = larger first-column entry
    a, c, b, d = A0[:, 0, 0], A0[:, 0, 1], A0[:, 1, 0], A0[:, 1, 1]   # A[b, col, row]
    swap = c.abs() > a.abs()
    p = torch.where(swap, c, a)
    lo = torch.where(swap, a, c)
    u01 = torch.where(swap, d, b)
    r1 = torch.where(swap, b, d)
    l = lo * (1.0 / p)
    u11 = torch.addcmul(r1, -l, u01)  # not fused on every backend: compare with a tolerance of 1 ulp
    assert torch.equal(db.A[:, 0, 0], p) and torch.equal(db.A[:, 0, 1], l) and torch.equal(db.A[:, 1, 0], u01)
    assert torch.allclose(db.A[:, 1, 1], u11, rtol=4e-16, atol=0)
    assert torch.equal(db.ipiv[:, 0], torch.where(swap, 2, 1).to(torch.int32))
    assert torch.equal(db.ipiv[:, 1], torch.full_like(db.ipiv[:, 1], 2))
    assert int(db.info.abs().max()) == 0
    # the tail beyond the first chunk against the oracle, bit for bit
    tail = A0[(1 << 24) - 8:].cpu().numpy()
    ref = tail.copy()
    ipr, _ = oracle.getrf_batched(ref, n)
    assert np.array_equal(db.A[(1 << 24) - 8:].cpu().numpy(), ref)
    assert np.array_equal(db.ipiv[(1 << 24) - 8:].cpu().numpy(), ipr)


# ---- the compiled reference GPU path as a second witness ------------------------------------------------------------------

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmagma_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libmagma_ref.so not built (oracle/build_ref.py)")
@pytest.mark.parametrize("n,batch,nrhs", [(32, 1000, 0), (16, 4000, 1), (128, 100, 0), (512, 6, 0), (512, 4, 16)])
def test_against_reference_gpu(gpu_queue, tmp_path, n, batch, nrhs):
    """Same dlarnv inputs through MAGMA 2.10.0's own kernels (+cuBLAS), run in a child process because both
    libraries export the same names: pivots identical, both inside the testers' 30 eps backward-error bar."""
    out = tmp_path / "ref.npz"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_run.py"), str(n), str(batch), str(nrhs), str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = np.load(out)
    A0, seed = oracle.random_batch(batch, n, n)
    if nrhs:
        B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
        db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
        db.upload(A0, B0)
        assert db.gesv() == 0
        LU, ipiv, info, X = db.download()
        assert oracle.solve_residual(oracle.MagmaNoTrans, A0, X, B0, n) < oracle.TOL
        assert oracle.solve_residual(oracle.MagmaNoTrans, A0, ref["X"], B0, n) < oracle.TOL
    else:
        LU, ipiv, info = run_getrf(gpu_queue, A0, n)
    assert np.array_equal(ipiv, ref["ipiv"]), "pivots differ from the reference's GPU path"
    assert np.array_equal(info, ref["info"])
    assert oracle.lu_backward_error(A0, LU, ipiv, n) < oracle.TOL
    assert oracle.lu_backward_error(A0, ref["LU"], ref["ipiv"], n) < oracle.TOL
    # same algorithm, different summation order in the trailing update: factors agree to rounding
    scale = np.max(np.abs(ref["LU"]))
    assert np.max(np.abs(LU - ref["LU"])) <= 1e-9 * scale


INTERPOSE_SO = os.path.join(ROOT, "magma_b200", "lib", "libmagma_b200_interpose.so")


@pytest.mark.skipif(not (os.path.exists(REF_SO) and os.path.exists(INTERPOSE_SO)),
                    reason="needs oracle/_ref/libmagma_ref.so and lib/libmagma_b200_interpose.so")
@pytest.mark.parametrize("n,batch,nrhs", [(16, 500, 1), (32, 300, 0), (100, 20, 0), (300, 4, 3)])
def test_interpose_mode(tmp_path, n, batch, nrhs):
    """INTEGRATION.md mode 2: our batched-LU entry points on a queue that belongs to a real libmagma (here the compiled
    reference), whose struct layout this library never touches. Results must equal the oracle bit for bit."""
    out = tmp_path / "ip.npz"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "interpose_run.py"), str(n), str(batch), str(nrhs), str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    A0, seed = oracle.random_batch(batch, n, n)
    if nrhs:
        B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
        Ar, Br = A0.copy(), B0.copy()
        ipr, infr = oracle.gesv_batched(Ar, Br, n)
        assert np.array_equal(got["X"], Br)
    else:
        Ar = A0.copy()
        ipr, infr = oracle.getrf_batched(Ar, n)
    assert np.array_equal(got["ipiv"], ipr) and np.array_equal(got["info"], infr)
    assert np.array_equal(got["LU"], Ar)


# ---- a C program linked against the library -----------------------------------------------------------------------------------

C_PROG = r'''
#include <stdio.h>
#include <stdlib.h>
#include "magma_v2.h"
/* testing/testing_zgetrf_batched.cpp:161-206 (z -> d): allocate, set pointers, factor, copy back */
int main(void)
{
    magma_int_t n = 48, batch = 6, ldda = 48, i, b;
    double *hA, *dA; magma_int_t *hipiv, *dipiv, *dinfo, hinfo[6]; double **dA_array; magma_int_t **dipiv_array;
    magma_queue_t queue;
    if (magma_init() != MAGMA_SUCCESS) return 2;
    magma_queue_create(0, &queue);
    hA = (double *)malloc(sizeof(double) * ldda * n * batch);
    hipiv = (magma_int_t *)malloc(sizeof(magma_int_t) * n * batch);
    for (i = 0; i < ldda * n * batch; ++i) hA[i] = (double)((i * 2654435761u) % 1000003u) / 1000003.0;
    if (magma_malloc((void **)&dA, sizeof(double) * ldda * n * batch) != MAGMA_SUCCESS) return 3;
    magma_malloc((void **)&dipiv, sizeof(magma_int_t) * n * batch);
    magma_malloc((void **)&dinfo, sizeof(magma_int_t) * batch);
    magma_malloc((void **)&dA_array, sizeof(double *) * batch);
    magma_malloc((void **)&dipiv_array, sizeof(magma_int_t *) * batch);
    magma_dsetmatrix(n, n * batch, hA, ldda, dA, ldda, queue);
    magma_dset_pointer(dA_array, dA, ldda, 0, 0, ldda * n, batch, queue);
    magma_iset_pointer(dipiv_array, dipiv, 1, 0, 0, n, batch, queue);
    if (magma_dgetrf_batched(n, n, dA_array, ldda, dipiv_array, dinfo, batch, queue) != 0) return 4;
    magma_queue_sync(queue);
    magma_dgetmatrix(n, n * batch, dA, ldda, hA, ldda, queue);
    magma_getvector(n * batch, sizeof(magma_int_t), dipiv, 1, hipiv, 1, queue);
    magma_getvector(batch, sizeof(magma_int_t), dinfo, 1, hinfo, 1, queue);
    for (b = 0; b < batch; ++b) {
        if (hinfo[b] != 0) return 5;
        for (i = 0; i < n; ++i) {
            magma_int_t p = hipiv[b * n + i];
            if (p < i + 1 || p > n) return 6;
            printf("%d ", (int)p);
        }
        printf("\n");
    }
    for (i = 0; i < ldda * n * batch; ++i) printf("%.17g\n", hA[i]);
    magma_free(dA); magma_free(dipiv); magma_free(dinfo); magma_free(dA_array); magma_free(dipiv_array);
    magma_queue_destroy(queue);
    magma_finalize();
    return 0;
}
'''


def test_c_program_links_and_runs(tmp_path, lib):
    src = tmp_path / "tester.c"
    src.write_text(C_PROG)
    exe = tmp_path / "tester"
    libdir = os.path.join(ROOT, "magma_b200", "lib")
    cmd = ["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-lmagma_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-1000:])
    lines = r.stdout.strip().splitlines()
    n, batch = 48, 6
    ipiv = np.array([[int(t) for t in ln.split()] for ln in lines[:batch]], dtype=np.int32)
    LU = np.array([float(t) for t in lines[batch:]]).reshape(batch, n, n)
    idx = np.arange(n * n * batch, dtype=np.uint64)
    A0 = (((idx * np.uint64(2654435761)) % np.uint64(2 ** 32)) % np.uint64(1000003)).astype(np.float64) / 1000003.0
    A0 = A0.reshape(batch, n, n)
    ref = A0.copy()
    ipr, infr = oracle.getrf_batched(ref, n)
    assert np.array_equal(ipiv, ipr)
    assert np.array_equal(LU, ref)


# ---- single-launch inverse (getri.cu) ---------------------------------------------------------------------------------

@pytest.mark.parametrize("n,ldda,lddia", [(1, 1, 1), (2, 2, 2), (5, 5, 8), (8, 8, 8), (9, 12, 9), (16, 16, 16), (24, 25, 24), (31, 31, 33),
                                          (32, 32, 32), (33, 34, 33), (40, 40, 44), (47, 48, 47), (48, 48, 48), (56, 56, 56),
                                          (63, 63, 64), (64, 64, 64), (64, 70, 66)])
def test_getri_fused(gpu_queue, n, ldda, lddia):
    """magma_dgetri_outofplace_batched for n <= 64 (one launch, DMMA block updates): bit-identical to the oracle and to the
    identity + getrs path it replaces; padding of A and of inv(A) untouched."""
    import torch
    batch = 11
    A0, _ = oracle.random_batch(batch, n, n)
    db = mb.DeviceBatch(batch, n, n, ldda=ldda, queue=gpu_queue)
    Ain = np.full((batch, n, ldda), 3.5)
    Ain[:, :, :n] = A0
    db.upload(Ain)
    assert db.getrf() == 0
    LU, ipiv, info = db.download()
    assert not info.any()
    out = {}
    for mode in (1, 0):
        mb.set_getri_fused(mode)
        try:
            X = torch.full((batch, n, lddia), -2.25, dtype=torch.float64, device="cuda")
            base = X.data_ptr()
            ptrs = torch.tensor([base + b * n * lddia * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
            assert mb.magma_dgetri_outofplace_batched(n, db.dA_array, db.ldda, db.dipiv_array, ptrs, lddia, db.info, batch,
                                                      gpu_queue) == 0
            gpu_queue.sync()
            out[mode] = X.cpu().numpy()
        finally:
            mb.set_getri_fused(1)
    Xr = oracle.getri_outofplace_batched(np.ascontiguousarray(LU[:, :, :n]), ipiv, n)
    assert np.array_equal(out[1][:, :, :n], Xr), f"max diff {np.max(np.abs(out[1][:, :, :n] - Xr))}"
    assert np.array_equal(out[0], out[1])
    assert np.all(out[1][:, :, n:] == -2.25)
    LU2, _, _ = db.download()
    assert np.array_equal(LU2, LU)  # the factors are read only


def test_getri_fused_identity_and_permutation(gpu_queue):
    """Structured factors: identity, a pure permutation, unit-diagonal triangular."""
    n = 64
    rng = np.random.default_rng(5)
    P = np.eye(n)[rng.permutation(n)]
    Lo = np.tril(rng.random((n, n)), -1) * 0.01 + np.eye(n)
    mats = np.stack([np.eye(n), P, Lo, Lo.T, 3.0 * np.eye(n)])
    batch = len(mats)
    db = mb.DeviceBatch(batch, n, n, nrhs=n, queue=gpu_queue)
    db.upload(mats, np.zeros((batch, n, n)))
    assert db.getrf() == 0
    assert mb.magma_dgetri_outofplace_batched(n, db.dA_array, db.ldda, db.dipiv_array, db.dB_array, db.lddb, db.info, batch,
                                              gpu_queue) == 0
    LU, ipiv, info, X = db.download()
    Xr = oracle.getri_outofplace_batched(LU, ipiv, n)
    assert np.array_equal(X, Xr)
    assert np.array_equal(X[0], np.eye(n)) and np.array_equal(X[1].T @ P.T, np.eye(n))


@pytest.mark.parametrize("rows", [2, 3, 5, 6, 9])
def test_register_tier_layout_switches_n32(gpu_queue, rows):
    """The alternative n = 32 layouts behind magma_b200_set_small_rows (two rows per lane generic / staged square /
    shuffle-broadcast square with 16 lanes per matrix / generic only): slower than the default, same bits."""
    mb.set_small_rows(rows)
    try:
        A0, _ = oracle.random_batch(41, 32, 32)
        check_against_oracle(gpu_queue, A0, 32)
    finally:
        mb.set_small_rows(0)


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("m,n,batch", [(128, 128, 7), (100, 100, 5), (97, 97, 5), (128, 40, 5), (120, 200, 4), (127, 96, 4), (96, 96, 5),
                                       (64, 64, 6), (65, 90, 4), (50, 45, 5), (45, 128, 4), (128, 127, 3), (101, 67, 4)])
def test_fused_tail_switch(gpu_queue, level, m, n, batch):
    """Left-looking driver, at most 128 rows: panels factored in the tail of the slab kernel (off by default: slower),
    every switch level gives the oracle's bits (square, tall, wide with a narrow last panel, odd sizes: plain-store path)."""
    mb.set_fused_tail(level)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_fused_tail(0)


@pytest.mark.parametrize("level", [1, 2])
def test_fused_tail_structured_and_vbatched(gpu_queue, level):
    import torch
    mb.set_fused_tail(level)
    try:
        rng = np.random.default_rng(19)
        n = 128
        mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)), rng.integers(-3, 4, size=(n, n)).astype(float)]
        Z = rng.random((n, n)); Z[:, 40] = 0.0; Z[:, 100] = 0.0; mats.append(Z)
        Z2 = rng.random((n, n)); Z2[n // 2:, :] = Z2[:n - n // 2, :]; mats.append(Z2)
        check_against_oracle(gpu_queue, np.stack(mats), n)
        # variable sizes inside the <= 128-row classes: members shorter and narrower than the class maximum
        batch = 40
        ms = rng.integers(33, 129, size=batch).astype(np.int32)
        ns = rng.integers(33, 129, size=batch).astype(np.int32)
        ms[0], ns[0] = 128, 128
        lds = ms.copy()
        mats = [oracle.random_batch(1, int(ms[b]), int(ns[b]))[0][0] for b in range(batch)]  # (n, ld = m)
        offs = np.cumsum([0] + [x.size for x in mats])
        flat = torch.from_numpy(np.concatenate([x.reshape(-1) for x in mats])).cuda()
        aptr = torch.tensor([flat.data_ptr() + int(offs[b]) * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
        kmax = int(np.minimum(ms, ns).max())
        ip = torch.zeros((batch, kmax), dtype=torch.int32, device="cuda")
        ipp = torch.tensor([ip.data_ptr() + b * kmax * 4 for b in range(batch)], dtype=torch.int64, device="cuda")
        info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        dm, dn, dl = (torch.from_numpy(x).cuda() for x in (ms, ns, lds))
        assert mb.magma_dgetrf_vbatched(dm, dn, aptr, dl, ipp, info, batch, gpu_queue) == 0
        gpu_queue.sync()
        out, iph = flat.cpu().numpy(), ip.cpu().numpy()
        for b in range(batch):
            m, nn = int(ms[b]), int(ns[b])
            ref = mats[b].copy().reshape(1, nn, m)
            ipr, infr = oracle.getrf_batched(ref, m)
            assert np.array_equal(iph[b, :min(m, nn)], ipr[0]), (b, m, nn)
            assert np.array_equal(out[offs[b]:offs[b + 1]].reshape(1, nn, m), ref), (b, m, nn)
    finally:
        mb.set_fused_tail(0)


@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("m,n,ld", [(101, 77, 101), (127, 127, 127), (128, 128, 129), (255, 255, 257), (300, 64, 301), (511, 200, 511),
                                    (96, 96, 97), (200, 333, 201), (65, 65, 65), (128, 128, 128), (256, 256, 256)])
def test_any_alignment_tma_staging(gpu_queue, shift, m, n, ld):
    """Odd m / ldda and matrices that start 8 bytes off a 16-byte boundary: the left-looking driver stages them by TMA through
    parity-shifted shared-memory columns (left_update_kernel<NW, true>). Bit-exact, padding and the bytes around each matrix
    untouched."""
    import torch
    batch = 5
    A0, _ = oracle.random_batch(batch, m, n)
    per = n * ld + 3  # an odd gap between matrices: their bases alternate between 0 and 8 mod 16
    buf = torch.full((batch * per + 2,), -1.5, dtype=torch.float64, device="cuda")
    host = np.full(batch * per + 2, -1.5)
    for b in range(batch):
        o = shift + b * per
        blk = np.full((n, ld), 9.75)
        blk[:, :m] = A0[b]
        host[o:o + n * ld] = blk.reshape(-1)
    buf.copy_(torch.from_numpy(host))
    ptrs = torch.tensor([buf.data_ptr() + (shift + b * per) * 8 for b in range(batch)], dtype=torch.int64, device="cuda")
    k = min(m, n)
    ip = torch.zeros((batch, k), dtype=torch.int32, device="cuda")
    ipp = torch.tensor([ip.data_ptr() + b * k * 4 for b in range(batch)], dtype=torch.int64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    assert mb.magma_dgetrf_batched(m, n, ptrs, ld, ipp, info, batch, gpu_queue) == 0
    gpu_queue.sync()
    out = buf.cpu().numpy()
    ref = A0.copy()
    ipr, infr = oracle.getrf_batched(ref, m)
    assert np.array_equal(ip.cpu().numpy(), ipr) and np.array_equal(info.cpu().numpy(), infr)
    want = host.copy()
    for b in range(batch):
        o = shift + b * per
        blk = want[o:o + n * ld].reshape(n, ld)
        blk[:, :m] = ref[b]
    assert np.array_equal(out, want)


@pytest.mark.parametrize("on", [0, 1])
@pytest.mark.parametrize("m,n,batch", [(512, 512, 2), (384, 384, 2), (256, 256, 3), (200, 130, 3), (130, 200, 3), (129, 129, 4), (300, 40, 3),
                                       (511, 77, 2), (160, 160, 3), (450, 300, 2), (257, 17, 3), (333, 333, 2)])
def test_tall_panel_switch(gpu_queue, on, m, n, batch):
    """Panels of 129..512 rows in the left-looking driver: two 16-column halves per thread (default) and the 32-column register
    panel give the oracle's bits (square, tall, wide, odd, a narrow last panel of fewer than 16 columns)."""
    mb.set_tall_panel(on)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_tall_panel(1)


def test_tall_panel_structured(gpu_queue):
    """Singular columns in both halves of a tall panel, ties, identity / permutation inputs (zero-pivot steps are skipped by
    the solve and by the rank-16 update exactly as the oracle skips them)."""
    rng = np.random.default_rng(23)
    n = 256
    mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)), rng.integers(-3, 4, size=(n, n)).astype(float)]
    Z = rng.random((n, n)); Z[:, 5] = 0.0; Z[:, 20] = 0.0; Z[:, 100] = 0.0; Z[:, 117] = 0.0; mats.append(Z)
    Z2 = rng.random((n, n)); Z2[n // 2:, :] = Z2[:n - n // 2, :]; mats.append(Z2)
    Z3 = rng.random((n, n)); Z3[:, :3] = 0.0; mats.append(Z3)
    check_against_oracle(gpu_queue, np.stack(mats), n)


@pytest.mark.parametrize("m,n", [(33, 5), (64, 16), (128, 64), (200, 17), (257, 40), (512, 16)])
@pytest.mark.parametrize("which", ["lower_unit", "upper_nonunit"])
def test_dtrsm_left_on_the_tensor_pipe(gpu_queue, which, m, n):
    """The two solves of an LU through magmablas_dtrsm_batched (side = Left): one sweep of the getrs DMMA kernel, bit-identical
    to the shared-memory DFMA solver (tier 4 forces it) and close to numpy; alpha applied on the way in."""
    import torch
    rng = np.random.default_rng(m + n)
    batch = 3
    T = rng.random((batch, m, m)) + 4 * np.eye(m)   # stored column-major: T[b].T is the matrix
    B = rng.random((batch, n, m + 2))               # lddb = m + 2
    uplo, diag = (mb.MagmaLower, mb.MagmaUnit) if which == "lower_unit" else (mb.MagmaUpper, mb.MagmaNonUnit)
    dT = torch.from_numpy(T).cuda()
    pT = torch.tensor([dT.data_ptr() + 8 * m * m * b for b in range(batch)], dtype=torch.int64, device="cuda")
    out = {}
    for alpha in (1.0, -0.5):
        for tier in (0, 4):
            mb.set_tier(tier)
            try:
                dB = torch.from_numpy(B).cuda()
                pB = torch.tensor([dB.data_ptr() + 8 * (m + 2) * n * b for b in range(batch)], dtype=torch.int64, device="cuda")
                mb.magmablas_dtrsm_batched(mb.MagmaLeft, uplo, mb.MagmaNoTrans, diag, m, n, alpha, pT, m, pB, m + 2, batch, gpu_queue)
                gpu_queue.sync()
                out[(alpha, tier)] = dB.cpu().numpy()
            finally:
                mb.set_tier(0)
        X = out[(alpha, 0)]
        assert np.array_equal(X, out[(alpha, 4)])
        assert np.array_equal(X[:, :, m:], B[:, :, m:])  # padding rows untouched
        for b in range(batch):
            Tm = T[b].T
            Tm = (np.tril(Tm, -1) + np.eye(m)) if which == "lower_unit" else np.triu(Tm)
            ref = np.linalg.solve(Tm, alpha * B[b, :, :m].T)
            assert np.allclose(X[b, :, :m].T, ref, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("parts", [1, 2, 3, 4])
@pytest.mark.parametrize("m,n,batch", [(128, 128, 9000), (300, 300, 2600), (64, 64, 5000)])
def test_batch_split_streams(gpu_queue, parts, m, n, batch):
    """The left-looking chain run as 1..4 slices of the batch on separate streams (forked from / joined to the queue's stream):
    identical bits on a bit-exact sample of every slice, residual on all."""
    mb.set_split(parts)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        db = mb.DeviceBatch(batch, m, n, queue=gpu_queue)
        db.upload(A0)
        assert db.getrf() == 0
        LU, ipiv, info = db.download()
        assert not info.any()
        sel = np.unique(np.concatenate([np.arange(0, batch, max(1, batch // 40)), np.arange(batch - 8, batch),
                                        np.arange(batch // 2 - 4, batch // 2 + 4), np.arange(batch // 3 - 2, batch // 3 + 2)]))
        ref = np.ascontiguousarray(A0[sel])
        ipr, _ = oracle.getrf_batched(ref, m)
        assert np.array_equal(LU[sel], ref) and np.array_equal(ipiv[sel], ipr)
        assert oracle.lu_backward_error(A0, LU, ipiv, m) < oracle.TOL
    finally:
        mb.set_split(0)
