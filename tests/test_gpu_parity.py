"""GPU parity tests (run with -m gpu on the B200 box). Every call goes through the C ABI of
libmagma_b200.so; oracle/ is only the checker. Bars:
  * pivots and info identical to the oracle (and to the LAPACK golden vectors),
  * factors / solutions of every DFMA kernel BIT-IDENTICAL to the oracle's canonical order,
  * the reference testers' checks: ||PA-LU||_F/(n||A||_F) < 30 eps, residual < 30 eps
    (testing/testing_zgetrf_batched.cpp:312-321, testing/testing_zgesv_batched.cpp:133-153).
"""
import os

import numpy as np
import pytest

import oracle
from magma_b200 import batched as mb

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lu_golden.npz"))


def run_getrf(q, A0, m, ldda=None):
    batch, n, _ = A0.shape
    ldda = m if ldda is None else ldda
    db = mb.DeviceBatch(batch, m, n, ldda=ldda, queue=q)
    Ain = np.zeros((batch, n, ldda))
    Ain[:, :, :m] = A0[:, :, :m]
    pad = 7.25  # rows >= m of each column must never be touched
    Ain[:, :, m:] = pad
    db.upload(Ain)
    rc = db.getrf()
    assert rc == 0
    LU, ipiv, info = db.download()
    assert np.all(LU[:, :, m:] == pad), "wrote into the ldda padding"
    return LU[:, :, :m].copy(), ipiv, info


def check_against_oracle(q, A0, m, ldda=None, exact=True):
    LU, ipiv, info = run_getrf(q, A0, m, ldda)
    ref = np.ascontiguousarray(A0[:, :, :m]).copy()
    ipiv_ref, info_ref = oracle.getrf_batched(ref, m)
    assert np.array_equal(info, info_ref)
    assert np.array_equal(ipiv, ipiv_ref), "pivot vectors differ from the oracle"
    if exact:
        assert np.array_equal(LU, ref), f"factors not bit-identical: max diff {np.max(np.abs(LU - ref))}"
    if not info_ref.any():
        err = oracle.lu_backward_error(np.ascontiguousarray(A0[:, :, :m]), LU, ipiv, m)
        assert err < oracle.TOL, f"backward error {err / oracle.EPS:.2f} eps"
    return LU, ipiv, info


# ---- register tier (m, n <= 32) ---------------------------------------------------------------

@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 20, 23, 24, 28, 31, 32])
def test_getrf_small_square(gpu_queue, n):
    A0, _ = oracle.random_batch(1003, n, n)  # odd batch: partially filled warps / CTAs
    check_against_oracle(gpu_queue, A0, n)


@pytest.mark.parametrize("m,n", [(32, 8), (8, 32), (20, 13), (13, 20), (1, 32), (32, 1), (16, 9), (5, 16)])
def test_getrf_small_rectangular(gpu_queue, m, n):
    A0, _ = oracle.random_batch(257, m, n)
    check_against_oracle(gpu_queue, A0, m)


@pytest.mark.parametrize("n,ldda", [(16, 32), (32, 64), (32, 33), (15, 17), (8, 9)])
def test_getrf_small_ldda(gpu_queue, n, ldda):
    A0, _ = oracle.random_batch(300, n, n)
    check_against_oracle(gpu_queue, A0, n, ldda=ldda)


def test_golden_vectors_on_gpu(gpu_queue):
    for key in sorted({k.rsplit("_", 1)[0] for k in GOLD.files if k.endswith("_A")}):
        m, n, r = (int(t[1:]) for t in key.split("_"))
        A0 = GOLD[key + "_A"]
        LU, ipiv, info = run_getrf(gpu_queue, A0, m)
        assert np.array_equal(ipiv, GOLD[key + "_ipiv"]), key
        assert np.array_equal(info, GOLD[key + "_info"]), key
        ref = GOLD[key + "_LU"]
        assert np.max(np.abs(LU - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), key


def test_singular_ties_and_structure(gpu_queue):
    n = 16
    mats = []
    mats.append(np.zeros((n, n)))                      # all zero: info=1, identity pivots
    mats.append(np.ones((n, n)))                       # every column ties: first row wins
    mats.append(np.eye(n))                             # no interchange needed
    mats.append(np.fliplr(np.eye(n)))                  # permutation matrix
    rng = np.random.default_rng(0)
    M = rng.integers(-3, 4, size=(n, n)).astype(float)  # many exact |x| ties
    mats.append(M)
    Z = rng.random((n, n))
    Z[:, 5] = 0.0                                      # zero column (column-major: A[j] is column j)
    mats.append(Z)
    D = rng.random((n, n))
    D[3, :] = D[9, :]                                  # duplicate "rows" of the stored layout
    mats.append(D)
    S = rng.random((n, n))
    S[:, :] = np.where(rng.random((n, n)) < 0.5, -S, S)
    mats.append(S)
    A0 = np.stack(mats)
    check_against_oracle(gpu_queue, A0, n)
    # the same structures in the blocked tier (n = 40) and as a 32x32
    for nn in (32, 40):
        big = np.zeros((len(mats), nn, nn))
        big[:, :n, :n] = A0
        big[:, n:, n:] = np.eye(nn - n)
        check_against_oracle(gpu_queue, big, nn)


# ---- blocked tier -----------------------------------------------------------------------------

@pytest.mark.parametrize("n,batch", [(33, 65), (48, 40), (64, 33), (100, 20), (128, 24), (200, 6), (256, 5),
                                     (512, 3)])
def test_getrf_blocked_square(gpu_queue, n, batch):
    A0, _ = oracle.random_batch(batch, n, n)
    check_against_oracle(gpu_queue, A0, n)


@pytest.mark.parametrize("m,n,batch", [(300, 70, 5), (70, 300, 5), (600, 40, 3), (40, 600, 3), (1100, 24, 2),
                                       (33, 32, 9), (32, 33, 9), (97, 97, 7), (129, 65, 4)])
def test_getrf_blocked_rectangular(gpu_queue, m, n, batch):
    A0, _ = oracle.random_batch(batch, m, n)
    check_against_oracle(gpu_queue, A0, m)


@pytest.mark.parametrize("n,ldda", [(100, 128), (100, 101), (65, 67)])
def test_getrf_blocked_ldda(gpu_queue, n, ldda):
    A0, _ = oracle.random_batch(6, n, n)
    check_against_oracle(gpu_queue, A0, n, ldda=ldda)


# ---- register-file tier (lu_mid.cu), forced up to its full range ------------------------------------

@pytest.fixture
def mid_tier_128():
    mb.set_mid_max(128)
    mb.set_small_rows(7)   # keep the register-file tier on 65..96 too (default: left-looking blocked driver there)
    yield
    mb.set_small_rows(0)
    mb.set_mid_max(128)


@pytest.mark.parametrize("m,n,batch", [(33, 33, 21), (64, 64, 9), (65, 65, 9), (96, 96, 7), (100, 100, 5), (128, 128, 6),
                                       (128, 40, 5), (40, 128, 5), (97, 90, 4), (33, 128, 3), (128, 33, 3), (127, 121, 3),
                                       (1, 100, 3), (100, 1, 3)])
def test_getrf_mid_tier(gpu_queue, mid_tier_128, m, n, batch):
    A0, _ = oracle.random_batch(batch, m, n)
    check_against_oracle(gpu_queue, A0, m)


@pytest.mark.parametrize("n,ldda", [(100, 128), (100, 101), (65, 67), (128, 130)])
def test_getrf_mid_tier_ldda(gpu_queue, mid_tier_128, n, ldda):
    A0, _ = oracle.random_batch(6, n, n)
    check_against_oracle(gpu_queue, A0, n, ldda=ldda)


def test_mid_tier_singular_and_ties(gpu_queue, mid_tier_128):
    rng = np.random.default_rng(5)
    for n in (40, 72, 128):
        mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)),
                rng.integers(-3, 4, size=(n, n)).astype(float)]
        Z = rng.random((n, n))
        Z[:, n // 3] = 0.0          # an exactly zero column of the stored layout
        mats.append(Z)
        Z2 = rng.random((n, n))
        Z2[n // 2:, :] = Z2[:n - n // 2, :]   # rank deficient: zero pivots late in the factorisation
        mats.append(Z2)
        check_against_oracle(gpu_queue, np.stack(mats), n)


def test_blocked_tier_on_small_sizes(gpu_queue):
    """Force the blocked kernels onto sizes the register tier normally takes."""
    mb.set_tier(2)
    try:
        for n in (1, 5, 16, 31, 32):
            A0, _ = oracle.random_batch(37, n, n)
            check_against_oracle(gpu_queue, A0, n)
    finally:
        mb.set_tier(0)


@pytest.mark.parametrize("m,n,batch", [(400, 400, 3), (512, 512, 2), (300, 70, 4), (300, 500, 3), (257, 300, 3),
                                       (500, 90, 3), (390, 33, 3)])
def test_left_looking_16_warps(gpu_queue, m, n, batch):
    """Tier 7: the left-looking slab driver with 16 warps per slab (257..512 rows; not the default there)."""
    mb.set_tier(7)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_tier(0)


@pytest.mark.parametrize("m,n,batch", [(33, 33, 21), (64, 64, 9), (65, 65, 9), (80, 80, 5), (96, 96, 7), (100, 100, 5),
                                       (128, 128, 6), (128, 40, 5), (40, 128, 5), (97, 90, 4), (33, 128, 3), (128, 33, 3),
                                       (127, 121, 3), (70, 200, 3), (100, 37, 4)])
def test_left_looking_4_warps(gpu_queue, m, n, batch):
    """The left-looking driver with 4 warps per slab (at most 128 rows): default for 65..96, forced here for 33..128."""
    mb.set_mid_max(32)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_mid_max(128)


@pytest.mark.parametrize("m,n,batch", [(200, 200, 4), (256, 256, 3), (70, 300, 4), (150, 40, 5)])
def test_right_looking_forced(gpu_queue, m, n, batch):
    """Tier 6: the right-looking flow on shapes the left-looking driver takes by default."""
    mb.set_tier(6)
    try:
        A0, _ = oracle.random_batch(batch, m, n)
        check_against_oracle(gpu_queue, A0, m)
    finally:
        mb.set_tier(0)


def test_left_looking_singular_and_ties(gpu_queue):
    """Zero columns, duplicate rows and integer ties through the left-looking driver (n = 160)."""
    rng = np.random.default_rng(11)
    n = 160
    mats = [rng.integers(-2, 3, size=(n, n)).astype(float)]
    Z = rng.random((n, n))
    Z[:, 40] = 0.0
    Z[:, 41] = 0.0
    mats.append(Z)
    Z2 = rng.random((n, n))
    Z2[n // 2:, :] = Z2[:n - n // 2, :]
    mats.append(Z2)
    mats.append(np.zeros((n, n)))
    check_against_oracle(gpu_queue, np.stack(mats), n)


# ---- solves -----------------------------------------------------------------------------------

def run_gesv(q, A0, B0, n, ldda=None, lddb=None):
    batch = A0.shape[0]
    nrhs = B0.shape[1]
    ldda = n if ldda is None else ldda
    lddb = n if lddb is None else lddb
    db = mb.DeviceBatch(batch, n, n, ldda=ldda, nrhs=nrhs, lddb=lddb, queue=q)
    Ain = np.zeros((batch, n, ldda))
    Ain[:, :, :n] = A0
    Bin = np.zeros((batch, nrhs, lddb))
    Bin[:, :, :n] = B0
    db.upload(Ain, Bin)
    assert db.gesv() == 0
    LU, ipiv, info, X = db.download()
    return LU[:, :, :n].copy(), ipiv, info, X[:, :, :n].copy()


@pytest.mark.parametrize("n", [1, 2, 4, 7, 8, 13, 16, 17, 24, 31, 32])
def test_gesv_small_fused(gpu_queue, n):
    A0, seed = oracle.random_batch(515, n, n)
    B0, _ = oracle.random_batch(515, n, 1, iseed=seed)
    LU, ipiv, info, X = run_gesv(gpu_queue, A0, B0, n)
    Ar, Br = A0.copy(), B0.copy()
    ipiv_ref, info_ref = oracle.gesv_batched(Ar, Br, n)
    assert np.array_equal(ipiv, ipiv_ref) and np.array_equal(info, info_ref)
    assert np.array_equal(LU, Ar)
    assert np.array_equal(X, Br), f"solution not bit-identical, max diff {np.max(np.abs(X - Br))}"
    assert oracle.solve_residual(oracle.MagmaNoTrans, A0, X, B0, n) < oracle.TOL


def test_gesv_small_ldda_lddb(gpu_queue):
    n = 16
    A0, seed = oracle.random_batch(100, n, n)
    B0, _ = oracle.random_batch(100, n, 1, iseed=seed)
    LU, ipiv, info, X = run_gesv(gpu_queue, A0, B0, n, ldda=32, lddb=32)
    Ar, Br = A0.copy(), B0.copy()
    oracle.gesv_batched(Ar, Br, n)
    assert np.array_equal(X, Br)


@pytest.mark.parametrize("n,nrhs,batch", [(16, 3, 50), (32, 16, 20), (40, 1, 30), (100, 5, 10), (128, 16, 6),
                                          (512, 16, 2), (200, 33, 3)])
def test_gesv_getrf_getrs_path(gpu_queue, n, nrhs, batch):
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    LU, ipiv, info, X = run_gesv(gpu_queue, A0, B0, n)
    Ar, Br = A0.copy(), B0.copy()
    ipiv_ref, info_ref = oracle.gesv_batched(Ar, Br, n)
    assert np.array_equal(ipiv, ipiv_ref) and np.array_equal(info, info_ref)
    assert np.array_equal(LU, Ar)
    assert np.array_equal(X, Br), f"solution not bit-identical, max diff {np.max(np.abs(X - Br))}"
    assert oracle.solve_residual(oracle.MagmaNoTrans, A0, X, B0, n) < oracle.TOL


@pytest.mark.parametrize("trans", [mb.MagmaNoTrans, mb.MagmaTrans, mb.MagmaConjTrans])
@pytest.mark.parametrize("n,nrhs", [(7, 2), (32, 1), (75, 4), (160, 16)])
def test_getrs(gpu_queue, trans, n, nrhs):
    batch = 9
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    db.upload(A0, B0)
    assert db.getrf() == 0
    assert db.getrs(trans) == 0
    LU, ipiv, info, X = db.download()
    Xr = B0.copy()
    oracle.getrs_batched(trans, LU, ipiv, Xr, n)
    assert np.array_equal(X, Xr), f"max diff {np.max(np.abs(X - Xr))}"
    assert oracle.solve_residual(trans, A0, X, B0, n) < oracle.TOL


@pytest.mark.parametrize("n,batch", [(1, 5), (7, 9), (16, 33), (32, 12), (33, 7), (75, 5), (128, 4), (200, 3), (300, 2)])
def test_getri_outofplace(gpu_queue, n, batch):
    """magma_dgetri_outofplace_batched: bit-identical to the oracle's solve on the identity,
    and the reference tester's check ||I - A inv(A)|| small (testing/testing_zgetri_batched.cpp)."""
    A0, _ = oracle.random_batch(batch, n, n)
    db = mb.DeviceBatch(batch, n, n, nrhs=n, queue=gpu_queue)
    db.upload(A0, np.full((batch, n, n), 7.0))  # stale contents must be overwritten
    assert db.getrf() == 0
    assert mb.magma_dgetri_outofplace_batched(n, db.dA_array, db.ldda, db.dipiv_array, db.dB_array, db.lddb,
                                              db.info, batch, gpu_queue) == 0
    LU, ipiv, info, X = db.download()
    assert not info.any()
    Xr = oracle.getri_outofplace_batched(LU, ipiv, n)
    assert np.array_equal(X, Xr), f"max diff {np.max(np.abs(X - Xr))}"
    # stored layout is [b][col][row]: A = A0[b].T, inv = X[b].T
    for b in range(batch):
        A, Ai = A0[b].T, X[b].T
        r = np.linalg.norm(np.eye(n) - A @ Ai, 1) / (n * np.linalg.norm(A, 1) * np.linalg.norm(Ai, 1))
        assert r < oracle.TOL


def test_getri_argument_errors(gpu_queue):
    assert mb.magma_dgetri_outofplace_batched(-1, None, 1, None, None, 1, None, 1, gpu_queue) == -1
    assert mb.magma_dgetri_outofplace_batched(4, None, 3, None, None, 4, None, 1, gpu_queue) == -3
    assert mb.magma_dgetri_outofplace_batched(4, None, 4, None, None, 3, None, 1, gpu_queue) == -6
    assert mb.magma_dgetri_outofplace_batched(0, None, 1, None, None, 1, None, 1, gpu_queue) == 0


def _dominant_batch(batch, m, n, seed=3):
    rng = np.random.default_rng(seed)
    A = rng.random((batch, n, m)) - 0.5
    for k in range(min(m, n)):
        A[:, k, k] += float(max(m, n))
    return A


@pytest.mark.parametrize("m,n,batch", [(1, 1, 5), (5, 5, 9), (16, 16, 33), (32, 32, 9), (20, 9, 7), (9, 20, 7), (33, 33, 6),
                                       (64, 64, 5), (100, 100, 4), (128, 128, 3), (200, 200, 3), (300, 70, 3), (70, 300, 3),
                                       (512, 512, 2), (257, 300, 2)])
def test_getrf_nopiv(gpu_queue, m, n, batch):
    """magma_dgetrf_nopiv_batched: bit-identical to oracle_dgetf2_nopiv."""
    A0 = _dominant_batch(batch, m, n)
    A0[0] = oracle.random_batch(1, m, n)[0][0]  # one general matrix: no pivoting is still well defined, just less stable
    db = mb.DeviceBatch(batch, m, n, queue=gpu_queue)
    db.upload(A0)
    db.info.fill_(-3)
    assert mb.magma_dgetrf_nopiv_batched(m, n, db.dA_array, db.ldda, db.info, batch, gpu_queue) == 0
    LU, _, info = db.download()
    ref = A0.copy()
    info_ref = oracle.getrf_nopiv_batched(ref, m)
    assert np.array_equal(info, info_ref)
    assert np.array_equal(LU, ref), f"max diff {np.max(np.abs(LU - ref))}"


def test_getrf_nopiv_zero_diagonal(gpu_queue):
    n, batch = 40, 4
    A0 = _dominant_batch(batch, n, n)
    A0[1, 7, :] = 0.0      # zero column: A(7,7) is zero at its turn
    A0[2, 0, 0] = 0.0      # zero first pivot
    A0[3, 35, :] = 0.0     # zero column in the second panel
    db = mb.DeviceBatch(batch, n, n, queue=gpu_queue)
    db.upload(A0)
    assert mb.magma_dgetrf_nopiv_batched(n, n, db.dA_array, db.ldda, db.info, batch, gpu_queue) == 0
    LU, _, info = db.download()
    ref = A0.copy()
    info_ref = oracle.getrf_nopiv_batched(ref, n)
    assert list(info_ref) == [0, 8, 1, 36]
    assert np.array_equal(info, info_ref)
    assert np.array_equal(LU, ref, equal_nan=True)


@pytest.mark.parametrize("n,nrhs", [(7, 2), (32, 1), (75, 4), (160, 16)])
def test_gesv_and_getrs_nopiv(gpu_queue, n, nrhs):
    batch = 6
    A0 = _dominant_batch(batch, n, n)
    B0 = np.random.default_rng(9).random((batch, nrhs, n))
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    db.upload(A0, B0)
    assert mb.magma_dgesv_nopiv_batched(n, nrhs, db.dA_array, db.ldda, db.dB_array, db.lddb, db.info, batch,
                                        gpu_queue) == 0
    LU, _, info, X = db.download()
    ref = A0.copy()
    assert not oracle.getrf_nopiv_batched(ref, n).any() and not info.any()
    assert np.array_equal(LU, ref)
    Xr = B0.copy()
    oracle.getrs_nopiv_batched(mb.MagmaNoTrans, ref, Xr, n)
    assert np.array_equal(X, Xr)
    assert oracle.solve_residual(mb.MagmaNoTrans, A0, X, B0, n) < oracle.TOL
    # transposed solve from the same factors
    db.B.copy_(db.torch.from_numpy(B0))
    assert mb.magma_dgetrs_nopiv_batched(mb.MagmaTrans, n, nrhs, db.dA_array, db.ldda, db.dB_array, db.lddb, db.info,
                                         batch, gpu_queue) == 0
    Xt = db.download()[3]
    Xtr = B0.copy()
    oracle.getrs_nopiv_batched(mb.MagmaTrans, ref, Xtr, n)
    assert np.array_equal(Xt, Xtr)


def test_nopiv_argument_errors(gpu_queue):
    assert mb.magma_dgetrf_nopiv_batched(-1, 4, None, 4, None, 1, gpu_queue) == -1
    assert mb.magma_dgetrf_nopiv_batched(4, -1, None, 4, None, 1, gpu_queue) == -2
    assert mb.magma_dgetrf_nopiv_batched(4, 4, None, 3, None, 1, gpu_queue) == -4
    assert mb.magma_dgetrs_nopiv_batched(5, 4, 1, None, 4, None, 4, None, 1, gpu_queue) == -1
    assert mb.magma_dgetrs_nopiv_batched(mb.MagmaNoTrans, 4, 1, None, 3, None, 4, None, 1, gpu_queue) == -5
    assert mb.magma_dgesv_nopiv_batched(4, 1, None, 4, None, 3, None, 1, gpu_queue) == -6


# ---- variable-size batch ------------------------------------------------------------------------

def _vbatched_case(q, ms, ns, lds=None, expert=False):
    import torch
    batch = len(ms)
    lds = [max(1, m) for m in ms] if lds is None else lds
    sizes = [ld * n for ld, n in zip(lds, ns)]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    mns = [min(m, n) for m, n in zip(ms, ns)]
    poffs = np.concatenate([[0], np.cumsum([max(1, k) for k in mns])]).astype(np.int64)
    x, _ = oracle.dlarnv(int(offs[-1]) + 1)
    host = x[:int(offs[-1])].copy()
    dev = torch.device("cuda", 0)
    dA = torch.from_numpy(host).to(dev)
    dip = torch.zeros(int(poffs[-1]), dtype=torch.int32, device=dev)
    dinfo = torch.full((batch,), -7, dtype=torch.int32, device=dev)
    pA = torch.tensor([dA.data_ptr() + 8 * int(o) for o in offs[:-1]], dtype=torch.int64, device=dev)
    pP = torch.tensor([dip.data_ptr() + 4 * int(o) for o in poffs[:-1]], dtype=torch.int64, device=dev)
    dm = torch.tensor(ms, dtype=torch.int32, device=dev)
    dn = torch.tensor(ns, dtype=torch.int32, device=dev)
    dl = torch.tensor(lds, dtype=torch.int32, device=dev)
    if expert:
        lwork = np.array([-1], dtype=np.int32)
        rc = mb.magma_dgetrf_vbatched_max_nocheck_work(dm, dn, max(ms), max(ns), max(mns), 0, pA, dl, pP, dinfo, 0,
                                                       lwork, batch, q)
        assert rc == 0 and lwork[0] > 0
        work = torch.zeros(int(lwork[0]), dtype=torch.uint8, device=dev)
        rc = mb.magma_dgetrf_vbatched_max_nocheck_work(dm, dn, max(ms), max(ns), max(mns), 0, pA, dl, pP, dinfo,
                                                       work, lwork, batch, q)
    else:
        rc = mb.magma_dgetrf_vbatched(dm, dn, pA, dl, pP, dinfo, batch, q)
    assert rc == 0
    q.sync()
    torch.cuda.synchronize()
    out = dA.cpu().numpy()
    piv = dip.cpu().numpy()
    info = dinfo.cpu().numpy()
    for b in range(batch):
        m, n, ld = ms[b], ns[b], lds[b]
        if m == 0 or n == 0:
            assert info[b] == 0
            continue
        ref = host[offs[b]:offs[b + 1]].reshape(n, ld).copy()
        rp = np.zeros(max(1, mns[b]), dtype=np.int32)
        rinfo = oracle.lib().oracle_dgetf2(m, n, ref.reshape(-1), ld, rp)
        got = out[offs[b]:offs[b + 1]].reshape(n, ld)
        assert info[b] == rinfo, (b, m, n)
        assert np.array_equal(piv[poffs[b]:poffs[b] + mns[b]], rp[:mns[b]]), (b, m, n)
        assert np.array_equal(got[:, :m], ref[:, :m]), (b, m, n)


def test_vbatched_mixed_sizes(gpu_queue):
    rng = np.random.default_rng(7)
    ns = rng.integers(1, 140, size=97).tolist()
    _vbatched_case(gpu_queue, ns, ns)


def test_vbatched_all_small(gpu_queue):
    rng = np.random.default_rng(8)
    ns = rng.integers(1, 33, size=301).tolist()
    _vbatched_case(gpu_queue, ns, ns)


def test_vbatched_rectangular_empty_and_padded(gpu_queue):
    ms = [5, 0, 40, 33, 70, 16, 1, 64, 90, 12]
    ns = [9, 4, 33, 40, 20, 0, 1, 64, 30, 50]
    lds = [max(1, m) + (3 if i % 2 else 0) for i, m in enumerate(ms)]
    _vbatched_case(gpu_queue, ms, ns, lds)


def test_vbatched_expert_async_entry(gpu_queue):
    rng = np.random.default_rng(9)
    ns = rng.integers(1, 100, size=60).tolist()
    _vbatched_case(gpu_queue, ns, ns, expert=True)


def test_vbatched_config4_shape_sample(gpu_queue):
    """BASELINE config 4 draws n uniformly in 16..512; a small sample of that distribution."""
    rng = np.random.default_rng(10)
    ns = rng.integers(16, 513, size=24).tolist()
    _vbatched_case(gpu_queue, ns, ns)


def test_vbatched_argument_check(gpu_queue, capfd):
    import torch
    dev = torch.device("cuda", 0)
    dm = torch.tensor([4, -1, 3], dtype=torch.int32, device=dev)
    dn = torch.tensor([4, 2, 3], dtype=torch.int32, device=dev)
    dl = torch.tensor([4, 1, 3], dtype=torch.int32, device=dev)
    z = torch.zeros(3, dtype=torch.int64, device=dev)
    info = torch.zeros(3, dtype=torch.int32, device=dev)
    assert mb.magma_dgetrf_vbatched(dm, dn, z, dl, z, info, 3, gpu_queue) == -1
    dm = torch.tensor([4, 2, 3], dtype=torch.int32, device=dev)
    dl = torch.tensor([4, 1, 3], dtype=torch.int32, device=dev)
    assert mb.magma_dgetrf_vbatched(dm, dn, z, dl, z, info, 3, gpu_queue) == -4


# ---- size-independent properties at (near) BASELINE sizes ------------------------------------------

def test_dlarnv_device_stream(gpu_queue):
    import torch
    n = 1_000_003
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, n, d, gpu_queue)
    gpu_queue.sync()
    x, s2 = oracle.dlarnv(n)
    assert np.array_equal(d.cpu().numpy(), x) and np.array_equal(seed, s2)


@pytest.mark.parametrize("n,batch", [(32, 10000), (16, 200000)])
def test_full_size_properties(gpu_queue, n, batch):
    """BASELINE config 1 at full size and config 2 at 1/5 size: residual property on every matrix
    (computed on the device in float64 with torch), bit-exactness on a sample through the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    db = mb.DeviceBatch(batch, n, n, nrhs=1, queue=gpu_queue)
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, batch * n * n, db.A, gpu_queue)
    mb.dlarnv_uniform(seed, batch * n, db.B, gpu_queue)
    gpu_queue.sync()
    A0 = db.A.clone()
    B0 = db.B.clone()
    assert db.gesv() == 0
    gpu_queue.sync()
    torch.cuda.synchronize()
    assert int(db.info.abs().max()) == 0
    # storage [b, j, i] = A_b(i, j)  ->  A_b = A0[b].T
    Am = A0.transpose(1, 2)
    X = db.B.transpose(1, 2)
    R = B0.transpose(1, 2) - Am @ X
    res = R.abs().sum(dim=1).amax(dim=1) / (
        n * Am.abs().sum(dim=2).amax(dim=1) * X.abs().sum(dim=1).amax(dim=1))
    assert float(res.max()) < oracle.TOL
    # pivots are a valid LAPACK interchange sequence
    ip = db.ipiv
    lo = torch.arange(1, n + 1, device=dev, dtype=torch.int32)
    assert bool((ip >= lo).all()) and bool((ip <= n).all())
    # sample vs oracle, bit exact
    idx = np.linspace(0, batch - 1, 64).astype(np.int64)
    As = A0[idx].cpu().numpy()
    Bs = B0[idx].cpu().numpy()
    ipr, _ = oracle.gesv_batched(As, Bs, n)
    assert np.array_equal(db.A[idx].cpu().numpy(), As)
    assert np.array_equal(db.B[idx].cpu().numpy(), Bs)
    assert np.array_equal(db.ipiv[idx].cpu().numpy(), ipr)


def _lu_backward_error_on_device(A0, LU, ipiv, n):
    """max_b ||P A - L U||_F / (n ||A||_F) computed with torch on the device (storage [b, j, i] = A_b(i, j))."""
    import torch
    A = A0.transpose(1, 2).clone()                      # [b, i, j]
    F = LU.transpose(1, 2)
    L = torch.tril(F, -1) + torch.eye(n, dtype=F.dtype, device=F.device)
    U = torch.triu(F)
    idx = torch.arange(A.shape[0], device=A.device)
    for k in range(n):                                   # LAPACK's forward interchanges
        p = (ipiv[:, k] - 1).long()
        rk = A[idx, k, :].clone()
        A[idx, k, :] = A[idx, p, :]
        A[idx, p, :] = rk
    num = torch.linalg.matrix_norm(A - L @ U)
    den = torch.linalg.matrix_norm(A0.transpose(1, 2)) * n
    return float((num / den).max())


@pytest.mark.parametrize("n,batch,nrhs", [(128, 20000, 0), (512, 600, 16)])
def test_full_size_properties_mid_and_blocked(gpu_queue, n, batch, nrhs):
    """BASELINE configs 3 and 5 at (near) full size: the testers' backward-error check on EVERY matrix,
    computed on the device, pivots in LAPACK range, info == 0, solve residual for config 5, and a
    bit-exact sample against the oracle."""
    import torch
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, batch * n * n, db.A, gpu_queue)
    if nrhs:
        mb.dlarnv_uniform(seed, batch * n * nrhs, db.B, gpu_queue)
    gpu_queue.sync()
    A0 = db.A.clone()
    B0 = db.B.clone() if nrhs else None
    assert db.getrf() == 0
    if nrhs:
        assert db.getrs() == 0
    gpu_queue.sync()
    torch.cuda.synchronize()
    assert int(db.info.abs().max()) == 0
    lo = torch.arange(1, n + 1, device=db.ipiv.device, dtype=torch.int32)
    assert bool((db.ipiv >= lo).all()) and bool((db.ipiv <= n).all())
    err = _lu_backward_error_on_device(A0, db.A, db.ipiv, n)
    assert err < oracle.TOL, f"backward error {err / oracle.EPS:.2f} eps"
    if nrhs:
        Am = A0.transpose(1, 2)
        X = db.B.transpose(1, 2)
        R = B0.transpose(1, 2) - Am @ X
        res = R.abs().sum(dim=1).amax(dim=1) / (n * Am.abs().sum(dim=2).amax(dim=1) * X.abs().sum(dim=1).amax(dim=1))
        assert float(res.max()) < oracle.TOL
    idx = np.linspace(0, batch - 1, 6).astype(np.int64)
    As = A0[idx].cpu().numpy()
    ipr, _ = oracle.getrf_batched(As, n)
    assert np.array_equal(db.A[idx].cpu().numpy(), As)
    assert np.array_equal(db.ipiv[idx].cpu().numpy(), ipr)
    if nrhs:
        Bs = B0[idx].cpu().numpy()
        oracle.getrs_batched(mb.MagmaNoTrans, As, ipr, Bs, n)
        assert np.array_equal(db.B[idx].cpu().numpy(), Bs)


@pytest.mark.parametrize("n,nrhs", [(33, 1), (40, 17), (100, 16), (129, 3), (250, 20), (256, 16), (257, 5), (511, 9)])
def test_getrs_tensor_pipe_shapes(gpu_queue, n, nrhs):
    """getrs_dmma_kernel: sizes that are not multiples of 8 / 32, right-hand-side counts around the 16-wide tile."""
    batch = 5
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=gpu_queue)
    db.upload(A0, B0)
    assert db.getrf() == 0
    assert db.getrs() == 0
    LU, ipiv, info, X = db.download()
    Xr = B0.copy()
    oracle.getrs_batched(mb.MagmaNoTrans, LU, ipiv, Xr, n)
    assert np.array_equal(X, Xr), f"max diff {np.max(np.abs(X - Xr))}"
    assert oracle.solve_residual(mb.MagmaNoTrans, A0, X, B0, n) < oracle.TOL


def test_host_front_end(gpu_queue):
    """magma_b200_dgesv_batched_host: pageable host buffers in, results out, chunked pipeline."""
    n, batch = 16, 5000
    A0, seed = oracle.random_batch(batch, n, n)
    B0, _ = oracle.random_batch(batch, n, 1, iseed=seed)
    hA, hB = A0.copy(), B0.copy()
    hip = np.zeros((batch, n), dtype=np.int32)
    hinfo = np.full(batch, -5, dtype=np.int32)
    assert mb.dgesv_batched_host(n, 1, hA, n, hip, hB, n, hinfo, batch, gpu_queue) == 0
    Ar, Br = A0.copy(), B0.copy()
    ipr, infr = oracle.gesv_batched(Ar, Br, n)
    assert np.array_equal(hA, Ar) and np.array_equal(hB, Br)
    assert np.array_equal(hip, ipr) and np.array_equal(hinfo, infr)
    # getrf front end, blocked tier
    n, batch = 96, 50
    A0, _ = oracle.random_batch(batch, n, n)
    hA = A0.copy()
    hip = np.zeros((batch, n), dtype=np.int32)
    hinfo = np.zeros(batch, dtype=np.int32)
    assert mb.dgetrf_batched_host(n, n, hA, n, hip, hinfo, batch, gpu_queue) == 0
    Ar = A0.copy()
    ipr, infr = oracle.getrf_batched(Ar, n)
    assert np.array_equal(hA, Ar) and np.array_equal(hip, ipr)


def test_standalone_blas_entry_points(gpu_queue):
    """magmablas_dtrsm_batched / magma_dgemm_batched_core / magma_dlaswp_rowserial_batched."""
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    batch, m, n = 5, 45, 7
    T = rng.random((batch, m, m)) + 4 * np.eye(m)
    B = rng.random((batch, n, m))
    dT = torch.from_numpy(T).to(dev)
    pT = torch.tensor([dT.data_ptr() + 8 * m * m * b for b in range(batch)], dtype=torch.int64, device=dev)
    for uplo in (mb.MagmaLower, mb.MagmaUpper):
        for trans in (mb.MagmaNoTrans, mb.MagmaTrans):
            for diag in (mb.MagmaNonUnit, mb.MagmaUnit):
                dB = torch.from_numpy(B).to(dev)
                pB = torch.tensor([dB.data_ptr() + 8 * m * n * b for b in range(batch)], dtype=torch.int64,
                                  device=dev)
                mb.magmablas_dtrsm_batched(mb.MagmaLeft, uplo, trans, diag, m, n, 2.0, pT, m, pB, m, batch,
                                           gpu_queue)
                gpu_queue.sync()
                X = dB.cpu().numpy()
                for b in range(batch):
                    Tm = T[b].T  # storage is column-major
                    Tm = np.tril(Tm) if uplo == mb.MagmaLower else np.triu(Tm)
                    if diag == mb.MagmaUnit:
                        Tm = Tm - np.diag(np.diag(Tm)) + np.eye(m)
                    op = Tm if trans == mb.MagmaNoTrans else Tm.T
                    ref = np.linalg.solve(op, 2.0 * B[b].T)
                    assert np.allclose(X[b].T, ref, rtol=1e-10, atol=1e-12)
    # gemm with offsets
    k = 19
    Am = rng.random((batch, k + 2, m + 3))   # lda = m+3, cols k+2
    Bm = rng.random((batch, n + 1, k + 4))
    Cm = rng.random((batch, n, m))
    dA_, dB_, dC_ = (torch.from_numpy(x).to(dev) for x in (Am, Bm, Cm))
    pa = torch.tensor([dA_.data_ptr() + 8 * Am[0].size * b for b in range(batch)], dtype=torch.int64, device=dev)
    pb = torch.tensor([dB_.data_ptr() + 8 * Bm[0].size * b for b in range(batch)], dtype=torch.int64, device=dev)
    pc = torch.tensor([dC_.data_ptr() + 8 * Cm[0].size * b for b in range(batch)], dtype=torch.int64, device=dev)
    mb.magma_dgemm_batched_core(mb.MagmaNoTrans, mb.MagmaNoTrans, m, n, k, -1.0, pa, 3, 2, m + 3, pb, 4, 1, k + 4,
                                1.0, pc, 0, 0, m, batch, gpu_queue)
    gpu_queue.sync()
    got = dC_.cpu().numpy()
    for b in range(batch):
        Ab = Am[b].T[3:3 + m, 2:2 + k]
        Bb = Bm[b].T[4:4 + k, 1:1 + n]
        assert np.allclose(got[b].T, Cm[b].T - Ab @ Bb, rtol=1e-12, atol=1e-12)
    # laswp
    P = np.stack([np.array([rng.integers(i + 1, m + 1) for i in range(m)], dtype=np.int32) for _ in range(batch)])
    dP = torch.from_numpy(P).to(dev)
    pP = torch.tensor([dP.data_ptr() + 4 * m * b for b in range(batch)], dtype=torch.int64, device=dev)
    dB = torch.from_numpy(B).to(dev)
    pB = torch.tensor([dB.data_ptr() + 8 * m * n * b for b in range(batch)], dtype=torch.int64, device=dev)
    mb.magma_dlaswp_rowserial_batched(n, pB, m, 1, m, pP, batch, gpu_queue)
    gpu_queue.sync()
    got = dB.cpu().numpy()
    for b in range(batch):
        ref = B[b].T.copy()
        for i in range(m):
            p = P[b, i] - 1
            ref[[i, p]] = ref[[p, i]]
        assert np.array_equal(got[b].T, ref)


def test_fortran_style_wrappers(gpu_queue, lib):
    import ctypes as C
    n, batch = 12, 10
    A0, _ = oracle.random_batch(batch, n, n)
    db = mb.DeviceBatch(batch, n, n, queue=gpu_queue)
    db.upload(A0)
    ci = lambda v: C.c_int(v)  # noqa: E731
    cp = lambda v: C.c_size_t(v)  # noqa: E731
    m_, n_, ld_, bc_, info_ = ci(n), ci(n), ci(n), ci(batch), ci(-1)
    pa, pp, pi, pq = cp(db.dA_array.data_ptr()), cp(db.dipiv_array.data_ptr()), cp(db.info.data_ptr()), cp(
        gpu_queue.handle)
    lib.magmaf_dgetrf_batched_(C.addressof(m_), C.addressof(n_), C.addressof(pa), C.addressof(ld_), C.addressof(pp),
                               C.addressof(pi), C.addressof(bc_), C.addressof(pq), C.addressof(info_))
    assert info_.value == 0
    LU, ipiv, info = db.download()
    ref = A0.copy()
    ipr, _ = oracle.getrf_batched(ref, n)
    assert np.array_equal(LU, ref) and np.array_equal(ipiv, ipr)
