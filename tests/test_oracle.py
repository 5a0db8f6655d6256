"""CPU tests: pin oracle/lu_oracle.c against (1) the committed LAPACK golden vectors, (2) the
known-answer values recorded in SURVEY.md section 8c, (3) the host LAPACK itself (the reference
testers' CPU path) on fresh dlarnv streams, and check the checker's own error formulas."""
import os

import numpy as np
import pytest

import oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lu_golden.npz"))
SHAPES = sorted({k.rsplit("_", 1)[0] for k in GOLD.files if k.endswith("_A")})


def _shape(key):
    m, n, r = key.split("_")
    return int(m[1:]), int(n[1:]), int(r[1:])


def test_dlarnv_known_answer():
    # dlarnv(1,{0,0,0,1}) -> 0.12062469795087694, 0.64384591082168541
    x, seed = oracle.dlarnv(2)
    assert x[0] == 0.12062469795087694 and x[1] == 0.64384591082168541
    assert np.array_equal(GOLD["first_two"], x)


def test_dlarnv_matches_lapack_stream():
    L = oracle.lapack()
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    y = np.empty(300001)
    L.lapack_dlarnv(1, seed, y.size, y)
    x, s2 = oracle.dlarnv(y.size)
    assert np.array_equal(x, y) and np.array_equal(seed, s2)
    # continuing the stream from the returned seed
    y2 = np.empty(1000)
    L.lapack_dlarnv(1, seed, y2.size, y2)
    x2, _ = oracle.dlarnv(1000, s2)
    assert np.array_equal(x2, y2)


def test_known_answer_4x4():
    # the first 4x4 of the stream factors with ipiv = 2 3 3 4, info = 0
    A, _ = oracle.random_batch(1, 4, 4)
    ipiv, info = oracle.getrf_batched(A, 4)
    assert ipiv.tolist() == [[2, 3, 3, 4]] and info.tolist() == [0]


@pytest.mark.parametrize("key", SHAPES)
def test_oracle_vs_golden(key):
    m, n, nrhs = _shape(key)
    A0 = GOLD[key + "_A"]
    A = A0.copy()
    if nrhs:
        B = GOLD[key + "_B"].copy()
        ipiv, info = oracle.gesv_batched(A, B, n)
        X = GOLD[key + "_X"]
        assert np.allclose(B, X, rtol=1e-9, atol=1e-11)
        assert oracle.solve_residual(oracle.MagmaNoTrans, A0, B, GOLD[key + "_B"], n) < oracle.TOL
    else:
        ipiv, info = oracle.getrf_batched(A, m)
    assert np.array_equal(ipiv, GOLD[key + "_ipiv"]), "pivots differ from LAPACK"
    assert np.array_equal(info, GOLD[key + "_info"])
    # factors agree with LAPACK's up to rounding-order differences
    ref = GOLD[key + "_LU"]
    assert np.max(np.abs(A - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref)))
    assert oracle.lu_backward_error(A0, A, ipiv, m) < oracle.TOL
    assert oracle.lu_backward_error(A0, ref, GOLD[key + "_ipiv"], m) < oracle.TOL


@pytest.mark.parametrize("m,n,batch", [(16, 16, 3000), (32, 32, 1500), (7, 7, 500), (50, 20, 200), (20, 50, 200),
                                       (128, 128, 40), (257, 257, 4)])
def test_pivots_identical_to_host_lapack(m, n, batch):
    L = oracle.lapack()
    A, _ = oracle.random_batch(batch, m, n, iseed=[1, 2, 3, 5])
    A0 = A.copy()
    ipiv, info = oracle.getrf_batched(A, m)
    mn = min(m, n)
    LU = A0.copy()
    ip2 = np.zeros((batch, mn), np.int32)
    inf2 = np.zeros(batch, np.int32)
    L.lapack_dgetrf_loop(m, n, LU.reshape(-1), m, m * n, ip2.reshape(-1), mn, inf2, batch)
    assert np.array_equal(ipiv, ip2)
    assert np.array_equal(info, inf2)
    assert oracle.lu_backward_error(A0, A, ipiv, m) < oracle.TOL


def test_getrs_trans_and_notrans_vs_lapack():
    L = oracle.lapack()
    n, nrhs, batch = 40, 3, 20
    A, seed = oracle.random_batch(batch, n, n)
    B, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    LU = A.copy()
    ipiv, _ = oracle.getrf_batched(LU, n)
    for trans in (oracle.MagmaNoTrans, oracle.MagmaTrans):
        X = B.copy()
        oracle.getrs_batched(trans, LU, ipiv, X, n)
        X2 = B.copy()
        L.lapack_dgetrs_loop(trans, n, nrhs, LU.reshape(-1), n, n * n, ipiv.reshape(-1), n, X2.reshape(-1), n,
                             n * nrhs, batch)
        assert np.allclose(X, X2, rtol=1e-9, atol=1e-11)
        assert oracle.solve_residual(trans, A, X, B, n) < oracle.TOL


def test_getri_vs_lapack():
    """oracle.getri_outofplace_batched against host LAPACK's inverse (numpy: dgesv on I) and
    the reference tester's residual ||I - A inv(A)||_1 / (n ||A||_1 ||inv(A)||_1) (testing/testing_zgetri_batched.cpp)."""
    for n, batch in ((1, 3), (9, 20), (40, 10), (130, 3)):
        A, _ = oracle.random_batch(batch, n, n)
        LU = A.copy()
        ipiv, info = oracle.getrf_batched(LU, n)
        assert not info.any()
        X = oracle.getri_outofplace_batched(LU, ipiv, n)
        for b in range(batch):
            Ab, Xb = A[b].T, X[b].T  # stored layout is [col][row]
            ref = np.linalg.inv(Ab)
            assert np.allclose(Xb, ref, rtol=1e-8, atol=1e-10 * np.max(np.abs(ref)))
            r = np.linalg.norm(np.eye(n) - Ab @ Xb, 1) / (n * np.linalg.norm(Ab, 1) * np.linalg.norm(Xb, 1))
            assert r < oracle.TOL


def _dominant_batch(batch, m, n, seed=3):
    """Matrices that need no pivoting: random + a dominant diagonal (stored layout [b][col][row])."""
    rng = np.random.default_rng(seed)
    A = rng.random((batch, n, m)) - 0.5
    for k in range(min(m, n)):
        A[:, k, k] += float(max(m, n))
    return A


def test_nopiv_oracle_vs_lapack():
    """oracle.getrf_nopiv_batched: on diagonally dominant matrices partial pivoting never
    interchanges, so host LAPACK's dgetrf must give identity pivots and (to rounding) the same factors; the solve
    from those factors passes the testers' residual check."""
    L = oracle.lapack()
    for m, n, batch in ((1, 1, 3), (7, 7, 9), (40, 40, 6), (33, 20, 4), (20, 33, 4), (130, 130, 2)):
        A0 = _dominant_batch(batch, m, n)
        LU = A0.copy()
        info = oracle.getrf_nopiv_batched(LU, m)
        assert not info.any()
        ref = A0.copy()
        mn = min(m, n)
        ipiv = np.zeros((batch, mn), dtype=np.int32)
        inf2 = np.zeros(batch, dtype=np.int32)
        L.lapack_dgetrf_loop(m, n, ref.reshape(-1), m, n * m, ipiv.reshape(-1), mn, inf2, batch)
        assert np.array_equal(ipiv, np.broadcast_to(np.arange(1, mn + 1, dtype=np.int32), (batch, mn)))
        assert np.allclose(LU, ref, rtol=1e-12, atol=1e-13)
    n, nrhs, batch = 40, 3, 5
    A0 = _dominant_batch(batch, n, n)
    B0 = np.random.default_rng(4).random((batch, nrhs, n))
    LU = A0.copy()
    oracle.getrf_nopiv_batched(LU, n)
    for trans in (oracle.MagmaNoTrans, oracle.MagmaTrans):
        X = B0.copy()
        oracle.getrs_nopiv_batched(trans, LU, X, n)
        assert oracle.solve_residual(trans, A0, X, B0, n) < oracle.TOL
    # zero diagonal: info = first such column, factorisation completed with the column unscaled
    Z = _dominant_batch(1, 6, 6)
    Z[0, 2, :] = 0.0   # column 2 of the matrix entirely zero -> A(2,2) stays 0 after two steps
    info = oracle.getrf_nopiv_batched(Z, 6)
    assert info[0] == 3 and np.isfinite(Z).all()


def test_singular_and_tie_semantics():
    # zero matrix: info = 1, ipiv = identity, nothing scaled (smallsq_noshfl.cu:90-91,106)
    Z = np.zeros((1, 5, 5))
    ipiv, info = oracle.getrf_batched(Z, 5)
    assert info.tolist() == [1] and ipiv.tolist() == [[1, 2, 3, 4, 5]] and not Z.any()
    # all ones: every column ties, LAPACK's idamax takes the first -> ipiv = identity, info = 2
    O = np.ones((1, 4, 4))
    ipiv, info = oracle.getrf_batched(O, 4)
    L = oracle.lapack()
    O2 = np.ones(16)
    ip2 = np.zeros(4, np.int32)
    inf2 = np.zeros(1, np.int32)
    L.lapack_dgetrf_loop(4, 4, O2, 4, 16, ip2, 4, inf2, 1)
    assert ipiv.reshape(-1).tolist() == ip2.tolist() and info.tolist() == inf2.tolist() == [2]
    # a zero column in the middle
    A, _ = oracle.random_batch(1, 6, 6)
    A[0, 2, :] = 0.0
    A2 = A.copy().reshape(-1)
    ipiv, info = oracle.getrf_batched(A, 6)
    ip2 = np.zeros(6, np.int32)
    L.lapack_dgetrf_loop(6, 6, A2, 6, 36, ip2, 6, inf2, 1)
    assert info.tolist() == inf2.tolist() == [3] and ipiv.reshape(-1).tolist() == ip2.tolist()


def test_checker_detects_errors():
    A, _ = oracle.random_batch(4, 12, 12)
    LU = A.copy()
    ipiv, _ = oracle.getrf_batched(LU, 12)
    assert oracle.lu_backward_error(A, LU, ipiv, 12) < oracle.TOL
    bad = LU.copy()
    bad[2, 5, 7] += 1e-6
    assert oracle.lu_backward_error(A, bad, ipiv, 12) > oracle.TOL
    bad = LU.copy()
    bad[1, 0, 0] = np.nan
    assert oracle.lu_backward_error(A, bad, ipiv, 12) > oracle.TOL
    badp = ipiv.copy()
    badp[0, 0], badp[0, 1] = 12, 12
    assert oracle.lu_backward_error(A, LU, badp, 12) > oracle.TOL


def test_flops_formulas():
    # testing/flops.h:79-91 evaluated at the BASELINE.md sizes
    assert round(oracle.flops_getrf(32, 32)) == 21360
    assert round(oracle.flops_getrf(128, 128)) == 1390016
    assert round(oracle.flops_getrf(512, 512)) == 89347840
    assert oracle.flops_getrs(512, 16) == 8380416
    assert round(oracle.flops_getrf(16, 16)) == 2616 and oracle.flops_getrs(16, 1) == 496


def test_rbt_restatement_is_consistent():
    """oracle.prbt / prbt_mtv / prbt_mv (numpy restatement of magmablas/zgerbt_kernels.cu): the vector routines define
    the dense butterflies U^T and V; the matrix routine must equal U^T A V, and the whole chain must solve A X = B."""
    rng = np.random.default_rng(7)
    for n in (1, 2, 3, 7, 8, 33, 64):
        u = np.exp((rng.random(2 * n) - 0.5) / 10)
        v = np.exp((rng.random(2 * n) - 0.5) / 10)
        eye = np.ascontiguousarray(np.eye(n)[None])          # [1, col, row]
        Ut = eye.copy(); oracle.prbt_mtv(Ut, n, u)            # column j of U^T
        V = eye.copy(); oracle.prbt_mv(V, n, v)
        Ut, V = Ut[0].T, V[0].T                               # [row, col]
        A = rng.random((2, n, n))
        ref = np.stack([(Ut @ A[b].T @ V).T for b in range(2)])
        got = A.copy(); oracle.prbt(got, n, u, v)
        if n >= 4:  # (a half of size 1, i.e. n <= 3: the reference's transposed vector routine returns early for
            #  n < 2 while its matrix routine scales -- zgerbt_kernels.cu:157 vs :21-80 -- and the restatement follows it)
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-15)
    n, batch, nrhs = 48, 5, 3
    A0 = rng.random((batch, n, n)); B0 = rng.random((batch, nrhs, n))
    u = np.exp((rng.random(2 * n) - 0.5) / 10); v = np.exp((rng.random(2 * n) - 0.5) / 10)
    A, B = A0.copy(), B0.copy()
    info = oracle.gesv_rbt_batched(A, B, n, u, v)
    assert not info.any()
    assert oracle.solve_residual(oracle.MagmaNoTrans, A0, B, B0, n) < 1e-9  # no pivoting: looser than the 30 eps bar


def test_scz_template_equals_double_oracle():
    """lu_oracle_tmpl.h instantiated in double ("q") reproduces oracle_dgetf2 / oracle_dgetrs / oracle_dgesv bit for bit:
    the s / c / z restatements are the same algorithm, not a second one."""
    A0, _ = oracle.random_batch(7, 45, 38)
    a, b = A0.copy(), A0.copy()
    ip1, in1 = oracle.getrf_batched(a, 45)
    ip2, in2 = oracle.getrf_batched_prec("q", b, 45)
    assert np.array_equal(a, b) and np.array_equal(ip1, ip2) and np.array_equal(in1, in2)
    A0, _ = oracle.random_batch(5, 30, 30)
    B0, _ = oracle.random_batch(5, 30, 4)
    LU = A0.copy()
    ip, _ = oracle.getrf_batched(LU, 30)
    for tr in (111, 112, 113):
        x1, x2 = B0.copy(), B0.copy()
        oracle.getrs_batched(tr, LU, ip, x1, 30)
        oracle.getrs_batched_prec("q", tr, LU, ip, x2, 30)
        assert np.array_equal(x1, x2)
    a1, b1, a2, b2 = A0.copy(), B0.copy(), A0.copy(), B0.copy()
    oracle.gesv_batched(a1, b1, 30)
    oracle.gesv_batched_prec("q", a2, b2, 30)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)


@pytest.mark.parametrize("p", ["s", "c", "z"])
def test_scz_oracle_vs_lapack(p):
    """oracle_{s,c,z}getf2 / getrs against LAPACK (scipy): identical pivots, factors and solutions to a small multiple of
    n * eps of the precision (the reference testers' criterion, testing/testing_zgetrf_batched.cpp:41-81)."""
    import scipy.linalg as sl
    eps = float(np.finfo(oracle.PREC_DTYPE[p]).eps)
    for n in (8, 32, 64, 150):
        A = oracle.random_batch_prec(p, 3, n, n, seed=n)
        LU = A.copy()
        ip, info = oracle.getrf_batched_prec(p, LU, n)
        assert not info.any()
        for b in range(3):
            lu, piv = sl.lu_factor(A[b].T)
            assert np.array_equal(piv + 1, ip[b])
            assert np.abs(lu - LU[b].T).max() <= 4 * n * eps * np.abs(lu).max()
        B = oracle.random_batch_prec(p, 3, n, 2, seed=n + 1)
        for tr in (111, 112, 113):
            X = B.copy()
            oracle.getrs_batched_prec(p, tr, LU, ip, X, n)
            for b in range(3):
                M = A[b].T
                op = {111: M, 112: M.T, 113: M.conj().T}[tr]
                r = np.linalg.norm(op @ X[b].T - B[b].T, 1) / (n * np.linalg.norm(M, 1) * np.linalg.norm(X[b].T, 1))
                assert r < 30 * eps
    # rectangular: P A = L U
    for (m, n) in ((40, 25), (25, 40)):
        A = oracle.random_batch_prec(p, 2, m, n, seed=m)
        LU = A.copy()
        ip, _ = oracle.getrf_batched_prec(p, LU, m)
        k = min(m, n)
        for b in range(2):
            M = A[b].T.copy()
            for i in range(k):
                pi = ip[b, i] - 1
                M[[i, pi]] = M[[pi, i]]
            Lm = np.tril(LU[b].T[:, :k], -1) + np.eye(m, k)
            Um = np.triu(LU[b].T[:k, :])
            assert np.abs(M - Lm @ Um).max() <= 4 * k * eps * np.abs(M).max()


GOLD_SCZ = np.load(os.path.join(os.path.dirname(__file__), "golden", "lu_golden_scz.npz"))


@pytest.mark.parametrize("key", sorted({k.rsplit("_", 1)[0] for k in GOLD_SCZ.files if k.endswith("_A")}))
def test_scz_oracle_vs_golden(key):
    """oracle_{s,c,z}getf2 / getrs against the committed LAPACK {s,c,z}getrf / getrs vectors (tests/golden/make_golden_scz.py):
    pivots identical, factors and solutions within a few n * eps of the precision."""
    p = key[0]
    A = GOLD_SCZ[key + "_A"]
    m, n = A.shape
    eps = float(np.finfo(oracle.PREC_DTYPE[p]).eps)
    LU = np.ascontiguousarray(A.T)[None].copy()  # stored layout (batch, n, ld = m)
    ip, info = oracle.getrf_batched_prec(p, LU, m)
    assert info[0] == 0
    assert np.array_equal(ip[0], GOLD_SCZ[key + "_ipiv"])
    want = GOLD_SCZ[key + "_LU"]
    assert np.abs(LU[0].T - want).max() <= 4 * max(m, n) * eps * np.abs(want).max()
    if m == n:
        B = GOLD_SCZ[key + "_B"]
        for tr, name in ((111, "N"), (112, "T"), (113, "C")):
            X = np.ascontiguousarray(B.T)[None].copy()
            oracle.getrs_batched_prec(p, tr, LU, ip, X, n)
            wx = GOLD_SCZ[key + "_X" + name]
            assert np.abs(X[0].T - wx).max() <= 40 * n * eps * np.abs(wx).max()
