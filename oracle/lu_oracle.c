/*
 * oracle/lu_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement, in plain C, of the arithmetic the reference's batched FP64 LU path performs:
 * LAPACK-style partial-pivoting LU ("first max |x|" pivot, reciprocal scaling, 1-based ipiv,
 * info = first exactly-zero pivot, keep going on singular input), forward/backward substitution
 * for getrs/gesv, the dlarnv input stream of the testers and the testers' error checks.
 *
 * What it follows in /root/reference (z masters -> d by the reference's codegen):
 *   - unblocked right-looking column loop, strict '>' pivot scan, zero-pivot handling, reciprocal
 *     scale, rank-1 update                         magmablas/zgetrf_batched_smallsq_noshfl.cu:76-116
 *   - fused panel with the same arithmetic          magmablas/zgetf2_devicefunc.cuh:215-294
 *   - blocked driver whose result this must equal   src/zgetrf_batched.cpp:81-213
 *   - solve = row interchanges, unit-lower forward, non-unit-upper backward
 *                                                   src/zgetrs_batched.cpp:118-146
 *   - fused small gesv                              magmablas/zgesv_batched_small.cu:47-116
 *   - trsm: multiply by the inverted diagonal       magmablas/trsm_template_device.cuh:54-58
 *   - input stream dlarnv(1, {0,0,0,1})             testing/testing_zgetrf_batched.cpp:136,180
 *   - checks ||PA-LU||_F/(n ||A||_F), residual      testing/testing_zgetrf_batched.cpp:41-81,
 *                                                   testing/testing_zgesv_batched.cpp:133-153
 *
 * Rounding model ("canonical order"): every element a(i,j) receives its updates
 *     a(i,j) <- fma(-l(i,k), u(k,j), a(i,j))   for k = 0,1,...,min(i,j)-1 in increasing k,
 * multipliers are l(i,k) = a(i,k) * (1/pivot) with an IEEE-correct reciprocal. Any blocked
 * variant that applies the same per-element fma sequence produces bit-identical factors; the
 * CUDA kernels of this repo are written to do exactly that, so the GPU tests demand bit
 * equality against this file, not just a backward-error bound. That includes the kernels on the
 * FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA) was measured on the B200 to equal a chain of four
 * FMAs with k increasing (tools/dmma_probe.cu, 262144 of 262144 outputs), and the kernels feed it
 * the operands in canonical k order.
 *
 * The host LAPACK the reference's testers call (third-party; OpenBLAS 0.3.31.dev bundled with
 * scipy in this image) is NOT restated here: tests/ pin this oracle against it (identical
 * pivots on the testers' dlarnv stream; factors equal to a few ulp; known-answer values from
 * SURVEY.md section 8c) and against the committed fixtures under tests/golden/.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * dlarnv(idist=1): LAPACK's dlaruv is the 48-bit multiplicative congruential generator
 *   x <- a*x mod 2^48,  a = 33952834046453,  u = x * 2^-48
 * (its 128x4 multiplier table holds a^1..a^128 split in 12-bit digits). iseed[0..3] are the
 * base-4096 digits of x, most significant first; iseed[3] must be odd. The returned seed is the
 * state after n draws, exactly as dlarnv leaves it (testers draw A then B from one stream,
 * testing/testing_zgesv_batched.cpp:96-97).
 * ------------------------------------------------------------------------------------------ */
#define LCG_A   33952834046453ULL
#define MASK48  ((1ULL << 48) - 1)

static uint64_t seed_pack(const int *iseed)
{
    return ((uint64_t)(iseed[0] & 4095) << 36) | ((uint64_t)(iseed[1] & 4095) << 24) |
           ((uint64_t)(iseed[2] & 4095) << 12) | (uint64_t)(iseed[3] & 4095);
}

static void seed_unpack(uint64_t x, int *iseed)
{
    iseed[0] = (int)((x >> 36) & 4095);
    iseed[1] = (int)((x >> 24) & 4095);
    iseed[2] = (int)((x >> 12) & 4095);
    iseed[3] = (int)(x & 4095);
}

void oracle_dlarnv_uniform01(int *iseed, long n, double *x)
{
    uint64_t s = seed_pack(iseed);
    for (long i = 0; i < n; ++i) {
        s = (s * LCG_A) & MASK48;
        x[i] = (double)s * 0x1.0p-48;
    }
    seed_unpack(s, iseed);
}

/* ------------------------------------------------------------------------------------------
 * One matrix: right-looking LU with partial pivoting in the canonical rounding order.
 * A is m x n column-major, lda >= m. ipiv has min(m,n) 1-based entries. Returns info.
 * ------------------------------------------------------------------------------------------ */
int oracle_dgetf2(int m, int n, double *A, int lda, int *ipiv)
{
    int info = 0;
    int mn = m < n ? m : n;
    for (int k = 0; k < mn; ++k) {
        double *ck = A + (size_t)k * lda;
        /* idamax: first index of the largest |x| (strict '>' scan) */
        int p = k;
        double best = fabs(ck[k]);
        for (int i = k + 1; i < m; ++i) {
            double v = fabs(ck[i]);
            if (v > best) { best = v; p = i; }
        }
        ipiv[k] = p + 1;
        if (ck[p] == 0.0) {
            /* exactly singular: record the first such column, no swap effect, no scaling */
            if (info == 0) info = k + 1;
            if (p != k) { /* cannot happen: all candidates are zero so p stays k */ }
            continue;
        }
        if (p != k) {
            for (int j = 0; j < n; ++j) {
                double t = A[k + (size_t)j * lda];
                A[k + (size_t)j * lda] = A[p + (size_t)j * lda];
                A[p + (size_t)j * lda] = t;
            }
        }
        double r = 1.0 / ck[k];
        for (int i = k + 1; i < m; ++i) ck[i] *= r;
        for (int j = k + 1; j < n; ++j) {
            double *cj = A + (size_t)j * lda;
            double u = cj[k];
            for (int i = k + 1; i < m; ++i) cj[i] = fma(-ck[i], u, cj[i]);
        }
    }
    return info;
}

/* ------------------------------------------------------------------------------------------
 * LU without pivoting: magma_dgetrf_nopiv_batched (src/zgetrf_nopiv_batched.cpp:75-170; panel arithmetic
 * magmablas/zgetf2_nopiv_kernels.cu:22-72: reciprocal of the diagonal, scale, rank-1 update). info = first i with
 * A(i,i) == 0 at its turn (1-based). At a zero diagonal the column is left unscaled ("reg = 1", :53) and the
 * update still runs; the reference then abandons the rest of the matrix (:32-34), this restatement -- like the
 * CUDA path it checks -- completes the factorisation the way LAPACK's dgetf2 would.
 * ------------------------------------------------------------------------------------------ */
int oracle_dgetf2_nopiv(int m, int n, double *A, int lda)
{
    int info = 0;
    int mn = m < n ? m : n;
    for (int k = 0; k < mn; ++k) {
        double *ck = A + (size_t)k * lda;
        double r = 1.0;
        if (ck[k] == 0.0) {
            if (info == 0) info = k + 1;
        } else {
            r = 1.0 / ck[k];
        }
        for (int i = k + 1; i < m; ++i) ck[i] *= r;
        for (int j = k + 1; j < n; ++j) {
            double *cj = A + (size_t)j * lda;
            double u = cj[k];
            for (int i = k + 1; i < m; ++i) cj[i] = fma(-ck[i], u, cj[i]);
        }
    }
    return info;
}

void oracle_dgetrf_nopiv_batched(int m, int n, double *A, int lda, long strideA, int *info, long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b) info[b] = oracle_dgetf2_nopiv(m, n, A + b * strideA, lda);
}

/* LAPACK dlaswp semantics on B (n rows touched by ipiv[0..k-1]), forward or backward. */
static void apply_pivots(int k, int nrhs, double *B, int ldb, const int *ipiv, int forward)
{
    if (forward) {
        for (int i = 0; i < k; ++i) {
            int p = ipiv[i] - 1;
            if (p != i)
                for (int j = 0; j < nrhs; ++j) {
                    double t = B[i + (size_t)j * ldb];
                    B[i + (size_t)j * ldb] = B[p + (size_t)j * ldb];
                    B[p + (size_t)j * ldb] = t;
                }
        }
    } else {
        for (int i = k - 1; i >= 0; --i) {
            int p = ipiv[i] - 1;
            if (p != i)
                for (int j = 0; j < nrhs; ++j) {
                    double t = B[i + (size_t)j * ldb];
                    B[i + (size_t)j * ldb] = B[p + (size_t)j * ldb];
                    B[p + (size_t)j * ldb] = t;
                }
        }
    }
}

/*
 * Solve with the factors. trans: 111 NoTrans, 112 Trans, 113 ConjTrans (MAGMA enum values,
 * include/magma_types.h:612-614). Canonical order:
 *   NoTrans  forward : b(i) <- fma(-l(i,k), b(k), b(i)) for k increasing;
 *            backward: for k = n-1..0: b(k) <- b(k)*(1/u(k,k)); b(i) <- fma(-u(i,k), b(k), b(i)), i<k.
 *            The diagonal is inverted once and multiplied, as the reference's batched trsm does
 *            (magmablas/trsm_template_device.cuh:54-58,143); its fused n<=32 gesv kernel divides
 *            instead (magmablas/zgesv_batched_small.cu:111) -- one rule is used everywhere here.
 *   Trans    LAPACK dgetrs('T'): solve U^T (non-unit) forward, L^T (unit) backward, then the
 *            interchanges in reverse. (The reference's batched Trans branch swaps the diag flags
 *            and applies the interchanges forward, src/zgetrs_batched.cpp:148-178; that is a
 *            defect -- SURVEY.md section 3.3 -- and is not restated.)
 */
void oracle_dgetrs(int trans, int n, int nrhs, const double *A, int lda, const int *ipiv,
                   double *B, int ldb)
{
    if (n == 0 || nrhs == 0) return;
    if (trans == 111) {
        apply_pivots(n, nrhs, B, ldb, ipiv, 1);
        for (int j = 0; j < nrhs; ++j) {
            double *b = B + (size_t)j * ldb;
            for (int k = 0; k < n; ++k) {
                double bk = b[k];
                const double *ck = A + (size_t)k * lda;
                for (int i = k + 1; i < n; ++i) b[i] = fma(-ck[i], bk, b[i]);
            }
            for (int k = n - 1; k >= 0; --k) {
                const double *ck = A + (size_t)k * lda;
                double bk = b[k] * (1.0 / ck[k]); /* reciprocal of the diagonal, as the reference's trsm */
                b[k] = bk;
                for (int i = 0; i < k; ++i) b[i] = fma(-ck[i], bk, b[i]);
            }
        }
    } else {
        for (int j = 0; j < nrhs; ++j) {
            double *b = B + (size_t)j * ldb;
            /* U^T y = b : row i of U^T is column i of U */
            for (int i = 0; i < n; ++i) {
                const double *ci = A + (size_t)i * lda;
                double s = b[i];
                for (int k = 0; k < i; ++k) s = fma(-ci[k], b[k], s);
                b[i] = s * (1.0 / ci[i]);
            }
            /* L^T x = y (unit) */
            for (int i = n - 1; i >= 0; --i) {
                const double *ci = A + (size_t)i * lda;
                double s = b[i];
                for (int k = n - 1; k > i; --k) s = fma(-ci[k], b[k], s);
                b[i] = s;
            }
        }
        apply_pivots(n, nrhs, B, ldb, ipiv, 0);
    }
}

int oracle_dgesv(int n, int nrhs, double *A, int lda, int *ipiv, double *B, int ldb)
{
    int info = oracle_dgetf2(n, n, A, lda, ipiv);
    /* the reference solves regardless of info (src/zgesv_batched.cpp:129-152 without CHECK_INFO) */
    oracle_dgetrs(111, n, nrhs, A, lda, ipiv, B, ldb);
    return info;
}

/* ------------------------------------------------------------------------------------------
 * Batched / variable-size wrappers over contiguous storage (stride in elements).
 * ------------------------------------------------------------------------------------------ */
void oracle_dgetrf_batched(int m, int n, double *A, int lda, long strideA, int *ipiv,
                           long stride_ipiv, int *info, long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b)
        info[b] = oracle_dgetf2(m, n, A + b * strideA, lda, ipiv + b * stride_ipiv);
}

void oracle_dgetrs_batched(int trans, int n, int nrhs, const double *A, int lda, long strideA,
                           const int *ipiv, long stride_ipiv, double *B, int ldb, long strideB,
                           long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b)
        oracle_dgetrs(trans, n, nrhs, A + b * strideA, lda, ipiv + b * stride_ipiv,
                      B + b * strideB, ldb);
}

void oracle_dgesv_batched(int n, int nrhs, double *A, int lda, long strideA, int *ipiv,
                          long stride_ipiv, double *B, int ldb, long strideB, int *info,
                          long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b)
        info[b] = oracle_dgesv(n, nrhs, A + b * strideA, lda, ipiv + b * stride_ipiv,
                               B + b * strideB, ldb);
}

/* offsets[b] = element offset of matrix b in A, ipiv_off[b] likewise for ipiv */
void oracle_dgetrf_vbatched(const int *m, const int *n, double *A, const int *lda,
                            const long *offsets, int *ipiv, const long *ipiv_off, int *info,
                            long batch)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (long b = 0; b < batch; ++b)
        info[b] = oracle_dgetf2(m[b], n[b], A + offsets[b], lda[b], ipiv + ipiv_off[b]);
}

/* ------------------------------------------------------------------------------------------
 * The testers' checks.
 * ------------------------------------------------------------------------------------------ */

/* ||P*A0 - L*U||_F / (||A0||_F * n)   (testing/testing_zgetrf_batched.cpp:41-81) */
double oracle_lu_backward_error(int m, int n, const double *A0, int lda0, const double *LU,
                                int ldlu, const int *ipiv)
{
    int mn = m < n ? m : n;
    double *PA = (double *)malloc(sizeof(double) * (size_t)m * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) PA[i + (size_t)j * m] = A0[i + (size_t)j * lda0];
    apply_pivots(mn, n, PA, m, ipiv, 1);
    long double num = 0, den = 0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            int kmax = i < j ? i : j; /* L(i,k) k<=i (unit diag), U(k,j) k<=j */
            long double s = 0;
            for (int k = 0; k <= kmax && k < mn; ++k) {
                long double l = (k == i) ? 1.0L : (long double)LU[i + (size_t)k * ldlu];
                s += l * (long double)LU[k + (size_t)j * ldlu];
            }
            long double d = (long double)PA[i + (size_t)j * m] - s;
            num += d * d;
            long double a = A0[i + (size_t)j * lda0];
            den += a * a;
        }
    free(PA);
    if (den == 0) return num == 0 ? 0.0 : INFINITY;
    return (double)(sqrtl(num) / (sqrtl(den) * n));
}

/* ||B - A*X||_inf / (n * ||A||_inf * ||X||_inf)   (testing/testing_zgesv_batched.cpp:133-153) */
double oracle_solve_residual(int trans, int n, int nrhs, const double *A, int lda,
                             const double *X, int ldx, const double *B, int ldb)
{
    double anorm = 0, xnorm = 0, rnorm = 0;
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j)
            s += fabs(trans == 111 ? A[i + (size_t)j * lda] : A[j + (size_t)i * lda]);
        if (s > anorm) anorm = s;
    }
    for (int i = 0; i < n; ++i) {
        double sx = 0, sr = 0;
        for (int c = 0; c < nrhs; ++c) {
            long double acc = B[i + (size_t)c * ldb];
            for (int j = 0; j < n; ++j) {
                double a = trans == 111 ? A[i + (size_t)j * lda] : A[j + (size_t)i * lda];
                acc -= (long double)a * (long double)X[j + (size_t)c * ldx];
            }
            sr += fabs((double)acc);
            sx += fabs(X[i + (size_t)c * ldx]);
        }
        if (sx > xnorm) xnorm = sx;
        if (sr > rnorm) rnorm = sr;
    }
    if (anorm == 0 || xnorm == 0) return rnorm == 0 ? 0.0 : INFINITY;
    return rnorm / (n * anorm * xnorm);
}

/* max over a contiguous batch, OpenMP */
double oracle_lu_backward_error_batched(int m, int n, const double *A0, int lda0, long strideA0,
                                        const double *LU, int ldlu, long strideLU,
                                        const int *ipiv, long stride_ipiv, long batch)
{
    double worst = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(max : worst)
    for (long b = 0; b < batch; ++b) {
        double e = oracle_lu_backward_error(m, n, A0 + b * strideA0, lda0, LU + b * strideLU, ldlu,
                                            ipiv + b * stride_ipiv);
        if (isnan(e)) e = INFINITY; /* a NaN anywhere must fail the check */
        if (e > worst) worst = e;
    }
    return worst;
}

double oracle_solve_residual_batched(int trans, int n, int nrhs, const double *A, int lda,
                                     long strideA, const double *X, int ldx, long strideX,
                                     const double *B, int ldb, long strideB, long batch)
{
    double worst = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(max : worst)
    for (long b = 0; b < batch; ++b) {
        double e = oracle_solve_residual(trans, n, nrhs, A + b * strideA, lda, X + b * strideX, ldx,
                                         B + b * strideB, ldb);
        if (isnan(e)) e = INFINITY;
        if (e > worst) worst = e;
    }
    return worst;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * The LU trailing update as a stand-alone product (magma_dgemm_batched with alpha = -1, beta = 1, NoTrans/NoTrans:
 * src/zgetrf_batched.cpp:195-200 calls magma_zgemm_batched_core this way): per element the canonical chain
 *     c(i,j) <- fma(-a(i,k), b(k,j), c(i,j)),  k = 0 .. K-1 in increasing order.
 * ------------------------------------------------------------------------------------------ */
void oracle_dgemm_lu_update(int m, int n, int k, const double *A, int lda, const double *B, int ldb, double *C, int ldc)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double c = C[i + (size_t)j * ldc];
            for (int kk = 0; kk < k; ++kk) c = fma(-A[i + (size_t)kk * lda], B[kk + (size_t)j * ldb], c);
            C[i + (size_t)j * ldc] = c;
        }
}
