/* TEST INFRASTRUCTURE -- body of the s / c / z restatements, included once per precision by lu_oracle_scz.c with
 *   PFX(name)   -> oracle_<p>name        R -> float | double        CPLX -> 0 | 1
 *   FMA, FABS   -> fmaf/fabsf | fma/fabs
 * Same algorithm and canonical rounding order as oracle_dgetf2 / oracle_dgetrs (lu_oracle.c), i.e. the reference's
 * src/zgetrf_batched.cpp:11 ("@precisions normal z -> s d c": one master, four precisions) and src/zgetrs_batched.cpp.
 * Complex specifics follow the reference's device code:
 *   pivot metric |re| + |im|           magmablas/zgetf2_devicefunc.cuh:39,248 (MAGMA_Z_ABS1, include/magma_types.h:132)
 *   multiplier = a * (1 / pivot)       magmablas/zgetf2_devicefunc.cuh:123,155 (MAGMA_Z_DIV(ONE, pivot) = cuCdiv: scaled by
 *                                      1 / (|re| + |im|) before the squares)
 * The order of the four FMAs of a complex multiply-subtract and of the products inside the reciprocal is fixed HERE
 * (nvcc is free to contract the reference's expressions either way); the CUDA kernels (csrc/lu_scz.cu) use exactly
 * these sequences, so results are bit-identical; against LAPACK the oracle is checked to n * eps (tests/test_oracle.py).
 */
#if CPLX
typedef struct { R x, y; } T;
static inline R PFX(abs1)(T a) { return FABS(a.x) + FABS(a.y); }
static inline int PFX(iszero)(T a) { return a.x == (R)0 && a.y == (R)0; }
static inline T PFX(rcp)(T p)
{
    R s = FABS(p.x) + FABS(p.y);
    R oos = (R)1 / s;
    R brs = p.x * oos, bis = p.y * oos;
    R d = FMA(brs, brs, bis * bis);
    R ood = (R)1 / d;
    T r;
    r.x = (brs * ood) * oos;
    r.y = (-(bis * ood)) * oos;
    return r;
}
static inline T PFX(mul)(T a, T b)
{
    T c;
    c.x = FMA(a.x, b.x, -(a.y * b.y));
    c.y = FMA(a.x, b.y, a.y * b.x);
    return c;
}
static inline T PFX(fnma)(T l, T u, T a) /* a - l*u */
{
    T c;
    c.x = FMA(-l.x, u.x, a.x);
    c.x = FMA(l.y, u.y, c.x);
    c.y = FMA(-l.x, u.y, a.y);
    c.y = FMA(-l.y, u.x, c.y);
    return c;
}
static inline T PFX(conj_if)(T a, int c) { if (c) a.y = -a.y; return a; }
#else
typedef R T;
static inline R PFX(abs1)(T a) { return FABS(a); }
static inline int PFX(iszero)(T a) { return a == (R)0; }
static inline T PFX(rcp)(T p) { return (R)1 / p; }
static inline T PFX(mul)(T a, T b) { return a * b; }
static inline T PFX(fnma)(T l, T u, T a) { return FMA(-l, u, a); }
static inline T PFX(conj_if)(T a, int c) { (void)c; return a; }
#endif

int PFX(getf2)(int m, int n, T *A, int lda, int *ipiv)
{
    int info = 0;
    int mn = m < n ? m : n;
    for (int k = 0; k < mn; ++k) {
        T *ck = A + (size_t)k * lda;
        int p = k;
        R best = PFX(abs1)(ck[k]);
        for (int i = k + 1; i < m; ++i) {
            R v = PFX(abs1)(ck[i]);
            if (v > best) { best = v; p = i; }
        }
        ipiv[k] = p + 1;
        if (PFX(iszero)(ck[p])) {
            if (info == 0) info = k + 1;
            continue;
        }
        if (p != k)
            for (int j = 0; j < n; ++j) {
                T t = A[k + (size_t)j * lda];
                A[k + (size_t)j * lda] = A[p + (size_t)j * lda];
                A[p + (size_t)j * lda] = t;
            }
        T r = PFX(rcp)(ck[k]);
        for (int i = k + 1; i < m; ++i) ck[i] = PFX(mul)(ck[i], r);
        for (int j = k + 1; j < n; ++j) {
            T *cj = A + (size_t)j * lda;
            T u = cj[k];
            for (int i = k + 1; i < m; ++i) cj[i] = PFX(fnma)(ck[i], u, cj[i]);
        }
    }
    return info;
}

static void PFX(swaprows)(int k, int nrhs, T *B, int ldb, const int *ipiv, int forward)
{
    for (int s = 0; s < k; ++s) {
        int i = forward ? s : k - 1 - s;
        int p = ipiv[i] - 1;
        if (p != i)
            for (int j = 0; j < nrhs; ++j) {
                T t = B[i + (size_t)j * ldb];
                B[i + (size_t)j * ldb] = B[p + (size_t)j * ldb];
                B[p + (size_t)j * ldb] = t;
            }
    }
}

/* trans: 111 N, 112 T, 113 C. Orders as oracle_dgetrs. */
void PFX(getrs)(int trans, int n, int nrhs, const T *A, int lda, const int *ipiv, T *B, int ldb)
{
    if (n == 0 || nrhs == 0) return;
    if (trans == 111) {
        PFX(swaprows)(n, nrhs, B, ldb, ipiv, 1);
        for (int j = 0; j < nrhs; ++j) {
            T *b = B + (size_t)j * ldb;
            for (int k = 0; k < n; ++k) {
                T bk = b[k];
                const T *ck = A + (size_t)k * lda;
                for (int i = k + 1; i < n; ++i) b[i] = PFX(fnma)(ck[i], bk, b[i]);
            }
            for (int k = n - 1; k >= 0; --k) {
                const T *ck = A + (size_t)k * lda;
                T bk = PFX(mul)(b[k], PFX(rcp)(ck[k]));
                b[k] = bk;
                for (int i = 0; i < k; ++i) b[i] = PFX(fnma)(ck[i], bk, b[i]);
            }
        }
    } else {
        const int cj_ = (trans == 113);
        for (int j = 0; j < nrhs; ++j) {
            T *b = B + (size_t)j * ldb;
            for (int i = 0; i < n; ++i) {
                const T *ci = A + (size_t)i * lda;
                T s = b[i];
                for (int k = 0; k < i; ++k) s = PFX(fnma)(PFX(conj_if)(ci[k], cj_), b[k], s);
                b[i] = PFX(mul)(s, PFX(rcp)(PFX(conj_if)(ci[i], cj_)));
            }
            for (int i = n - 1; i >= 0; --i) {
                const T *ci = A + (size_t)i * lda;
                T s = b[i];
                for (int k = n - 1; k > i; --k) s = PFX(fnma)(PFX(conj_if)(ci[k], cj_), b[k], s);
                b[i] = s;
            }
        }
        PFX(swaprows)(n, nrhs, B, ldb, ipiv, 0);
    }
}

void PFX(getrf_batched)(int m, int n, T *A, int lda, long strideA, int *ipiv, long stride_ipiv, int *info, long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b) info[b] = PFX(getf2)(m, n, A + b * strideA, lda, ipiv + b * stride_ipiv);
}

void PFX(getrs_batched)(int trans, int n, int nrhs, const T *A, int lda, long strideA, const int *ipiv, long stride_ipiv, T *B,
                        int ldb, long strideB, long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b)
        PFX(getrs)(trans, n, nrhs, A + b * strideA, lda, ipiv + b * stride_ipiv, B + b * strideB, ldb);
}

/* the reference solves regardless of info (src/zgesv_batched.cpp:129-152) */
void PFX(gesv_batched)(int n, int nrhs, T *A, int lda, long strideA, int *ipiv, long stride_ipiv, T *B, int ldb, long strideB,
                       int *info, long batch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < batch; ++b) {
        info[b] = PFX(getf2)(n, n, A + b * strideA, lda, ipiv + b * stride_ipiv);
        PFX(getrs)(111, n, nrhs, A + b * strideA, lda, ipiv + b * stride_ipiv, B + b * strideB, ldb);
    }
}

void PFX(getrf_vbatched)(const int *m, const int *n, T *A, const int *lda, const long *offsets, int *ipiv, const long *ipiv_off,
                         int *info, long batch)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (long b = 0; b < batch; ++b) info[b] = PFX(getf2)(m[b], n[b], A + offsets[b], lda[b], ipiv + ipiv_off[b]);
}
