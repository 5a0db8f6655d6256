/*
 * oracle/lapack_loop.c -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's own CPU path for batched LU is not MAGMA code: its testers time a host
 * LAPACK loop,
 *     #pragma omp parallel for schedule(dynamic)   over the batch,
 *     one dgetrf_ (then dgetrs_ / dgesv_) per matrix, BLAS threads forced to 1,
 * testing/testing_zgetrf_batched.cpp:254-278 and testing/testing_zgesv_batched.cpp:158-180.
 * This file is that loop. The LAPACK it drives is the third-party library the reference would
 * be linked with; in this image that is OpenBLAS 0.3.31.dev as bundled by scipy
 * (site-packages/scipy.libs/libscipy_openblas-*.so, LP64, symbols prefixed "scipy_"). It is
 * opened with dlopen at run time (path supplied by the Python side) so nothing here is a
 * link-time dependency of the repo.
 *
 * Used for: (1) pinning oracle/lu_oracle.c (pivots, factors, dlarnv stream) in tests/,
 * (2) bench.py's cpu_baseline leg and `--impl reference` arm.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*getrf_fn)(const int *, const int *, double *, const int *, int *, int *);
typedef void (*getrs_fn)(const char *, const int *, const int *, const double *, const int *,
                         const int *, double *, const int *, int *, size_t);
typedef void (*gesv_fn)(const int *, const int *, double *, const int *, int *, double *,
                        const int *, int *);
typedef void (*larnv_fn)(const int *, int *, const int *, double *);
typedef void (*setthr_fn)(int);

static void *g_handle;
static getrf_fn g_getrf;
static getrs_fn g_getrs;
static gesv_fn g_gesv;
static larnv_fn g_larnv;
static setthr_fn g_setthr;
static char g_desc[256] = "uninitialised";

static void *sym2(const char *a, const char *b)
{
    void *p = dlsym(g_handle, a);
    if (!p && b) p = dlsym(g_handle, b);
    return p;
}

/* returns 0 on success */
int lapack_loop_init(const char *libpath)
{
    if (g_handle) return 0;
    g_handle = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!g_handle) {
        snprintf(g_desc, sizeof g_desc, "dlopen failed: %s", dlerror());
        return -1;
    }
    g_getrf = (getrf_fn)sym2("scipy_dgetrf_", "dgetrf_");
    g_getrs = (getrs_fn)sym2("scipy_dgetrs_", "dgetrs_");
    g_gesv = (gesv_fn)sym2("scipy_dgesv_", "dgesv_");
    g_larnv = (larnv_fn)sym2("scipy_dlarnv_", "dlarnv_");
    g_setthr = (setthr_fn)sym2("scipy_openblas_set_num_threads", "openblas_set_num_threads");
    if (!g_getrf || !g_getrs || !g_gesv || !g_larnv) {
        snprintf(g_desc, sizeof g_desc, "LAPACK symbols missing in %s", libpath);
        return -2;
    }
    /* BLAS threads pinned to 1: parallelism is the OpenMP loop, as in the testers */
    if (g_setthr) g_setthr(1);
    typedef char *(*cfg_fn)(void);
    cfg_fn cfg = (cfg_fn)sym2("scipy_openblas_get_config", "openblas_get_config");
    snprintf(g_desc, sizeof g_desc, "%s", cfg ? cfg() : libpath);
    return 0;
}

const char *lapack_loop_describe(void) { return g_desc; }

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the benchmark's reference arm asks for the box's cores back */
void lapack_loop_set_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

int lapack_loop_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void lapack_dlarnv(int idist, int *iseed, long n, double *x)
{
    /* dlarnv takes an int count: feed it in chunks, the seed carries the stream */
    const long chunk = 1L << 30;
    for (long off = 0; off < n; off += chunk) {
        int c = (int)((n - off) < chunk ? (n - off) : chunk);
        g_larnv(&idist, iseed, &c, x + off);
    }
}

static double now(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0;
#endif
}

/* each returns the wall time of the loop in seconds */
double lapack_dgetrf_loop(int m, int n, double *A, int lda, long strideA, int *ipiv,
                          long stride_ipiv, int *info, long batch)
{
    double t0 = now();
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < batch; ++b)
        g_getrf(&m, &n, A + b * strideA, &lda, ipiv + b * stride_ipiv, info + b);
    return now() - t0;
}

double lapack_dgetrs_loop(int trans, int n, int nrhs, const double *A, int lda, long strideA,
                          const int *ipiv, long stride_ipiv, double *B, int ldb, long strideB,
                          long batch)
{
    const char *t = trans == 111 ? "N" : "T";
    double t0 = now();
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < batch; ++b) {
        int inf;
        g_getrs(t, &n, &nrhs, A + b * strideA, &lda, ipiv + b * stride_ipiv, B + b * strideB, &ldb,
                &inf, 1);
    }
    return now() - t0;
}

double lapack_dgesv_loop(int n, int nrhs, double *A, int lda, long strideA, int *ipiv,
                         long stride_ipiv, double *B, int ldb, long strideB, int *info, long batch)
{
    double t0 = now();
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < batch; ++b)
        g_gesv(&n, &nrhs, A + b * strideA, &lda, ipiv + b * stride_ipiv, B + b * strideB, &ldb,
               info + b);
    return now() - t0;
}

double lapack_dgetrf_vloop(const int *m, const int *n, double *A, const int *lda,
                           const long *offsets, int *ipiv, const long *ipiv_off, int *info,
                           long batch)
{
    double t0 = now();
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < batch; ++b)
        g_getrf(m + b, n + b, A + offsets[b], lda + b, ipiv + ipiv_off[b], info + b);
    return now() - t0;
}
