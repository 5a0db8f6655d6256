// Source-compatibility entry points: the panel-level and BLAS-level pieces of the reference's batched LU that are
// declared in its public header (include/magma_zbatched.h:330-335,472-478,829-855, z -> d) and that callers of the
// reference may link against directly. Here they are thin fronts over the same kernels the drivers use: a panel
// factorisation at an offset is just an LU of the displaced sub-matrix (pivots relative to the panel, info offset
// by gbstep), so nothing is re-derived.
#include "lu_common.cuh"

using namespace mb200;

namespace {

inline int imax(int a, int b) { return a > b ? a : b; }

__global__ void merge_info_kernel(int *__restrict__ info, const int *__restrict__ tinfo, int gbstep, long batch)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    // zgetf2_fused_kernel_batched (magmablas/zgetf2_kernels.cu:840-866): linfo starts from info_array unless
    // gbstep == 0; a zero pivot at panel step i records gbstep + i + 1 only if nothing was recorded before
    const int prev = (gbstep == 0) ? 0 : info[b];
    const int t = tinfo[b];
    info[b] = prev != 0 ? prev : (t != 0 ? t + gbstep : 0);
}

// new[r] = old[pivinfo[r] - 1] for the rows r in [0, height) (written to dout) and for the rows those came from
// (written in place), as zlaswp_rowparallel_devfunc does (magmablas/zlaswp_device.cuh:25-84) -- but with every read
// before the first write. One CTA per (matrix, 8-column group), thread = one of the `height` top rows.
constexpr int RP_COLS = 8;
__global__ void __launch_bounds__(1024)
laswp_rowparallel_kernel(int n, double **__restrict__ din, int ii, int ij, int ldi, double **__restrict__ dout, int oi, int oj,
                         int ldo, int height, int **__restrict__ pivinfo, int groups)
{
    const long b = blockIdx.x / groups;
    const int c0 = (blockIdx.x % groups) * RP_COLS;
    const int w = (n - c0) < RP_COLS ? (n - c0) : RP_COLS;
    const int tid = threadIdx.x;
    double *A = din[b] + (size_t)ij * ldi + ii + (size_t)c0 * ldi;
    double *O = dout[b] + (size_t)oj * ldo + oi + (size_t)c0 * ldo;
    const int *piv = pivinfo[b];
    double v1[RP_COLS], v2[RP_COLS];
    int r1 = 0;
    const bool live = tid < height;
    if (live) {
        r1 = piv[tid] - 1;
        const int r2 = piv[r1] - 1;
#pragma unroll
        for (int i = 0; i < RP_COLS; ++i) {
            if (i < w) {
                v1[i] = A[r1 + (size_t)i * ldi];
                v2[i] = A[r2 + (size_t)i * ldi];
            }
        }
    }
    __syncthreads();
    if (live) {
#pragma unroll
        for (int i = 0; i < RP_COLS; ++i)
            if (i < w) A[r1 + (size_t)i * ldi] = v2[i];
    }
    __syncthreads();
    if (live) {
#pragma unroll
        for (int i = 0; i < RP_COLS; ++i)
            if (i < w) O[tid + (size_t)i * ldo] = v1[i];
    }
}

// x <- op(A)^-1 x, one CTA (128 threads) per system. 32-wide diagonal blocks: the block is solved by warp 0 with one
// shuffle per unknown (lane = row), the rest of the vector then receives the block's contribution from all threads
// (NoTrans: thread = row, coalesced column reads; Trans: thread = column of the transposed product, a warp per dot product).
constexpr int TV_THREADS = 128;
__global__ void __launch_bounds__(TV_THREADS)
trsv_kernel(int uplo, int trans, int diag, int n, double **__restrict__ dA, int ldda, double **__restrict__ dB, int incb)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *x = reinterpret_cast<double *>(smem_raw);  // [n]
    const long b = blockIdx.x;
    const double *__restrict__ A = dA[b];
    double *__restrict__ X = dB[b];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < n; i += TV_THREADS) x[i] = X[(size_t)i * incb];
    __syncthreads();
    const bool lower = (uplo == MagmaLower);
    const bool notrans = (trans == MagmaNoTrans);
    // effective triangle of op(A): lower-NoTrans and upper-Trans sweep forward, the others backward
    const bool forward = (lower == notrans);
    const int nblk = (n + 31) / 32;
    for (int bi = 0; bi < nblk; ++bi) {
        const int kb = forward ? bi : nblk - 1 - bi;
        const int r0 = 32 * kb;
        const int nb = (n - r0) < 32 ? (n - r0) : 32;
        if (wid == 0) {
            // op(A)(r0+i, r0+k): element (i,k) of the diagonal block of op(A)
            double xi = lane < nb ? x[r0 + lane] : 0.0;
            for (int s = 0; s < nb; ++s) {
                const int k = forward ? s : nb - 1 - s;
                // finish unknown k
                double xk = __shfl_sync(0xffffffffu, xi, k);
                if (diag == MagmaNonUnit) {
                    const double dkk = A[(size_t)(r0 + k) + (size_t)(r0 + k) * ldda];
                    xk = xk / dkk;
                }
                if (lane == k) xi = xk;
                const bool rest = forward ? (lane > k) : (lane < k);
                if (rest && lane < nb) {
                    const double aik = notrans ? A[(size_t)(r0 + lane) + (size_t)(r0 + k) * ldda]
                                               : A[(size_t)(r0 + k) + (size_t)(r0 + lane) * ldda];
                    xi = fma(-aik, xk, xi);
                }
            }
            if (lane < nb) x[r0 + lane] = xi;
        }
        __syncthreads();
        // the other rows: x(i) -= sum_k op(A)(i, r0+k) x(r0+k), k increasing
        const int lo = forward ? r0 + nb : 0, hi = forward ? n : r0;
        if (notrans) {
            for (int i = lo + tid; i < hi; i += TV_THREADS) {
                double acc = x[i];
                for (int k = 0; k < nb; ++k) acc = fma(-A[(size_t)i + (size_t)(r0 + k) * ldda], x[r0 + k], acc);
                x[i] = acc;
            }
        } else {
            // op(A)(i, r0+k) = A(r0+k, i): column i of A, rows r0..r0+nb-1 -- one warp per i, lanes over k
            for (int i = lo + wid; i < hi; i += TV_THREADS / 32) {
                double p = (lane < nb) ? A[(size_t)(r0 + lane) + (size_t)i * ldda] * x[r0 + lane] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) x[i] -= p;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += TV_THREADS) X[(size_t)i * incb] = x[i];
}

// LU of the m x n block at (ai, aj) of every matrix: pivots 1-based relative to row ai, written at ipiv[b] + ai;
// info merged with the caller's running value the way the reference's panels do.
magma_int_t panel_lu_offset(const char *func, int m, int n, double **dA_array, int ai, int aj, int ldda, int **ipiv_array,
                            int *info_array, int gbstep, long batch, magma_queue_t queue)
{
    if (m == 0 || n == 0 || batch <= 0) return 0;
    cudaStream_t s = MB200_Q(queue)->stream;
    const size_t bytes = (size_t)batch * (sizeof(double *) + sizeof(int *) + sizeof(int));
    char *scr = (char *)queue_dscratch(queue, bytes, 0);
    if (!scr) {
        magma_xerbla(func, -MAGMA_ERR_DEVICE_ALLOC);
        return MAGMA_ERR_DEVICE_ALLOC;
    }
    double **pA = (double **)scr;
    int **pP = (int **)(pA + batch);
    int *tinfo = (int *)(pP + batch);
    displace_pointers_launch((void **)pA, (void **)dA_array, sizeof(double), ldda, ai, aj, batch, s);
    displace_pointers_launch((void **)pP, (void **)ipiv_array, sizeof(int), 1, ai, 0, batch, s);
    const magma_int_t rc = magma_dgetrf_batched(m, n, pA, ldda, pP, tinfo, (magma_int_t)batch, queue);
    if (rc != 0) return rc;
    merge_info_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>(info_array, tinfo, gbstep, batch);
    count_launch();
    MB200_CHECK_LAUNCH("merge_info_kernel");
    return 0;
}

}  // namespace

extern "C" {

// magmablas/zgetf2_kernels.cu:1005-1058: m x n panel, n <= 32, at (ai, aj); gbstep = aj.
magma_int_t magma_dgetf2_fused_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ai, magma_int_t aj,
                                       magma_int_t ldda, magma_int_t **dipiv_array, magma_int_t *info_array,
                                       magma_int_t batchCount, magma_queue_t queue)
{
    if (m < 0) return -1;
    if (n < 0 || n > 32) {
        fprintf(stderr, "%s: n = %4lld not supported, must be between 0 and %4lld\n", __func__, (long long)n, (long long)32);
        return -2;
    }
    return panel_lu_offset(__func__, m, n, dA_array, ai, aj, ldda, dipiv_array, info_array, aj, batchCount, queue);
}

// src/zgetf2_batched.cpp:243-287 (dpivinfo_array is scratch of the reference's implementation; unused here)
magma_int_t magma_dgetf2_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ai, magma_int_t aj,
                                 magma_int_t lda, magma_int_t **ipiv_array, magma_int_t **dpivinfo_array,
                                 magma_int_t *info_array, magma_int_t gbstep, magma_int_t batchCount, magma_queue_t queue)
{
    (void)dpivinfo_array;
    magma_int_t arginfo = 0;
    if (m < 0) arginfo = -1;
    else if (n < 0) arginfo = -2;
    else if (ai < 0) arginfo = -4;
    else if (aj < 0 || aj != ai) arginfo = -5;
    else if (lda < imax(1, m)) arginfo = -6;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    return panel_lu_offset(__func__, m, n, dA_array, ai, aj, lda, ipiv_array, info_array, gbstep, batchCount, queue);
}

// src/zgetrf_panel_batched.cpp:101-196 (min_recpnb and dpivinfo_array steer the reference's recursion; unused here)
magma_int_t magma_dgetrf_recpanel_batched(magma_int_t m, magma_int_t n, magma_int_t min_recpnb, double **dA_array,
                                          magma_int_t ai, magma_int_t aj, magma_int_t ldda, magma_int_t **dipiv_array,
                                          magma_int_t **dpivinfo_array, magma_int_t *info_array, magma_int_t gbstep,
                                          magma_int_t batchCount, magma_queue_t queue)
{
    (void)min_recpnb; (void)dpivinfo_array;
    magma_int_t arginfo = 0;
    if (m < 0) arginfo = -1;
    else if (n < 0) arginfo = -2;
    else if (ai < 0) arginfo = -4;
    else if (aj < 0 || aj != ai) arginfo = -5;
    else if (ldda < imax(1, m)) arginfo = -6;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    return panel_lu_offset(__func__, m, n, dA_array, ai, aj, ldda, dipiv_array, info_array, gbstep, batchCount, queue);
}

// magmablas/zlaswp_batched.cu:47-87
void magma_dlaswp_rowparallel_batched(magma_int_t n, double **input_array, magma_int_t input_i, magma_int_t input_j,
                                      magma_int_t ldi, double **output_array, magma_int_t output_i, magma_int_t output_j,
                                      magma_int_t ldo, magma_int_t k1, magma_int_t k2, magma_int_t **pivinfo_array,
                                      magma_int_t batchCount, magma_queue_t queue)
{
    if (n == 0 || batchCount <= 0) return;
    const int height = k2 - k1;
    if (height <= 0) return;
    if (height > 1024) {
        fprintf(stderr, "%s: n=%lld > 1024, not supported\n", __func__, (long long)n);
        magma_xerbla(__func__, -MAGMA_ERR_NOT_SUPPORTED);
        return;
    }
    const int groups = (n + RP_COLS - 1) / RP_COLS;
    const int threads = ((height + 31) / 32) * 32;
    const long per = 0x7fffffffL / groups;
    for (long off = 0; off < batchCount; off += per) {
        const long cnt = batchCount - off < per ? batchCount - off : per;
        laswp_rowparallel_kernel<<<(unsigned)(cnt * groups), threads, 0, MB200_Q(queue)->stream>>>(
            n, input_array + off, input_i, input_j, ldi, output_array + off, output_i, output_j, ldo, height,
            pivinfo_array + off, groups);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("laswp_rowparallel_kernel");
    }
}

// magmablas/ztrsv_batched.cu:258-298: x_b <- op(A_b)^-1 x_b in place, x stored with increment incb
void magmablas_dtrsv_batched(magma_uplo_t uplo, magma_trans_t transA, magma_diag_t diag, magma_int_t n, double **dA_array,
                             magma_int_t ldda, double **dB_array, magma_int_t incb, magma_int_t batchCount,
                             magma_queue_t queue)
{
    magma_int_t info = 0;
    if (uplo != MagmaUpper && uplo != MagmaLower) info = -1;
    else if (transA != MagmaNoTrans && transA != MagmaTrans && transA != MagmaConjTrans) info = -2;
    else if (diag != MagmaUnit && diag != MagmaNonUnit) info = -3;
    else if (n < 0) info = -4;
    else if (ldda < imax(1, n)) info = -6;
    else if (incb <= 0) info = -8;  // positive increments only
    else if (n > 28000) info = -MAGMA_ERR_NOT_SUPPORTED;  // x lives in shared memory
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return;
    }
    if (n == 0 || batchCount <= 0) return;
    const size_t smem = sizeof(double) * (size_t)n;
    static DevOnce once;
    smem_optin(once, trsv_kernel, 227 * 1024);
    const int t = (transA == MagmaNoTrans) ? MagmaNoTrans : MagmaTrans;
    for (long off = 0; off < batchCount; off += 0x7fffffffL) {
        const long cnt = batchCount - off < 0x7fffffffL ? batchCount - off : 0x7fffffffL;
        trsv_kernel<<<(unsigned)cnt, TV_THREADS, smem, MB200_Q(queue)->stream>>>(uplo, t, diag, n, dA_array + off, ldda, dB_array + off,
                                                                       incb);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("trsv_kernel");
    }
}

}  // extern "C"
