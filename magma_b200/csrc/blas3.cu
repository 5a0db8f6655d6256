// Standalone batched BLAS-3 behind the reference names: magma_dgemm_batched[_core]
// (magmablas/zgemm_batched.cpp:49-100,269-286) on the FP64 tensor pipe for all four transpose combinations, and
// magmablas_dtrsm_batched side = Right (magmablas/ztrsm_batched_core.cpp:299-350). Plus the strided device front
// ends of section 8(f).4: dA + b*stride instead of pointer arrays.
//
// gemm_dmma_tt_kernel: one 64 x 64 tile of C per CTA of four warps (warp tile 32 x 32 = 4 x 4 DMMA.8x8x4 fragments),
// op(A) and op(B) staged 16 k at a time into padded, conflict-free shared layouts (As[k][r], Bs[c][k]); the next
// chunk is fetched into registers while the current one is multiplied. Accumulation order per element: k increasing,
// four products per DMMA chained in order (tools/dmma_probe.cu), so with alpha = -1, beta = 1 the result equals the
// chain c <- fma(-a(i,k), b(k,j), c) of the LU trailing update bit for bit (the kernel then starts from C and feeds
// -A); otherwise acc = sum_k a b from zero and C <- fma(alpha, acc, beta C)  (beta = 0: C is not read).
#include "lu_common.cuh"

using namespace mb200;

namespace {

inline int imax(int a, int b) { return a > b ? a : b; }

constexpr int GT = 64, GK = 16, G_THREADS = 128;
constexpr int G_LDA = GT + 4;  // As[kk*G_LDA + r]: A fragments (k = 4s+q, row 8i+g) hit 32 distinct banks
constexpr int G_LDB = GK + 4;  // Bs[c*G_LDB + kk]

__device__ __forceinline__ void g_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(G_THREADS)
gemm_dmma_tt_kernel(int m, int n, int k, double alpha, double const *const *__restrict__ dA, int Ai, int Aj, int ldda,
                    double const *const *__restrict__ dB, int Bi, int Bj, int lddb, double beta, double **__restrict__ dC,
                    int Ci, int Cj, int lddc, int mt, int nt)
{
    __shared__ __align__(16) double As[GK * G_LDA];
    __shared__ __align__(16) double Bs[GT * G_LDB];
    const long b = blockIdx.x / (mt * nt);
    const int t = blockIdx.x % (mt * nt);
    const int r0 = (t % mt) * GT, c0 = (t / mt) * GT;
    const double *__restrict__ A = dA[b] + Ai + (size_t)Aj * ldda;
    const double *__restrict__ B = dB[b] + Bi + (size_t)Bj * lddb;
    double *__restrict__ C = dC[b] + Ci + (size_t)Cj * lddc;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wr = w & 1, wc = w >> 1;
    const bool lu_mode = (alpha == -1.0 && beta == 1.0);  // accumulate into C itself, A negated: the LU update's chain

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jt = 0; jt < 4; ++jt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = r0 + wr * 32 + 8 * i + g, c = c0 + wc * 32 + 8 * jt + 2 * q + e;
                acc[i][jt][e] = (lu_mode && r < m && c < n) ? C[(size_t)r + (size_t)c * lddc] : 0.0;
            }

    // staging assignment: 64 x 16 elements of op(A) and of op(B) per chunk, 8 + 8 per thread; the fast index of the
    // global read follows the operand's storage order
    double ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * G_THREADS;
            // op(A)(r, kk): NoTrans A[r + kk*lda] (r fast), Trans A[kk + r*lda] (kk fast)
            const int ar = TA ? idx / GK : idx % GT, ak = TA ? idx % GK : idx / GT;
            const bool aok = (r0 + ar < m) && (k0 + ak < k);
            ra[u] = aok ? (TA ? A[(size_t)(k0 + ak) + (size_t)(r0 + ar) * ldda] : A[(size_t)(r0 + ar) + (size_t)(k0 + ak) * ldda]) : 0.0;
            // op(B)(kk, c): NoTrans B[kk + c*ldb] (kk fast), Trans B[c + kk*ldb] (c fast)
            const int bc = TB ? idx % GT : idx / GK, bk = TB ? idx / GT : idx % GK;
            const bool bok = (c0 + bc < n) && (k0 + bk < k);
            rb[u] = bok ? (TB ? B[(size_t)(c0 + bc) + (size_t)(k0 + bk) * lddb] : B[(size_t)(k0 + bk) + (size_t)(c0 + bc) * lddb]) : 0.0;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * G_THREADS;
            const int ar = TA ? idx / GK : idx % GT, ak = TA ? idx % GK : idx / GT;
            As[ak * G_LDA + ar] = lu_mode ? -ra[u] : ra[u];
            const int bc = TB ? idx % GT : idx / GK, bk = TB ? idx / GT : idx % GK;
            Bs[bc * G_LDB + bk] = rb[u];
        }
    };
    if (k > 0) fetch(0);
    for (int k0 = 0; k0 < k; k0 += GK) {
        __syncthreads();  // the previous chunk has been consumed
        stage();
        __syncthreads();
        if (k0 + GK < k) fetch(k0 + GK);
        const double *Ap = As + q * G_LDA + wr * 32 + g;
        const double *Bp = Bs + (wc * 32 + g) * G_LDB + q;
#pragma unroll
        for (int ks = 0; ks < GK / 4; ++ks) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = Ap[ks * 4 * G_LDA + 8 * i];
#pragma unroll
            for (int jt = 0; jt < 4; ++jt) bf[jt] = Bp[8 * jt * G_LDB + ks * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jt = 0; jt < 4; ++jt) g_dmma(acc[i][jt][0], acc[i][jt][1], af[i], bf[jt]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jt = 0; jt < 4; ++jt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = r0 + wr * 32 + 8 * i + g, c = c0 + wc * 32 + 8 * jt + 2 * q + e;
                if (r < m && c < n) {
                    double *dst = C + (size_t)r + (size_t)c * lddc;
                    if (lu_mode) *dst = acc[i][jt][e];
                    else *dst = (beta == 0.0) ? alpha * acc[i][jt][e] : fma(alpha, acc[i][jt][e], beta * (*dst));
                }
            }
}

// X op(A) = alpha B, X overwrites B (m x n), A is n x n. One thread per row of B (rows are independent), columns of
// X in the order the triangle dictates, 32 at a time in registers; a(k,j) is the same address for every thread of
// the CTA (a broadcast through L1), the row elements b(i, j) are coalesced across threads.
// j-th column of op(A): x_j = (alpha b_j - sum_{k solved} x_k opA(k,j)) / opA(j,j), accumulated with k in solve order.
constexpr int TR_THREADS = 128, TR_BLK = 32;
template <bool UP>  // op(A) upper triangular (columns solved in ascending order) or lower (descending)
__global__ void __launch_bounds__(TR_THREADS)
trsm_right_kernel(int trans, int diag, int m, int n, double alpha, double **__restrict__ dA, int ldda,
                  double **__restrict__ dB, int lddb, int row_tiles)
{
    const long b = blockIdx.x / row_tiles;
    const int i = (blockIdx.x % row_tiles) * TR_THREADS + threadIdx.x;
    if (i >= m) return;
    const double *__restrict__ A = dA[b];
    double *__restrict__ B = dB[b] + i;
    const bool nt = (trans == MagmaNoTrans), unit = (diag == MagmaUnit);
    // X opA = B: column j of X needs the columns k with opA(k,j) != 0, k != j. opA upper (upper-NoTrans or lower-Trans):
    // k < j, sweep j ascending; opA lower: k > j, sweep descending.
    constexpr bool up = UP;
    auto opA = [&](int kk, int jj) -> double { return nt ? A[(size_t)kk + (size_t)jj * ldda] : A[(size_t)jj + (size_t)kk * ldda]; };
    const int nblk = (n + TR_BLK - 1) / TR_BLK;
    for (int bi = 0; bi < nblk; ++bi) {
        const int jb0 = (up ? bi : nblk - 1 - bi) * TR_BLK;
        const int jw = (n - jb0) < TR_BLK ? (n - jb0) : TR_BLK;
        double x[TR_BLK];
#pragma unroll
        for (int jj = 0; jj < TR_BLK; ++jj) x[jj] = (jj < jw) ? alpha * B[(size_t)(jb0 + jj) * lddb] : 0.0;
        // contributions of the columns solved in earlier blocks (already final in B)
        if (up) {
            for (int kk = 0; kk < jb0; ++kk) {
                const double xk = B[(size_t)kk * lddb];
#pragma unroll
                for (int jj = 0; jj < TR_BLK; ++jj)
                    if (jj < jw) x[jj] = fma(-xk, opA(kk, jb0 + jj), x[jj]);
            }
        } else {
            for (int kk = n - 1; kk >= jb0 + jw; --kk) {
                const double xk = B[(size_t)kk * lddb];
#pragma unroll
                for (int jj = 0; jj < TR_BLK; ++jj)
                    if (jj < jw) x[jj] = fma(-xk, opA(kk, jb0 + jj), x[jj]);
            }
        }
        // the diagonal block
#pragma unroll
        for (int s = 0; s < TR_BLK; ++s) {
            const int jj = up ? s : TR_BLK - 1 - s;
            if (jj < jw) {
                if (!unit) x[jj] = x[jj] / opA(jb0 + jj, jb0 + jj);
#pragma unroll
                for (int t2 = 0; t2 < TR_BLK; ++t2) {
                    const bool later = up ? (t2 > jj) : (t2 < jj);
                    if (later && t2 < jw) x[t2] = fma(-x[jj], opA(jb0 + jj, jb0 + t2), x[t2]);
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < TR_BLK; ++jj)
            if (jj < jw) B[(size_t)(jb0 + jj) * lddb] = x[jj];
    }
}

}  // namespace

namespace mb200 {

void gemm_dmma_launch(int transA, int transB, int m, int n, int k, double alpha, double const *const *dA, int Ai, int Aj,
                      int ldda, double const *const *dB, int Bi, int Bj, int lddb, double beta, double **dC, int Ci, int Cj,
                      int lddc, long batch, cudaStream_t s)
{
    if (m <= 0 || n <= 0 || batch <= 0) return;
    const int mt = (m + GT - 1) / GT, nt = (n + GT - 1) / GT;
    const bool ta = (transA != MagmaNoTrans), tb = (transB != MagmaNoTrans);
    const long per = 0x7fffffffL / ((long)mt * nt);
    for (long off = 0; off < batch; off += per) {
        const long cnt = batch - off < per ? batch - off : per;
        const unsigned grid = (unsigned)(cnt * mt * nt);
#define MB200_GEMM(TA_, TB_)                                                                                              \
    gemm_dmma_tt_kernel<TA_, TB_><<<grid, G_THREADS, 0, s>>>(m, n, k, alpha, dA + off, Ai, Aj, ldda, dB + off, Bi, Bj, lddb, \
                                                            beta, dC + off, Ci, Cj, lddc, mt, nt)
        if (!ta && !tb) MB200_GEMM(false, false);
        else if (ta && !tb) MB200_GEMM(true, false);
        else if (!ta && tb) MB200_GEMM(false, true);
        else MB200_GEMM(true, true);
#undef MB200_GEMM
        count_launch();
        MB200_CHECK_LAUNCH_VOID("gemm_dmma_tt_kernel");
    }
}

void trsm_right_launch(int uplo, int trans, int diag, int m, int n, double alpha, double **dA, int ldda, double **dB,
                       int lddb, long batch, cudaStream_t s)
{
    if (m <= 0 || n <= 0 || batch <= 0) return;
    const int row_tiles = (m + TR_THREADS - 1) / TR_THREADS;
    const long per = 0x7fffffffL / row_tiles;
    for (long off = 0; off < batch; off += per) {
        const long cnt = batch - off < per ? batch - off : per;
        const int t = (trans == MagmaNoTrans) ? MagmaNoTrans : MagmaTrans;
        if ((uplo == MagmaLower) != (t == MagmaNoTrans))
            trsm_right_kernel<true><<<(unsigned)(cnt * row_tiles), TR_THREADS, 0, s>>>(t, diag, m, n, alpha, dA + off, ldda, dB + off, lddb, row_tiles);
        else
            trsm_right_kernel<false><<<(unsigned)(cnt * row_tiles), TR_THREADS, 0, s>>>(t, diag, m, n, alpha, dA + off, ldda, dB + off, lddb, row_tiles);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("trsm_right_kernel");
    }
}

}  // namespace mb200

extern "C" {

// magmablas/zgemm_batched.cpp:269-286
void magma_dgemm_batched(magma_trans_t transA, magma_trans_t transB, magma_int_t m, magma_int_t n, magma_int_t k,
                         double alpha, double const *const *dA_array, magma_int_t ldda, double const *const *dB_array,
                         magma_int_t lddb, double beta, double **dC_array, magma_int_t lddc, magma_int_t batchCount,
                         magma_queue_t queue)
{
    magma_dgemm_batched_core(transA, transB, m, n, k, alpha, dA_array, 0, 0, ldda, dB_array, 0, 0, lddb, beta, dC_array, 0, 0,
                             lddc, batchCount, queue);
}

// ---- strided device front ends: matrix b at dA + b*strideA, pivots at dipiv + b*stride_piv ----------
// New surface (the reference has strided forms for gemm / gbtrf / gbsv only: include/magma_zbatched.h:135,538-622).
// The pointer arrays the kernels consume are built in per-queue scratch by magma_dset_pointer's kernel.
static int strided_ptrs(magma_queue_t queue, long batch, double *dA, long lda_elems, double ***pA, int *dipiv, long pstride,
                        int ***pP, double *dB, long ldb_elems, double ***pB)
{
    const size_t bytes = (size_t)batch * 3 * sizeof(void *);
    char *scr = (char *)queue_dscratch(queue, bytes, 0);
    if (!scr) return MAGMA_ERR_DEVICE_ALLOC;
    cudaStream_t s = MB200_Q(queue)->stream;
    *pA = (double **)scr;
    *pP = (int **)(scr + (size_t)batch * sizeof(void *));
    *pB = (double **)(scr + (size_t)batch * 2 * sizeof(void *));
    set_pointer_launch((void **)*pA, (char *)dA, sizeof(double), 1, 0, 0, lda_elems, batch, s);
    if (dipiv) set_pointer_launch((void **)*pP, (char *)dipiv, sizeof(int), 1, 0, 0, pstride, batch, s);
    if (dB) set_pointer_launch((void **)*pB, (char *)dB, sizeof(double), 1, 0, 0, ldb_elems, batch, s);
    return 0;
}

magma_int_t magma_dgetrf_batched_strided(magma_int_t m, magma_int_t n, double *dA, magma_int_t ldda, magma_int_t strideA,
                                         magma_int_t *dipiv, magma_int_t stride_piv, magma_int_t *info_array,
                                         magma_int_t batchCount, magma_queue_t queue)
{
    if (batchCount <= 0) return 0;
    double **pA, **pB;
    int **pP;
    if (strided_ptrs(queue, batchCount, dA, strideA, &pA, dipiv, stride_piv, &pP, nullptr, 0, &pB) != 0) return MAGMA_ERR_DEVICE_ALLOC;
    return magma_dgetrf_batched(m, n, pA, ldda, pP, info_array, batchCount, queue);
}

magma_int_t magma_dgetrs_batched_strided(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, double *dA, magma_int_t ldda,
                                         magma_int_t strideA, magma_int_t *dipiv, magma_int_t stride_piv, double *dB,
                                         magma_int_t lddb, magma_int_t strideB, magma_int_t batchCount, magma_queue_t queue)
{
    if (batchCount <= 0) return 0;
    double **pA, **pB;
    int **pP;
    if (strided_ptrs(queue, batchCount, dA, strideA, &pA, dipiv, stride_piv, &pP, dB, strideB, &pB) != 0) return MAGMA_ERR_DEVICE_ALLOC;
    return magma_dgetrs_batched(trans, n, nrhs, pA, ldda, pP, pB, lddb, batchCount, queue);
}

magma_int_t magma_dgesv_batched_strided(magma_int_t n, magma_int_t nrhs, double *dA, magma_int_t ldda, magma_int_t strideA,
                                        magma_int_t *dipiv, magma_int_t stride_piv, double *dB, magma_int_t lddb,
                                        magma_int_t strideB, magma_int_t *dinfo_array, magma_int_t batchCount,
                                        magma_queue_t queue)
{
    if (batchCount <= 0) return 0;
    double **pA, **pB;
    int **pP;
    if (strided_ptrs(queue, batchCount, dA, strideA, &pA, dipiv, stride_piv, &pP, dB, strideB, &pB) != 0) return MAGMA_ERR_DEVICE_ALLOC;
    return magma_dgesv_batched(n, nrhs, pA, ldda, pP, pB, lddb, dinfo_array, batchCount, queue);
}

}  // extern "C"
