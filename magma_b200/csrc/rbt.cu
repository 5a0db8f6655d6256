// Solve with a random butterfly transformation instead of pivoting.
// Replaces src/zgesv_rbt_batched.cpp, src/zgerbt_batched.cpp and magmablas/zgerbt_func_batched.cu / zgerbt_kernels.cu
// (z -> d): A <- U^T A V with two-level recursive butterflies U, V (each stored as 2n scalars: [0,n) the outer level,
// [n,2n) the two half-size inner butterflies), B <- U^T B, LU without pivoting, two triangular solves, X <- V Y.
// The factorisation and the solves are this library's no-pivoting paths (left-looking slab driver + getrs kernels);
// the butterflies are pure streaming kernels: the matrix kernel applies BOTH levels per launch pair, one thread per
// 2x2 element group, with the reference's operation order and no FMA contraction (products and sums rounded
// separately: oracle.prbt_* in oracle/__init__.py restates exactly this arithmetic).
#include <stdlib.h>

#include "lu_common.cuh"

using namespace mb200;

namespace {

inline int imax(int a, int b) { return a > b ? a : b; }

// One butterfly level on the Am x An block at (Ai, Aj) of every matrix (magmablas_zelementary_multiplication_devfunc,
// magmablas/zgerbt_kernels.cu:21-80): with r1 = ceil(Am/2), c1 = ceil(An/2) and (i, j) in the top-left r1 x c1 part,
//   [a00 a01; a10 a11] = A(i | i+r1, j | j+c1)  (missing elements = 0)
//   A00 = u1 v1 ((a00+a01) + (a10+a11)),  A01 = u1 v2 ((a00-a01) + (a10-a11)),
//   A10 = u2 v1 ((a00+a01) - (a10+a11)),  A11 = u2 v2 ((a00-a01) - (a10-a11)),   u1 = u(i), u2 = u(i+r1), v likewise.
// blockIdx.y selects the block (level 1: the four quadrants; level 2: the whole matrix), blockIdx.x = (matrix, tile).
struct RbtBlock {
    int Am, An, Ai, Aj, Ui, Vi;
};
struct RbtBlocks {
    RbtBlock b[4];
};

constexpr int RBT_TX = 32, RBT_TY = 8;

__global__ void __launch_bounds__(RBT_TX *RBT_TY)
prbt_level_kernel(RbtBlocks blocks, double **__restrict__ dA, int ldda, const double *__restrict__ du,
                  const double *__restrict__ dv, int tiles_x, int tiles_y)
{
    const RbtBlock &k = blocks.b[blockIdx.y];
    const int tiles = tiles_x * tiles_y;
    const long mat = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const int r1 = (k.Am + 1) / 2, r2 = k.Am - r1, c1 = (k.An + 1) / 2, c2 = k.An - c1;
    const int i = (t % tiles_x) * RBT_TX + threadIdx.x;
    const int j = (t / tiles_x) * RBT_TY + threadIdx.y;
    if (i >= r1 || j >= c1) return;
    double *A = dA[mat] + (size_t)k.Aj * ldda + k.Ai + i + (size_t)j * ldda;
    const bool v01 = j < c2, v10 = i < r2, v11 = v01 && v10;
    const double a00 = A[0];
    const double a01 = v01 ? A[(size_t)ldda * c1] : 0.0;
    const double a10 = v10 ? A[r1] : 0.0;
    const double a11 = v11 ? A[(size_t)ldda * c1 + r1] : 0.0;
    const double u1 = du[k.Ui + i], v1 = dv[k.Vi + j];
    const double u2 = v10 ? du[k.Ui + r1 + i] : 0.0, v2 = v01 ? dv[k.Vi + c1 + j] : 0.0;
    const double b1 = __dadd_rn(a00, a01), b2 = __dadd_rn(a10, a11), b3 = __dsub_rn(a00, a01), b4 = __dsub_rn(a10, a11);
    A[0] = __dmul_rn(__dmul_rn(u1, v1), __dadd_rn(b1, b2));
    if (v01) A[(size_t)ldda * c1] = __dmul_rn(__dmul_rn(u1, v2), __dadd_rn(b3, b4));
    if (v10) A[r1] = __dmul_rn(__dmul_rn(u2, v1), __dsub_rn(b1, b2));
    if (v11) A[(size_t)ldda * c1 + r1] = __dmul_rn(__dmul_rn(u2, v2), __dsub_rn(b3, b4));
}

// x <- V x on rows [off, off+n) of every right-hand side (magmablas_zapply_vector_devfunc, zgerbt_kernels.cu:105-125)
__global__ void __launch_bounds__(256)
prbt_mv_level_kernel(int n, int nrhs, const double *__restrict__ dv, int offv, double **__restrict__ dB, int lddb, int off,
                     int tiles)
{
    if (n < 1) return;
    const long mat = blockIdx.x / tiles;
    const int idx = (blockIdx.x % tiles) * 256 + threadIdx.x;
    const int n1 = (n + 1) / 2, n2 = n - n1;
    if (idx >= n1) return;
    const double u0 = dv[offv + idx], u1 = idx < n2 ? dv[offv + n1 + idx] : 0.0;
    for (int c = 0; c < nrhs; ++c) {
        double *b = dB[mat] + (size_t)c * lddb + off + idx;
        const double a1 = __dmul_rn(u0, b[0]);
        const double a2 = idx < n2 ? __dmul_rn(u1, b[n1]) : 0.0;
        b[0] = __dadd_rn(a1, a2);
        if (idx < n2) b[n1] = __dsub_rn(a1, a2);
    }
}

// x <- U^T x (magmablas_zapply_transpose_vector_devfunc, zgerbt_kernels.cu:150-170; nothing to do for n < 2)
__global__ void __launch_bounds__(256)
prbt_mtv_level_kernel(int n, int nrhs, const double *__restrict__ du, int offu, double **__restrict__ dB, int lddb, int off,
                      int tiles)
{
    if (n < 2) return;
    const long mat = blockIdx.x / tiles;
    const int idx = (blockIdx.x % tiles) * 256 + threadIdx.x;
    const int n1 = (n + 1) / 2, n2 = n - n1;
    if (idx >= n1) return;
    const double u0 = du[offu + idx], u1 = idx < n2 ? du[offu + n1 + idx] : 0.0;
    for (int c = 0; c < nrhs; ++c) {
        double *b = dB[mat] + (size_t)c * lddb + off + idx;
        const double lo = idx < n2 ? b[n1] : 0.0;
        const double a1 = __dadd_rn(b[0], lo), a2 = __dsub_rn(b[0], lo);
        b[0] = __dmul_rn(u0, a1);
        if (idx < n2) b[n1] = __dmul_rn(u1, a2);
    }
}

constexpr long RBT_MAX_MATS = 1L << 20;  // matrices per launch (grid.x = matrices x tiles stays below 2^31)

}  // namespace

extern "C" {

// magmablas/zgerbt_func_batched.cu:183-210: A <- U^T A V, both levels
void magmablas_dprbt_batched(magma_int_t n, double **dA_array, magma_int_t ldda, double *du, double *dv,
                             magma_int_t batchCount, magma_queue_t queue)
{
    if (n <= 0 || batchCount <= 0) return;
    const int n1 = (n + 1) / 2, n2 = n - n1;
    cudaStream_t s = MB200_Q(queue)->stream;
    RbtBlocks l1, l2;
    // inner level: the four quadrants, butterfly entries [n, 2n)
    l1.b[0] = {n1, n1, 0, 0, n + 0, n + 0};
    l1.b[1] = {n1, n2, 0, n1, n + 0, n + n1};
    l1.b[2] = {n2, n1, n1, 0, n + n1, n + 0};
    l1.b[3] = {n2, n2, n1, n1, n + n1, n + n1};
    // outer level: the whole matrix, entries [0, n)
    l2.b[0] = {n, n, 0, 0, 0, 0};
    l2.b[1] = l2.b[2] = l2.b[3] = l2.b[0];
    const int q1 = (n1 + 1) / 2;  // rows / columns of the largest top-left part at each level
    for (long off = 0; off < batchCount; off += RBT_MAX_MATS) {
        const long cnt = batchCount - off < RBT_MAX_MATS ? batchCount - off : RBT_MAX_MATS;
        {
            const int tx = (q1 + RBT_TX - 1) / RBT_TX, ty = (q1 + RBT_TY - 1) / RBT_TY;
            prbt_level_kernel<<<dim3((unsigned)(cnt * tx * ty), n2 > 0 ? 4 : 1), dim3(RBT_TX, RBT_TY), 0, s>>>(
                l1, dA_array + off, ldda, du, dv, tx, ty);
            count_launch();
        }
        {
            const int tx = (n1 + RBT_TX - 1) / RBT_TX, ty = (n1 + RBT_TY - 1) / RBT_TY;
            prbt_level_kernel<<<dim3((unsigned)(cnt * tx * ty), 1), dim3(RBT_TX, RBT_TY), 0, s>>>(l2, dA_array + off, ldda, du,
                                                                                              dv, tx, ty);
            count_launch();
        }
        MB200_CHECK_LAUNCH_VOID("prbt_level_kernel");
    }
}

// magmablas/zgerbt_func_batched.cu:117-138: B <- V B (outer level first, then the two halves)
void magmablas_dprbt_mv_batched(magma_int_t n, magma_int_t nrhs, double *dv, double **db_array, magma_int_t lddb,
                                magma_int_t batchCount, magma_queue_t queue)
{
    if (n <= 0 || nrhs <= 0 || batchCount <= 0) return;
    const int n1 = (n + 1) / 2, n2 = n - n1;
    cudaStream_t s = MB200_Q(queue)->stream;
    for (long off = 0; off < batchCount; off += RBT_MAX_MATS) {
        const long cnt = batchCount - off < RBT_MAX_MATS ? batchCount - off : RBT_MAX_MATS;
        int t = ((n1 + 255) / 256);
        prbt_mv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n, nrhs, dv, 0, db_array + off, lddb, 0, t);
        t = (((n1 + 1) / 2 + 255) / 256);
        prbt_mv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n1, nrhs, dv, n, db_array + off, lddb, 0, t);
        if (n2 > 0) prbt_mv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n2, nrhs, dv, n + n1, db_array + off, lddb, n1, t);
        count_launch(n2 > 0 ? 3 : 2);
        MB200_CHECK_LAUNCH_VOID("prbt_mv_level_kernel");
    }
}

// magmablas/zgerbt_func_batched.cu:57-78: B <- U^T B (the two halves first, then the outer level)
void magmablas_dprbt_mtv_batched(magma_int_t n, magma_int_t nrhs, double *du, double **db_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue)
{
    if (n <= 0 || nrhs <= 0 || batchCount <= 0) return;
    const int n1 = (n + 1) / 2, n2 = n - n1;
    cudaStream_t s = MB200_Q(queue)->stream;
    for (long off = 0; off < batchCount; off += RBT_MAX_MATS) {
        const long cnt = batchCount - off < RBT_MAX_MATS ? batchCount - off : RBT_MAX_MATS;
        int t = (((n1 + 1) / 2 + 255) / 256);
        prbt_mtv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n1, nrhs, du, n, db_array + off, lddb, 0, t);
        if (n2 > 0) prbt_mtv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n2, nrhs, du, n + n1, db_array + off, lddb, n1, t);
        t = ((n1 + 255) / 256);
        prbt_mtv_level_kernel<<<(unsigned)(cnt * t), 256, 0, s>>>(n, nrhs, du, 0, db_array + off, lddb, 0, t);
        count_launch(n2 > 0 ? 3 : 2);
        MB200_CHECK_LAUNCH_VOID("prbt_mtv_level_kernel");
    }
}

// src/zgerbt_batched.cpp:118-181. gen = MagmaTrue: U and V (host arrays of 2n) are generated here exactly as the reference
// does (exp((rand()/RAND_MAX - 0.5)/10), u then v per entry); gen = MagmaFalse: the caller's values are used.
magma_int_t magma_dgerbt_batched(magma_bool_t gen, magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                 double **dB_array, magma_int_t lddb, double *U, double *V, magma_int_t *info,
                                 magma_int_t batchCount, magma_queue_t queue)
{
    *info = 0;
    if (!(gen == MagmaTrue) && !(gen == MagmaFalse)) *info = -1;
    else if (n < 0) *info = -2;
    else if (nrhs < 0) *info = -3;
    else if (ldda < imax(1, n)) *info = -5;
    else if (lddb < imax(1, n)) *info = -7;
    if (*info != 0) {
        magma_xerbla(__func__, -(*info));
        return *info;
    }
    if (nrhs == 0 || n == 0) return *info;
    if (gen == MagmaTrue) {
        for (int idx = 0; idx < 2 * n; ++idx) {
            U[idx] = exp((((rand() * 1.0) / RAND_MAX) - 0.5) / 10);
            V[idx] = exp((((rand() * 1.0) / RAND_MAX) - 0.5) / 10);
        }
    }
    double *duv = (double *)queue_dscratch(queue, sizeof(double) * 4 * (size_t)n, 0);
    if (!duv) {
        *info = MAGMA_ERR_DEVICE_ALLOC;
        return *info;
    }
    cudaStream_t s = MB200_Q(queue)->stream;
    cudaMemcpyAsync(duv, U, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(duv + 2 * n, V, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s);
    magmablas_dprbt_batched(n, dA_array, ldda, duv, duv + 2 * n, batchCount, queue);
    magmablas_dprbt_mtv_batched(n, nrhs, duv, dB_array, lddb, batchCount, queue);
    cudaStreamSynchronize(s);  // U, V are host arrays the caller may reuse (the reference syncs here too)
    return *info;
}

// src/zgesv_rbt_batched.cpp:81-166
magma_int_t magma_dgesv_rbt_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                    double **dB_array, magma_int_t lddb, magma_int_t *dinfo_array, magma_int_t batchCount,
                                    magma_queue_t queue)
{
    magma_int_t info = 0;
    if (n < 0) info = -1;
    else if (nrhs < 0) info = -2;
    else if (ldda < imax(1, n)) info = -4;
    else if (lddb < imax(1, n)) info = -6;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return info;
    }
    if (n == 0 || nrhs == 0) return info;
    double *huv = (double *)malloc(sizeof(double) * 4 * (size_t)n);
    if (!huv) return MAGMA_ERR_HOST_ALLOC;
    double *hu = huv, *hv = huv + 2 * n;
    magma_int_t ginfo = 0;
    info = magma_dgerbt_batched(MagmaTrue, n, nrhs, dA_array, ldda, dB_array, lddb, hu, hv, &ginfo, batchCount, queue);
    if (info == MAGMA_SUCCESS) info = magma_dgetrf_nopiv_batched(n, n, dA_array, ldda, dinfo_array, batchCount, queue);
    if (info == MAGMA_SUCCESS)
        info = magma_dgetrs_nopiv_batched(MagmaNoTrans, n, nrhs, dA_array, ldda, dB_array, lddb, dinfo_array, batchCount, queue);
    if (info == MAGMA_SUCCESS) {
        // X = V Y. magma_dgerbt_batched left V on the device behind U in the queue's staging scratch
        double *duv = (double *)queue_dscratch(queue, sizeof(double) * 4 * (size_t)n, 0);
        cudaStream_t s = MB200_Q(queue)->stream;
        cudaMemcpyAsync(duv + 2 * n, hv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s);
        magmablas_dprbt_mv_batched(n, nrhs, duv + 2 * n, dB_array, lddb, batchCount, queue);
        cudaStreamSynchronize(s);  // hv is freed below
    }
    free(huv);
    return info;
}

}  // extern "C"
