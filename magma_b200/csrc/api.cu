// Host drivers: the magma_v2 batched-LU entry points, argument checks and tier dispatch.
// Replaces src/z{getrf,getrs,gesv}_batched.cpp, src/zgetrf_vbatched.cpp and the tuning tables in
// control/get_batched_crossover.cpp / control/get_ntcol.cpp for this path.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lu_common.cuh"

using namespace mb200;

namespace {

inline int imax(int a, int b) { return a > b ? a : b; }
inline int imin(int a, int b) { return a < b ? a : b; }

Dims uniform_dims(int m, int n, int ldda)
{
    Dims d;
    d.m = m;
    d.n = n;
    d.ldda = ldda;
    d.vm = d.vn = d.vldda = nullptr;
    return d;
}

// Grid-x limits: batches above 2^31-1 CTAs are split here, never inside a kernel.
constexpr long MAX_CHUNK = 1L << 24;

// Largest dimension routed to the register-file tier (lu_mid.cu); above it the blocked tier runs.
// Measured crossover on B200 (tools/gpu_probe.py, PROBE_MID): see DESIGN.md "tiers and crossovers".
std::atomic<int> g_mid_max{128};
std::atomic<int> g_fused_max{0};  // largest dimension routed to the single-launch shared-memory tier (lu_fused.cu); 0 = off (default:
                      // measured slower than the slab drivers)
std::atomic<int> g_host_chunk_mb{64};  // host front ends: payload per staging buffer (MB200_HOST_CHUNK_MB overrides, for sweeps)

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// Tuning tables (re-derived for B200; see DESIGN.md "tiers and crossovers").
// ---------------------------------------------------------------------------------------------
void magma_get_dgetrf_batched_nbparam(magma_int_t n, magma_int_t *nb, magma_int_t *recnb)
{
    // Reference: nb = 128, recnb = 32 for every n (control/get_batched_crossover.cpp:300-305).
    // Here the outer step IS the register panel width: 32 columns while the panel is <= 512 rows
    // tall, halving each time the height doubles; there is no inner recursion (recnb == nb).
    // n stands for the panel height (square matrices: the first, tallest panel); the driver asks the same table
    // with the rows that are left at each step
    const int w = panel_width_for_rows(n);
    *nb = w;
    *recnb = w;
}

void magma_get_dgetrf_vbatched_nbparam(magma_int_t max_m, magma_int_t max_n, magma_int_t *nb, magma_int_t *recnb)
{
    (void)max_n;
    magma_get_dgetrf_batched_nbparam(max_m, nb, recnb);  // the panel width follows the panel HEIGHT
}

magma_int_t magma_get_dgetrf_batched_ntcol(magma_int_t m, magma_int_t n)
{
    // matrices per warp in the register tier (the reference's "ntcol" = matrices per CTA,
    // control/get_ntcol.cpp:197-210, is 1 for n >= 17 on anything newer than Volta)
    const int per_warp = small_tier_matrices_per_warp(m, n);
    return per_warp > 0 ? per_warp * 4 : 1;  // x 4 warps per CTA; above the register tier one matrix per CTA
}

// which: 0 register tier (max(m,n) <=), 1 register-file tier (<=), 2 left-looking slab driver (rows <=),
// 3 single-launch shared-memory tier (<=, 0 = off)
magma_int_t magma_b200_get_dgetrf_batched_crossover(magma_int_t which)
{
    switch (which) {
        case 0: return XOVER_SMALL;
        case 1: return XOVER_MID;
        case 2: return XOVER_LEFT_ROWS;
        case 3: return g_fused_max;
        default: return -1;
    }
}

magma_int_t magma_get_dtrsm_batched_stop_nb(magma_side_t side, magma_int_t m, magma_int_t n)
{
    (void)side; (void)m; (void)n;
    return 32;  // the solver blocks the triangle by 32 (one warp per diagonal block), no recursion
}

// ---------------------------------------------------------------------------------------------
magma_int_t magma_dgetrf_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ldda,
                                 magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                 magma_queue_t queue)
{
    magma_int_t arginfo = 0;
    if (m < 0) arginfo = -1;
    else if (n < 0) arginfo = -2;
    else if (ldda < imax(1, m)) arginfo = -4;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (m == 0 || n == 0 || batchCount <= 0) return 0;

    const Dims d = uniform_dims(m, n, ldda);
    cudaStream_t s = MB200_Q(queue)->stream;
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        magma_int_t rc = -100;
        if (magma_get_dgetrf_batched_ntcol(m, n) > 1 && g_tier != 2)  // register tier
            rc = lu_small_launch(d, m, n, dA_array + off, ipiv_array + off, info_array + off, 0, nullptr, 0, cnt,
                                 nullptr, s);
        else if (m <= g_fused_max && n <= g_fused_max && g_tier != 2)
            rc = lu_fused_launch(d, m, n, dA_array + off, ipiv_array + off, info_array + off, cnt, nullptr, s);
        else if (m <= g_mid_max && n <= g_mid_max && g_tier != 2)
            rc = lu_mid_launch(d, m, n, dA_array + off, ipiv_array + off, info_array + off, cnt, nullptr, s);
        if (rc == -100) {
            const size_t rec_bytes = (lu_blocked_workspace_bytes(cnt) + 255) & ~(size_t)255;
            const size_t perm_bytes = lu_blocked_perm_bytes(cnt, m, n);
            void *ws = queue_dscratch(queue, rec_bytes + perm_bytes);
            if (!ws) {
                magma_xerbla(__func__, -MAGMA_ERR_DEVICE_ALLOC);
                return MAGMA_ERR_DEVICE_ALLOC;
            }
            rc = lu_blocked_launch(d, m, n, dA_array + off, ipiv_array + off, info_array + off, cnt, nullptr, ws, s,
                                   perm_bytes ? (char *)ws + rec_bytes : nullptr);
        }
        if (rc != 0) {
            magma_xerbla(__func__, -rc);
            return rc;
        }
    }
    return 0;
}

magma_int_t magma_dgetrf_batched_smallsq_noshfl(magma_int_t n, double **dA_array, magma_int_t ldda,
                                                magma_int_t **ipiv_array, magma_int_t *info_array,
                                                magma_int_t batchCount, magma_queue_t queue)
{
    // magmablas/zgetrf_batched_smallsq_noshfl.cu:202-213: -1 n (n<0 or n>32), -3 ldda
    magma_int_t arginfo = 0;
    if (n < 0 || n > 32) arginfo = -1;
    else if (ldda < imax(1, n)) arginfo = -3;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (n == 0 || batchCount <= 0) return 0;
    const Dims d = uniform_dims(n, n, ldda);
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        magma_int_t rc = lu_small_launch(d, n, n, dA_array + off, ipiv_array + off, info_array + off, 0, nullptr, 0,
                                         cnt, nullptr, MB200_Q(queue)->stream);
        if (rc != 0) return rc;
    }
    return 0;
}

magma_int_t magma_dgesv_batched_small(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                      magma_int_t **dipiv_array, double **dB_array, magma_int_t lddb,
                                      magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t arginfo = 0;
    if (n < 0) arginfo = -1;
    else if (nrhs < 0) arginfo = -2;
    else if (ldda < imax(1, n)) arginfo = -4;
    else if (lddb < imax(1, n)) arginfo = -6;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (n == 0 || nrhs == 0 || batchCount <= 0) return 0;
    if (n > 32 || nrhs != 1 || g_tier == 2) return -100;  // outside the fused kernel: caller falls back
    const Dims d = uniform_dims(n, n, ldda);
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        magma_int_t rc = lu_small_launch(d, n, n, dA_array + off, dipiv_array + off, dinfo_array + off, nrhs,
                                         dB_array + off, lddb, cnt, nullptr, MB200_Q(queue)->stream);
        if (rc != 0) return rc;
    }
    return 0;
}

magma_int_t magma_dgetrs_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, double **dA_array,
                                 magma_int_t ldda, magma_int_t **dipiv_array, double **dB_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t info = 0;
    if (trans != MagmaNoTrans && trans != MagmaTrans && trans != MagmaConjTrans) info = -1;
    else if (n < 0) info = -2;
    else if (nrhs < 0) info = -3;
    else if (ldda < imax(1, n)) info = -5;
    else if (lddb < imax(1, n)) info = -8;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return info;
    }
    if (n == 0 || nrhs == 0 || batchCount <= 0) return 0;
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        magma_int_t rc = getrs_launch(trans, n, nrhs, dA_array + off, ldda, dipiv_array + off, dB_array + off, lddb,
                                      cnt, MB200_Q(queue)->stream);
        if (rc != 0) {
            magma_xerbla(__func__, -rc);
            return rc;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LU without pivoting.   src/zgetrf_nopiv_batched.cpp:75-170, src/zgetrs_nopiv_batched.cpp,
// src/zgesv_nopiv_batched.cpp. Same kernels as the pivoted path: register panels with the diagonal as pivot,
// left-looking slab updates; no interchange pass. At most 512 rows (the register panel's reach).
// ---------------------------------------------------------------------------------------------
magma_int_t magma_dgetrf_nopiv_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ldda,
                                       magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t arginfo = 0;
    if (m < 0) arginfo = -1;
    else if (n < 0) arginfo = -2;
    else if (ldda < imax(1, m)) arginfo = -4;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (batchCount <= 0) return 0;
    cudaStream_t s = MB200_Q(queue)->stream;
    cudaMemsetAsync(info_array, 0, sizeof(int) * (size_t)batchCount, s);  // src/zgetrf_nopiv_batched.cpp:85
    if (m == 0 || n == 0) return 0;
    if (m > 512) {
        magma_xerbla(__func__, -MAGMA_ERR_NOT_SUPPORTED);
        return MAGMA_ERR_NOT_SUPPORTED;
    }
    const Dims d = uniform_dims(m, n, ldda);
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        const size_t rec_bytes = (lu_blocked_workspace_bytes(cnt) + 255) & ~(size_t)255;
        const size_t perm_bytes = lu_blocked_perm_bytes(cnt, m, n, true);
        void *ws = queue_dscratch(queue, rec_bytes + perm_bytes);
        if (!ws) {
            magma_xerbla(__func__, -MAGMA_ERR_DEVICE_ALLOC);
            return MAGMA_ERR_DEVICE_ALLOC;
        }
        const magma_int_t rc = lu_blocked_launch(d, m, n, dA_array + off, nullptr, info_array + off, cnt, nullptr, ws, s,
                                                 (char *)ws + rec_bytes, 1);
        if (rc != 0) {
            magma_xerbla(__func__, -rc);
            return rc;
        }
    }
    return 0;
}

magma_int_t magma_dgetrs_nopiv_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, double **dA_array,
                                       magma_int_t ldda, double **dB_array, magma_int_t lddb, magma_int_t *info_array,
                                       magma_int_t batchCount, magma_queue_t queue)
{
    (void)info_array;  // not written by the reference either (src/zgetrs_nopiv_batched.cpp)
    magma_int_t info = 0;
    if (trans != MagmaNoTrans && trans != MagmaTrans && trans != MagmaConjTrans) info = -1;
    else if (n < 0) info = -2;
    else if (nrhs < 0) info = -3;
    else if (ldda < imax(1, n)) info = -5;
    else if (lddb < imax(1, n)) info = -8;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return info;
    }
    if (n == 0 || nrhs == 0 || batchCount <= 0) return 0;
    for (long off = 0; off < batchCount; off += MAX_CHUNK) {
        const long cnt = std::min<long>(MAX_CHUNK, batchCount - off);
        const magma_int_t rc = getrs_launch(trans, n, nrhs, dA_array + off, ldda, nullptr, dB_array + off, lddb, cnt,
                                            MB200_Q(queue)->stream);
        if (rc != 0) {
            magma_xerbla(__func__, -rc);
            return rc;
        }
    }
    return 0;
}

magma_int_t magma_dgesv_nopiv_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                      double **dB_array, magma_int_t lddb, magma_int_t *info_array,
                                      magma_int_t batchCount, magma_queue_t queue)
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
    if (n == 0 || nrhs == 0) return 0;
    info = magma_dgetrf_nopiv_batched(n, n, dA_array, ldda, info_array, batchCount, queue);
    if (info != MAGMA_SUCCESS) return info;
    return magma_dgetrs_nopiv_batched(MagmaNoTrans, n, nrhs, dA_array, ldda, dB_array, lddb, info_array, batchCount, queue);
}

// inv(A_b) from the factors, out of place.   src/zgetri_outofplace_batched.cpp:81-141
// The reference forms U^-1 L^-1 I and then applies the column interchanges in reverse; column j of that product is
// U^-1 L^-1 e_pi(j), i.e. exactly the solve A X = I with the interchanges applied to the identity first: one
// identity fill + the getrs path (same kernels, same canonical order as oracle_dgetrs), nothing is re-derived.
// n <= 64 runs the same arithmetic in one launch (getri.cu: permuted identity built in registers, DMMA block updates).
// Like the reference, info_array is not written and dA_array is left untouched.
magma_int_t magma_dgetri_outofplace_batched(magma_int_t n, double **dA_array, magma_int_t ldda,
                                            magma_int_t **dipiv_array, double **dinvA_array, magma_int_t lddia,
                                            magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue)
{
    (void)info_array;
    magma_int_t info = 0;
    if (n < 0) info = -1;
    else if (ldda < imax(1, n)) info = -3;
    else if (lddia < imax(1, n)) info = -6;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return info;
    }
    if (n == 0 || batchCount <= 0) return 0;
    const magma_int_t rc = getri_fused_launch(n, dA_array, ldda, dipiv_array, dinvA_array, lddia, batchCount, MB200_Q(queue)->stream);
    if (rc != -100) return rc;
    identity_launch(n, dinvA_array, lddia, batchCount, MB200_Q(queue)->stream);
    return magma_dgetrs_batched(MagmaNoTrans, n, n, dA_array, ldda, dipiv_array, dinvA_array, lddia, batchCount, queue);
}

magma_int_t magma_dgesv_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                magma_int_t **dipiv_array, double **dB_array, magma_int_t lddb,
                                magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue)
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
    if (n == 0 || nrhs == 0) return 0;
    info = magma_dgesv_batched_small(n, nrhs, dA_array, ldda, dipiv_array, dB_array, lddb, dinfo_array, batchCount,
                                     queue);
    if (info == 0) return 0;
    info = magma_dgetrf_batched(n, n, dA_array, ldda, dipiv_array, dinfo_array, batchCount, queue);
    if (info != MAGMA_SUCCESS) return info;
    return magma_dgetrs_batched(MagmaNoTrans, n, nrhs, dA_array, ldda, dipiv_array, dB_array, lddb, batchCount,
                                queue);
}

// ---------------------------------------------------------------------------------------------
// Variable-size LU. Workspace layout (ints): [ 7 index lists of `batch` | counts (8) ], then the blocked
// tier's pivot records. Matrices are binned by max(m, n) and every bin runs the tier built for it.
// ---------------------------------------------------------------------------------------------
static size_t vbatched_lists_bytes(long batch) { return (((size_t)(7 * batch + 8) * sizeof(int)) + 511) & ~(size_t)511; }
// lists, pivot records, and (when the total still fits the int-sized lwork of the reference API) the
// step-permutation records that let every class of at most 512 rows run the left-looking driver
static size_t vbatched_base_bytes(long batch)
{
    return (vbatched_lists_bytes(batch) + lu_blocked_workspace_bytes(batch) + 255) & ~(size_t)255;
}
// int_cap: the caller-provided workspace of magma_dgetrf_vbatched_max_nocheck_work is sized through an int (lwork), so
// there the step-permutation records are dropped once the total no longer fits; the entries that use queue scratch have
// no such limit and always keep the left-looking driver.
static size_t vbatched_perm_bytes(long batch, bool int_cap)
{
    const size_t p = lu_blocked_perm_bytes(batch, 512, 512);
    return (!int_cap || vbatched_base_bytes(batch) + p <= 0x7fffff00ull) ? p : 0;
}
static size_t vbatched_work_bytes(long batch, bool int_cap) { return vbatched_base_bytes(batch) + vbatched_perm_bytes(batch, int_cap); }

// known[c] >= 0: size of bin c (read back by the synchronous driver); < 0: unknown (asynchronous expert
// entries: every list is pre-filled with -1 and launched over `batch` slots, CTAs that draw -1 exit).
static magma_int_t vbatched_run(magma_int_t *m, magma_int_t *n, int max_m, int max_n, double **dA_array,
                                magma_int_t *ldda, magma_int_t **ipiv_array, magma_int_t *info_array, void *work,
                                long batch, magma_queue_t queue, const int *known, bool int_cap)
{
    cudaStream_t s = MB200_Q(queue)->stream;
    Dims d;
    d.m = max_m;
    d.n = max_n;
    d.ldda = 0;
    d.vm = m;
    d.vn = n;
    d.vldda = ldda;
    // empty matrices get info = 0 and are otherwise untouched
    cudaMemsetAsync(info_array, 0, sizeof(int) * batch, s);
    if (max_m <= 32 && max_n <= 32 && g_tier != 2)
        return lu_small_launch(d, max_m, max_n, dA_array, ipiv_array, info_array, 0, nullptr, 0, batch, nullptr, s);

    int *lists = (int *)work;
    int *counts = lists + 7 * batch;
    void *recs = (char *)work + vbatched_lists_bytes(batch);
    long cnt[7];
    for (int c = 0; c < 7; ++c) cnt[c] = known ? known[c] : batch;
    if (!known) cudaMemsetAsync(lists, 0xFF, sizeof(int) * 7 * (size_t)batch, s);
    const int mid_max = (g_tier == 2) ? 32 : g_mid_max.load();
    vbatched_partition_launch(m, n, batch, lists, counts, mid_max, s);
    magma_int_t rc = 0;
    if (cnt[0] > 0) {
        if (g_tier != 2) rc = lu_small_launch(d, 32, 32, dA_array, ipiv_array, info_array, 0, nullptr, 0, cnt[0], lists, s);
        else rc = lu_blocked_launch(d, 32, 32, dA_array, ipiv_array, info_array, cnt[0], lists, recs, s);
    }
    static const int cap[3] = {64, 96, 128};
    void *perm = vbatched_perm_bytes(batch, int_cap) ? (char *)work + vbatched_base_bytes(batch) : nullptr;
    for (int c = 1; c <= 3 && rc == 0; ++c)
        if (cnt[c] > 0) {
            rc = lu_mid_launch(d, imin(max_m, cap[c - 1]), imin(max_n, cap[c - 1]), dA_array, ipiv_array, info_array,
                               cnt[c], lists + (size_t)c * batch, s);
            if (rc == -100)  // the register-file tier hands the 65..96 window to the left-looking blocked driver
                rc = lu_blocked_launch(d, imin(max_m, cap[c - 1]), imin(max_n, cap[c - 1]), dA_array, ipiv_array, info_array,
                                       cnt[c], lists + (size_t)c * batch, recs, s, perm);
        }
    // blocked tier, one step sequence per size class (the pivot records are reused: the stream serialises them)
    static const int bcap[3] = {256, 384, 0x7fffffff};
    for (int c = 4; c <= 6 && rc == 0; ++c)
        if (cnt[c] > 0)
            rc = lu_blocked_launch(d, imin(max_m, bcap[c - 4]), imin(max_n, bcap[c - 4]), dA_array, ipiv_array, info_array,
                                   cnt[c], lists + (size_t)c * batch, recs, s, perm);
    return rc;
}

magma_int_t magma_dgetrf_vbatched_max_nocheck_work(magma_int_t *m, magma_int_t *n, magma_int_t max_m,
                                                   magma_int_t max_n, magma_int_t max_minmn, magma_int_t max_mxn,
                                                   double **dA_array, magma_int_t *ldda, magma_int_t **dipiv_array,
                                                   magma_int_t *info_array, void *work, magma_int_t *lwork,
                                                   magma_int_t batchCount, magma_queue_t queue)
{
    (void)max_minmn; (void)max_mxn;
    const size_t need = vbatched_work_bytes(batchCount, true);
    if (need > 0x7fffffffull) {  // even the index lists and pivot records do not fit an int-sized lwork: say so
        magma_xerbla(__func__, 13);
        return -13;
    }
    if (lwork[0] < 0) {  // workspace query (src/zgetrf_vbatched.cpp:257-261)
        lwork[0] = (magma_int_t)need;
        return 0;
    }
    if ((size_t)lwork[0] < need) {
        magma_xerbla(__func__, 12);
        return -12;
    }
    if (batchCount <= 0 || max_m == 0 || max_n == 0) return 0;
    return vbatched_run(m, n, max_m, max_n, dA_array, ldda, dipiv_array, info_array, work, batchCount, queue, nullptr, true);
}

magma_int_t magma_dgetrf_vbatched_max_nocheck(magma_int_t *m, magma_int_t *n, magma_int_t *minmn, magma_int_t max_m,
                                              magma_int_t max_n, magma_int_t max_minmn, magma_int_t max_mxn,
                                              magma_int_t nb, magma_int_t recnb, double **dA_array, magma_int_t *ldda,
                                              magma_int_t **ipiv_array, magma_int_t **pivinfo_array,
                                              magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue)
{
    (void)minmn; (void)max_minmn; (void)max_mxn; (void)nb; (void)recnb; (void)pivinfo_array;
    if (batchCount <= 0 || max_m == 0 || max_n == 0) return 0;
    void *work = queue_dscratch(queue, vbatched_work_bytes(batchCount, false));
    if (!work) {
        magma_xerbla(__func__, -MAGMA_ERR_DEVICE_ALLOC);
        return MAGMA_ERR_DEVICE_ALLOC;
    }
    return vbatched_run(m, n, max_m, max_n, dA_array, ldda, ipiv_array, info_array, work, batchCount, queue, nullptr, false);
}

magma_int_t magma_dgetrf_vbatched(magma_int_t *m, magma_int_t *n, double **dA_array, magma_int_t *ldda,
                                  magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                  magma_queue_t queue)
{
    if (batchCount < 0) {
        magma_xerbla(__func__, 7);
        return -7;
    }
    if (batchCount == 0) return 0;
    // one statistics kernel + one D2H read (the reference: checker kernel + read, setup kernel + read)
    char *scr = (char *)queue_dscratch(queue, vbatched_work_bytes(batchCount, false) + 64);  // 64: the statistics block
    if (!scr) {
        magma_xerbla(__func__, -MAGMA_ERR_DEVICE_ALLOC);
        return MAGMA_ERR_DEVICE_ALLOC;
    }
    int *stats = (int *)scr;  // 16 ints, then the partition workspace
    void *work = scr + 64;
    vbatched_stats_launch(m, n, ldda, batchCount, stats, MB200_Q(queue)->stream);
    int h[16];
    cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, MB200_Q(queue)->stream);
    cudaStreamSynchronize(MB200_Q(queue)->stream);
    if (h[4] != 0) {
        const int arg = 8 - h[4];  // 1: m, 2: n, 4: ldda (src/zgetrf_vbatched.cpp:356-364 via the checker)
        magma_xerbla(__func__, arg);
        return -arg;
    }
    const int max_m = h[0], max_n = h[1];
    if (max_m == 0 || max_n == 0 || h[6] == 0) {
        cudaMemsetAsync(info_array, 0, sizeof(int) * (size_t)batchCount, MB200_Q(queue)->stream);
        return 0;
    }
    // bin sizes as the partition kernel will produce them (classes above mid_max fall into the last bin)
    const int mid_max = (g_tier == 2) ? 32 : g_mid_max.load();
    int known[7] = {h[5], h[8], h[9], h[10], h[11], h[12], 0};
    // classes above mid_max belong to the blocked bins (all three are <= 256, so they join bin 4)
    if (mid_max < 64) { known[4] += known[1]; known[1] = 0; }
    if (mid_max < 96) { known[4] += known[2]; known[2] = 0; }
    if (mid_max < 128) { known[4] += known[3]; known[3] = 0; }
    known[6] = h[6] - known[0] - known[1] - known[2] - known[3] - known[4] - known[5];
    magma_int_t rc = vbatched_run(m, n, max_m, max_n, dA_array, ldda, ipiv_array, info_array, work, batchCount, queue,
                                  known, false);
    // the reference's driver returns after a queue sync (src/zgetrf_vbatched.cpp:392)
    cudaStreamSynchronize(MB200_Q(queue)->stream);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone pieces kept for source compatibility.
// ---------------------------------------------------------------------------------------------
void magma_dlaswp_rowserial_batched(magma_int_t n, double **dA_array, magma_int_t lda, magma_int_t k1,
                                    magma_int_t k2, magma_int_t **ipiv_array, magma_int_t batchCount,
                                    magma_queue_t queue)
{
    laswp_rowserial_launch(n, dA_array, lda, k1, k2, ipiv_array, batchCount, MB200_Q(queue)->stream);
}

void magmablas_dtrsm_batched(magma_side_t side, magma_uplo_t uplo, magma_trans_t transA, magma_diag_t diag,
                             magma_int_t m, magma_int_t n, double alpha, double **dA_array, magma_int_t ldda,
                             double **dB_array, magma_int_t lddb, magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t info = 0;
    if (side != MagmaLeft && side != MagmaRight) info = -1;
    else if (uplo != MagmaUpper && uplo != MagmaLower) info = -2;
    else if (transA != MagmaNoTrans && transA != MagmaTrans && transA != MagmaConjTrans) info = -3;
    else if (diag != MagmaUnit && diag != MagmaNonUnit) info = -4;
    else if (m < 0) info = -5;
    else if (n < 0) info = -6;
    else if (ldda < imax(1, side == MagmaLeft ? m : n)) info = -9;
    else if (lddb < imax(1, m)) info = -11;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return;
    }
    if (side == MagmaRight) {
        trsm_right_launch(uplo, transA, diag, m, n, alpha, dA_array, ldda, dB_array, lddb, batchCount, MB200_Q(queue)->stream);
        return;
    }
    trsm_left_launch(uplo, transA, diag, m, n, alpha, dA_array, ldda, dB_array, lddb, batchCount, MB200_Q(queue)->stream);
}

void magma_dgemm_batched_core(magma_trans_t transA, magma_trans_t transB, magma_int_t m, magma_int_t n,
                              magma_int_t k, double alpha, double const *const *dA_array, magma_int_t Ai,
                              magma_int_t Aj, magma_int_t ldda, double const *const *dB_array, magma_int_t Bi,
                              magma_int_t Bj, magma_int_t lddb, double beta, double **dC_array, magma_int_t Ci,
                              magma_int_t Cj, magma_int_t lddc, magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t info = 0;
    if (transA != MagmaNoTrans && transA != MagmaTrans && transA != MagmaConjTrans) info = -1;
    else if (transB != MagmaNoTrans && transB != MagmaTrans && transB != MagmaConjTrans) info = -2;
    else if (m < 0) info = -3;
    else if (n < 0) info = -4;
    else if (k < 0) info = -5;
    else if (ldda < imax(1, transA == MagmaNoTrans ? m : k)) info = -8;
    else if (lddb < imax(1, transB == MagmaNoTrans ? k : n)) info = -10;
    else if (lddc < imax(1, m)) info = -13;
    if (info != 0) {
        magma_xerbla(__func__, -info);
        return;
    }
    if (g_tier == 4)  // DFMA kernels only (A/B runs): the SIMT tile, NoTrans x NoTrans
        if (transA == MagmaNoTrans && transB == MagmaNoTrans) {
            gemm_nn_launch(m, n, k, alpha, dA_array, Ai, Aj, ldda, dB_array, Bi, Bj, lddb, beta, dC_array, Ci, Cj, lddc,
                           batchCount, MB200_Q(queue)->stream);
            return;
        }
    gemm_dmma_launch(transA, transB, m, n, k, alpha, dA_array, Ai, Aj, ldda, dB_array, Bi, Bj, lddb, beta, dC_array, Ci, Cj,
                     lddc, batchCount, MB200_Q(queue)->stream);
}

void magma_dset_pointer(double **output_array, double *input, magma_int_t lda, magma_int_t row, magma_int_t column,
                        magma_int_t batch_offset, magma_int_t batchCount, magma_queue_t queue)
{
    set_pointer_launch((void **)output_array, (char *)input, sizeof(double), lda, row, column, batch_offset,
                       batchCount, MB200_Q(queue)->stream);
}

void magma_iset_pointer(magma_int_t **output_array, magma_int_t *input, magma_int_t lda, magma_int_t row,
                        magma_int_t column, magma_int_t batchSize, magma_int_t batchCount, magma_queue_t queue)
{
    set_pointer_launch((void **)output_array, (char *)input, sizeof(magma_int_t), lda, row, column, batchSize,
                       batchCount, MB200_Q(queue)->stream);
}

void magma_ddisplace_pointers(double **output_array, double **input_array, magma_int_t lda, magma_int_t row,
                              magma_int_t column, magma_int_t batchCount, magma_queue_t queue)
{
    displace_pointers_launch((void **)output_array, (void **)input_array, sizeof(double), lda, row, column,
                             batchCount, MB200_Q(queue)->stream);
}

void magma_idisplace_pointers(magma_int_t **output_array, magma_int_t **input_array, magma_int_t lda,
                              magma_int_t row, magma_int_t column, magma_int_t batchCount, magma_queue_t queue)
{
    displace_pointers_launch((void **)output_array, (void **)input_array, sizeof(magma_int_t), lda, row, column,
                             batchCount, MB200_Q(queue)->stream);
}

// ---------------------------------------------------------------------------------------------
// Additions
// ---------------------------------------------------------------------------------------------
int64_t magma_b200_rcp_selftest(int64_t n, magma_queue_t queue) { return rcp_selftest_run((long)n, MB200_Q(queue)->stream); }
void magma_b200_set_fused_max(int n) { g_fused_max = n > 128 ? 128 : (n < 0 ? 0 : n); }
void magma_b200_set_mid_max(int n) { g_mid_max = n >= 128 ? 128 : (n >= 96 ? 96 : (n >= 64 ? 64 : 32)); }

void magma_b200_dlarnv_uniform(magma_int_t *iseed, int64_t n, double *dx, magma_queue_t queue)
{
    const unsigned long long A = 33952834046453ull, MASK = (1ull << 48) - 1ull;
    unsigned long long s = ((unsigned long long)(iseed[0] & 4095) << 36) | ((unsigned long long)(iseed[1] & 4095) << 24) |
                           ((unsigned long long)(iseed[2] & 4095) << 12) | (unsigned long long)(iseed[3] & 4095);
    dlarnv_launch(s, n, dx, MB200_Q(queue)->stream);
    // advance the host seed by n draws: s * A^n
    unsigned long long p = A, e = (unsigned long long)n;
    while (e) {
        if (e & 1ull) s = (s * p) & MASK;
        p = (p * p) & MASK;
        e >>= 1;
    }
    iseed[0] = (int)((s >> 36) & 4095);
    iseed[1] = (int)((s >> 24) & 4095);
    iseed[2] = (int)((s >> 12) & 4095);
    iseed[3] = (int)(s & 4095);
}

double magma_b200_fp64_peak_tflops(int kind, magma_queue_t queue) { return fp64_peak_run(kind, MB200_Q(queue)->stream); }
double magma_b200_hbm_copy_gbs(size_t bytes, magma_queue_t queue) { return hbm_copy_run(bytes, MB200_Q(queue)->stream); }

magma_int_t magma_b200_dgetrf_batched_mgpu(magma_int_t ngpu, magma_int_t m, magma_int_t n, double ***dA_array,
                                           magma_int_t ldda, magma_int_t ***ipiv_array, magma_int_t **info_array,
                                           const magma_int_t *batchCount, magma_queue_t *queues)
{
    int prev = 0;
    cudaGetDevice(&prev);
    magma_int_t rc = 0;
    for (int g = 0; g < ngpu && rc == 0; ++g) {
        cudaSetDevice(MB200_Q(queues[g])->device);
        rc = magma_dgetrf_batched(m, n, dA_array[g], ldda, ipiv_array[g], info_array[g], batchCount[g], queues[g]);
    }
    cudaSetDevice(prev);
    return rc;
}

magma_int_t magma_b200_dgesv_batched_mgpu(magma_int_t ngpu, magma_int_t n, magma_int_t nrhs, double ***dA_array,
                                          magma_int_t ldda, magma_int_t ***dipiv_array, double ***dB_array,
                                          magma_int_t lddb, magma_int_t **dinfo_array, const magma_int_t *batchCount,
                                          magma_queue_t *queues)
{
    int prev = 0;
    cudaGetDevice(&prev);
    magma_int_t rc = 0;
    for (int g = 0; g < ngpu && rc == 0; ++g) {
        cudaSetDevice(MB200_Q(queues[g])->device);
        rc = magma_dgesv_batched(n, nrhs, dA_array[g], ldda, dipiv_array[g], dB_array[g], lddb, dinfo_array[g],
                                 batchCount[g], queues[g]);
    }
    cudaSetDevice(prev);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer front ends: chunked, double-buffered H2D -> compute -> D2H.
// ---------------------------------------------------------------------------------------------
static void ensure_aux(magma_queue_t queue)
{
    auto *q = MB200_Q(queue);
    if (q->aux_ready) return;
    // the auxiliary streams and events belong to the queue's device, whatever device is current
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != q->device) cudaSetDevice(q->device);
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&q->aux_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 12; ++i) cudaEventCreateWithFlags(&q->aux_event[i], cudaEventDisableTiming);
    q->aux_ready = true;
    if (prev != q->device) cudaSetDevice(prev);
}

static magma_int_t host_pipeline(bool solve, int m, int n, int nrhs, double *hA, int lda, int *hipiv, double *hB,
                                 int ldb, int *hinfo, long batch, magma_queue_t queue)
{
    ensure_aux(queue);
    if (const char *e = getenv("MB200_HOST_CHUNK_MB")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 1024) g_host_chunk_mb = v;
    }
    const int mn = imin(m, n);
    const size_t a_elems = (size_t)lda * n, b_elems = solve ? (size_t)ldb * nrhs : 0;
    const size_t per_mat = (a_elems + b_elems) * 8 + (size_t)mn * 4 + 4 + 3 * 8;
    // chunk: ~64 MiB of payload per buffer (pipeline fill + drain = one chunk each way; measured 50.0 ms per
    // headline step at 64 MiB vs 51.9 at 256 MiB and 59 at 8 MiB: PCIe-bound), at least 1 matrix, two buffers
    long chunk = (long)std::max<size_t>(1, ((size_t)g_host_chunk_mb << 20) / per_mat);
    if (chunk > batch) chunk = batch;
    const size_t stride = ((per_mat * chunk + 64 + 255) / 256) * 256;
    // staging buffers in the ring (MB200_HOST_NBUF = 2..4 for sweeps). Measured on the headline step, same box: 2 / 3 / 4
    // buffers 50.9 / 50.8 / 50.7 ms against 44.9 ms for the bare copies: the ring depth is not what separates them
    static int nbuf = 0;
    if (nbuf == 0) {
        const char *e = getenv("MB200_HOST_NBUF");
        nbuf = e ? atoi(e) : 2;
        if (nbuf < 2) nbuf = 2;
        if (nbuf > 4) nbuf = 4;
    }
    char *dev = (char *)queue_dscratch(queue, (size_t)nbuf * stride, 0);
    if (!dev) {
        magma_xerbla(solve ? "magma_b200_dgesv_batched_host" : "magma_b200_dgetrf_batched_host", -MAGMA_ERR_DEVICE_ALLOC);
        return MAGMA_ERR_DEVICE_ALLOC;
    }
    cudaStream_t sc = MB200_Q(queue)->stream;
    cudaStream_t sh = MB200_Q(queue)->aux_stream[0];  // H2D
    cudaStream_t sd = MB200_Q(queue)->aux_stream[1];  // D2H
    cudaEvent_t *ev_in = MB200_Q(queue)->aux_event;        // [4] H2D done
    cudaEvent_t *ev_done = MB200_Q(queue)->aux_event + 4;  // [4] compute done
    cudaEvent_t *ev_out = MB200_Q(queue)->aux_event + 8;   // [4] D2H done (buffer free)
    magma_int_t rc = 0;
    long it = 0;
    for (long off = 0; off < batch; off += chunk, ++it) {
        const long cnt = std::min<long>(chunk, batch - off);
        const int bsel = (int)(it % nbuf);
        char *base = dev + (size_t)bsel * stride;
        double *dAm = (double *)base;
        double *dBm = dAm + a_elems * chunk;
        int *dip = (int *)(dBm + b_elems * chunk);
        int *dinf = dip + (size_t)mn * chunk;
        // pointer arrays (8-byte aligned region after the ints)
        char *pbase = (char *)(dinf + chunk);
        pbase = (char *)(((uintptr_t)pbase + 7) & ~(uintptr_t)7);
        double **pA = (double **)pbase;
        double **pB = pA + chunk;
        int **pP = (int **)(pB + chunk);
        if (it >= nbuf) cudaStreamWaitEvent(sh, ev_out[bsel], 0);  // buffer reuse
        cudaMemcpyAsync(dAm, hA + (size_t)off * a_elems, a_elems * 8 * cnt, cudaMemcpyHostToDevice, sh);
        if (solve) cudaMemcpyAsync(dBm, hB + (size_t)off * b_elems, b_elems * 8 * cnt, cudaMemcpyHostToDevice, sh);
        cudaEventRecord(ev_in[bsel], sh);
        cudaStreamWaitEvent(sc, ev_in[bsel], 0);
        if (it >= nbuf) cudaStreamWaitEvent(sc, ev_out[bsel], 0);
        set_pointer_launch((void **)pA, (char *)dAm, 8, lda, 0, 0, (long)a_elems, cnt, sc);
        set_pointer_launch((void **)pP, (char *)dip, 4, 1, 0, 0, mn, cnt, sc);
        if (solve) {
            set_pointer_launch((void **)pB, (char *)dBm, 8, ldb, 0, 0, (long)b_elems, cnt, sc);
            rc = magma_dgesv_batched(n, nrhs, pA, lda, pP, pB, ldb, dinf, (magma_int_t)cnt, queue);
        } else {
            rc = magma_dgetrf_batched(m, n, pA, lda, pP, dinf, (magma_int_t)cnt, queue);
        }
        if (rc != 0) break;
        cudaEventRecord(ev_done[bsel], sc);
        cudaStreamWaitEvent(sd, ev_done[bsel], 0);
        cudaMemcpyAsync(hA + (size_t)off * a_elems, dAm, a_elems * 8 * cnt, cudaMemcpyDeviceToHost, sd);
        if (solve) cudaMemcpyAsync(hB + (size_t)off * b_elems, dBm, b_elems * 8 * cnt, cudaMemcpyDeviceToHost, sd);
        cudaMemcpyAsync(hipiv + (size_t)off * mn, dip, (size_t)mn * 4 * cnt, cudaMemcpyDeviceToHost, sd);
        cudaMemcpyAsync(hinfo + off, dinf, 4 * cnt, cudaMemcpyDeviceToHost, sd);
        cudaEventRecord(ev_out[bsel], sd);
    }
    cudaStreamSynchronize(sh);
    cudaStreamSynchronize(sc);
    cudaStreamSynchronize(sd);
    if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = MAGMA_ERR_UNKNOWN;
    return rc;
}

magma_int_t magma_b200_dgetrf_batched_host(magma_int_t m, magma_int_t n, double *hA, magma_int_t lda,
                                           magma_int_t *hipiv, magma_int_t *hinfo, magma_int_t batchCount,
                                           magma_queue_t queue)
{
    magma_int_t arginfo = 0;
    if (m < 0) arginfo = -1;
    else if (n < 0) arginfo = -2;
    else if (lda < imax(1, m)) arginfo = -4;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (m == 0 || n == 0 || batchCount <= 0) return 0;
    return host_pipeline(false, m, n, 0, hA, lda, hipiv, nullptr, 0, hinfo, batchCount, queue);
}

magma_int_t magma_b200_dgesv_batched_host(magma_int_t n, magma_int_t nrhs, double *hA, magma_int_t lda,
                                          magma_int_t *hipiv, double *hB, magma_int_t ldb, magma_int_t *hinfo,
                                          magma_int_t batchCount, magma_queue_t queue)
{
    magma_int_t arginfo = 0;
    if (n < 0) arginfo = -1;
    else if (nrhs < 0) arginfo = -2;
    else if (lda < imax(1, n)) arginfo = -4;
    else if (ldb < imax(1, n)) arginfo = -6;
    if (arginfo != 0) {
        magma_xerbla(__func__, -arginfo);
        return arginfo;
    }
    if (n == 0 || nrhs == 0 || batchCount <= 0) return 0;
    return host_pipeline(true, n, n, nrhs, hA, lda, hipiv, hB, ldb, hinfo, batchCount, queue);
}

// ---------------------------------------------------------------------------------------------
// F77-style wrappers (control/magma_df77.cpp conventions: scalars by reference, device pointers
// and the queue as integer handles, trailing info).
// ---------------------------------------------------------------------------------------------
void magmaf_dgetrf_batched_(magma_int_t *m, magma_int_t *n, devptr_t *dA_array, magma_int_t *ldda,
                            devptr_t *ipiv_array, devptr_t *info_array, magma_int_t *batchCount, devptr_t *queue,
                            magma_int_t *info)
{
    *info = magma_dgetrf_batched(*m, *n, (double **)(*dA_array), *ldda, (magma_int_t **)(*ipiv_array),
                                 (magma_int_t *)(*info_array), *batchCount, (magma_queue_t)(*queue));
}

void magmaf_dgetrs_batched_(const char *trans, magma_int_t *n, magma_int_t *nrhs, devptr_t *dA_array,
                            magma_int_t *ldda, devptr_t *dipiv_array, devptr_t *dB_array, magma_int_t *lddb,
                            magma_int_t *batchCount, devptr_t *queue, magma_int_t *info)
{
    magma_trans_t t = MagmaNoTrans;
    switch (*trans) {
        case 'N': case 'n': t = MagmaNoTrans; break;
        case 'T': case 't': t = MagmaTrans; break;
        case 'C': case 'c': t = MagmaConjTrans; break;
        default: t = (magma_trans_t)0; break;
    }
    *info = magma_dgetrs_batched(t, *n, *nrhs, (double **)(*dA_array), *ldda, (magma_int_t **)(*dipiv_array),
                                 (double **)(*dB_array), *lddb, *batchCount, (magma_queue_t)(*queue));
}

void magmaf_dgesv_batched_(magma_int_t *n, magma_int_t *nrhs, devptr_t *dA_array, magma_int_t *ldda,
                           devptr_t *dipiv_array, devptr_t *dB_array, magma_int_t *lddb, devptr_t *dinfo_array,
                           magma_int_t *batchCount, devptr_t *queue, magma_int_t *info)
{
    *info = magma_dgesv_batched(*n, *nrhs, (double **)(*dA_array), *ldda, (magma_int_t **)(*dipiv_array),
                                (double **)(*dB_array), *lddb, (magma_int_t *)(*dinfo_array), *batchCount,
                                (magma_queue_t)(*queue));
}

void magmaf_dgetrf_vbatched_(devptr_t *m, devptr_t *n, devptr_t *dA_array, devptr_t *ldda, devptr_t *ipiv_array,
                             devptr_t *info_array, magma_int_t *batchCount, devptr_t *queue, magma_int_t *info)
{
    *info = magma_dgetrf_vbatched((magma_int_t *)(*m), (magma_int_t *)(*n), (double **)(*dA_array),
                                  (magma_int_t *)(*ldda), (magma_int_t **)(*ipiv_array),
                                  (magma_int_t *)(*info_array), *batchCount, (magma_queue_t)(*queue));
}

}  // extern "C"
