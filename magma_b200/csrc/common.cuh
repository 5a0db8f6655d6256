// Internal definitions shared by the host drivers and kernels of libmagma_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "magma_b200.h"

// The opaque queue (replaces control/magma_internal.h:90-211). No cuBLAS/cuSPARSE handles and no
// 3x65534 pointer scratch: nothing on this path calls a vendor BLAS or displaces pointer arrays.
// Two deployment modes:
//   standalone (default): this library owns the queue type below and exports the minimal runtime around it;
//   interpose (-DMB200_INTERPOSE): the queue belongs to a real libmagma and stays opaque -- stream and device are read
//     through ITS magma_queue_get_cuda_stream / magma_queue_get_device (interface_cuda/interface.cpp:798,815), never by
//     struct layout, and this library's per-queue state (scratch, auxiliary streams) lives in a side table keyed by
//     the handle. runtime.cu's entry points are compiled out; only the batched-LU symbols are defined.
#ifdef MB200_INTERPOSE
struct magma_queue;  // opaque: owned by the real libmagma
namespace mb200 {
struct QState {
#else
struct magma_queue {
#endif
    magma_device_t device;
    cudaStream_t stream;
    bool own_stream;
    // lazily grown per-queue scratch (vbatched statistics, host front-end staging)
    void *dscratch[2];        // [0] host front-end staging, [1] kernel workspace (pivot records, bins)
    size_t dscratch_bytes[2];
    void *hscratch;  // pinned
    size_t hscratch_bytes;
    // host front ends: two extra streams + events for the H2D / compute / D2H pipeline
    cudaStream_t aux_stream[2];
    cudaEvent_t aux_event[12];
    bool aux_ready;
};
#ifdef MB200_INTERPOSE
QState *qstate(magma_queue_t q);  // finds or creates the entry; refreshes stream and device from the real queue
}  // namespace mb200
#define MB200_Q(q) (mb200::qstate(q))
#else
#define MB200_Q(q) (q)
#endif

namespace mb200 {

extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_tier;  // 0 auto, 1 force small/register tier, 2 force blocked tier
extern std::atomic<int> g_small_rows;  // register tier: rows per lane, 0 = tuned default

// ---- B200 tuning table: the ONE place tier boundaries and panel widths are written down. The drivers below and the
// reference-named getters (magma_get_dgetrf_batched_nbparam / _ntcol, magma_b200_get_dgetrf_batched_crossover) all
// read it (control/get_batched_crossover.cpp:300-305, control/get_ntcol.cpp:197-210 are its reference counterparts).
constexpr int XOVER_SMALL = 32;       // max(m,n) <= 32: register tier, one launch (lu_small*.cu)
constexpr int XOVER_MID = 44;         // 33..44: register-file tier (lu_mid.cu); above: left-looking slab driver
constexpr int XOVER_LEFT_ROWS = 512;  // at most 512 rows: left-looking slab driver; more: right-looking driver
// width of the register panel for a panel `rows` tall: 32 columns while one thread per row fits a CTA (512 rows),
// halving each time the height doubles; above 8192 rows the global-memory panel (8 columns)
inline int panel_width_for_rows(int rows)
{
    if (rows <= 512) return 32;
    if (rows <= 1024) return 16;
    if (rows <= 2048) return 8;
    if (rows <= 4096) return 4;
    if (rows <= 8192) return 2;
    return 8;
}
// matrices per warp in the register tier
inline int small_tier_matrices_per_warp(int m, int n)
{
    const int k = m > n ? m : n;
    return k <= 8 ? 4 : (k <= 16 ? 2 : (k <= XOVER_SMALL ? 1 : 0));
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Opt-in to more than 48 KB of dynamic shared memory. The attribute is PER DEVICE: a process that drives queues on
// several GPUs (magma_b200_d{getrf,gesv}_batched_mgpu) must set it on each of them, so the "already done" flag is
// kept per device, not per process.
struct DevOnce {
    bool set[64] = {};
};
template <typename F>
inline void smem_optin(DevOnce &once, F kernel, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !once.set[dev]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (dev >= 0 && dev < 64) once.set[dev] = true;
    }
}

// Scratch accessors (grow-only, never shrink; freed with the queue).
void *queue_dscratch(magma_queue_t q, size_t bytes, int slot = 1);
void *queue_hscratch(magma_queue_t q, size_t bytes);

// Launch error check: the product path fails loudly (no silent fallback).
#define MB200_CHECK_LAUNCH(name)                                                          \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            fprintf(stderr, "libmagma_b200: launch of %s failed: %s\n", name,             \
                    cudaGetErrorString(e__));                                             \
            return MAGMA_ERR_UNKNOWN;                                                     \
        }                                                                                 \
    } while (0)

#define MB200_CHECK_LAUNCH_VOID(name)                                                     \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            fprintf(stderr, "libmagma_b200: launch of %s failed: %s\n", name,             \
                    cudaGetErrorString(e__));                                             \
        }                                                                                 \
    } while (0)

// Per-matrix dimensions: uniform (fixed-size batched) or read from device arrays (vbatched).
struct Dims {
    int m, n, ldda;            // uniform values, or the maxima in variable mode
    const int *vm, *vn, *vldda;  // non-null => variable mode
};

__device__ __forceinline__ void dims_of(const Dims &d, long b, int &m, int &n, int &ld)
{
    if (d.vm) {
        m = d.vm[b];
        n = d.vn[b];
        ld = d.vldda[b];
    } else {
        m = d.m;
        n = d.n;
        ld = d.ldda;
    }
}

// ---- kernel launchers implemented in the .cu files -------------------------------------------

// lu_small.cu: whole matrix (m,n <= 32) per warp / half warp / quarter warp, optional fused solve
// with nrhs right-hand sides (nrhs == 0: factor only). Returns 0, or -100 when the shape is not
// covered (caller falls through to the blocked tier).
magma_int_t lu_small_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv,
                            int *dinfo, int nrhs, double **dB, int lddb, long batch,
                            const int *index_list, cudaStream_t s);

// lu_small_sq.cu: square fast path of the register tier (n in {8, 16, 32}, nrhs <= 1); -100 = not covered
magma_int_t lu_sq_launch(int n, double **dA, int ldda, int **dipiv, int *dinfo, int nrhs, double **dB,
                         int lddb, long batch, cudaStream_t s);

// lu_mid.cu: whole matrix in the register file of one CTA, 32 < max(m,n) <= 128; -100 = not covered
magma_int_t lu_mid_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                          long batch, const int *index_list, cudaStream_t s);

// lu_fused.cu: whole LU in one launch, matrix resident in shared memory (TMA bulk in/out), m, n <= 128; -100 = not covered
magma_int_t lu_fused_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                            long batch, const int *index_list, cudaStream_t s);

long rcp_selftest_run(long n, cudaStream_t s);
// lu_fused.cu: 32-column panel by single-warp pivot chains (panels of at most 128 rows); -100 = not covered
magma_int_t panel_chain_launch(const Dims &d, double **dA, int **dipiv, int *dinfo, int j, int T, long batch, const int *il,
                               cudaStream_t s, unsigned short *sinv, int sinv_rows, int sinv_blocks);
extern std::atomic<int> g_fused_tail;   // left-looking driver, at most 128 rows: panels factored in the slab kernel's tail (0 off = default: measured slower; 1: 33..96 rows, 2: <= 96 rows)
extern std::atomic<int> g_split;        // left-looking driver: the batch runs as this many independent parts on their own streams (0 = auto: 2 above 256 rows; 1 = off)
extern std::atomic<int> g_tall_panel2;  // left-looking driver: panels of 129..512 rows in two 16-column halves (panel2_kernel); 0: panel_kernel
extern std::atomic<int> g_chain_panel;  // level L > 0: panels of (128 - 32 L, 128] rows go to panel_chain_launch (default 3); 0: panel_kernel

// lu_blocked.cu: blocked right-looking LU for any m x n (two kernels per panel step).
// `workspace` must hold lu_blocked_workspace_bytes(batch) bytes (one 512-byte pivot record per matrix).
// `perm_workspace` (lu_blocked_perm_bytes, may be null): enables the left-looking driver for max_m <= 512.
size_t lu_blocked_workspace_bytes(long batch);
size_t lu_blocked_perm_bytes(long batch, int max_m, int max_n, bool any_width = false);
magma_int_t lu_blocked_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv,
                              int *dinfo, long batch, const int *index_list, void *workspace,
                              cudaStream_t s, void *perm_workspace = nullptr, int nopiv = 0);

// getrs.cu
magma_int_t getrs_launch(int trans, int n, int nrhs, double **dA, int ldda, int **dipiv,
                         double **dB, int lddb, long batch, cudaStream_t s);
void laswp_rowserial_launch(int n, double **dA, int lda, int k1, int k2, int **dipiv, long batch,
                            cudaStream_t s);
void trsm_left_launch(int uplo, int trans, int diag, int m, int n, double alpha, double **dA,
                      int ldda, double **dB, int lddb, long batch, cudaStream_t s);
void gemm_nn_launch(int m, int n, int k, double alpha, double const *const *dA, int Ai, int Aj,
                    int ldda, double const *const *dB, int Bi, int Bj, int lddb, double beta,
                    double **dC, int Ci, int Cj, int lddc, long batch, cudaStream_t s);

// blas3.cu: C <- alpha op(A) op(B) + beta C on the FP64 tensor pipe; X op(A) = alpha B
void gemm_dmma_launch(int transA, int transB, int m, int n, int k, double alpha, double const *const *dA, int Ai, int Aj,
                      int ldda, double const *const *dB, int Bi, int Bj, int lddb, double beta, double **dC, int Ci, int Cj,
                      int lddc, long batch, cudaStream_t s);
void trsm_right_launch(int uplo, int trans, int diag, int m, int n, double alpha, double **dA, int ldda, double **dB,
                       int lddb, long batch, cudaStream_t s);

// aux.cu
void set_pointer_launch(void **out, char *base, long elem, long lda, long row, long col,
                        long batch_offset, long batch, cudaStream_t s);
void displace_pointers_launch(void **out, void **in, long elem, long lda, long row, long col,
                              long batch, cudaStream_t s);
void memset_int_launch(int *p, int v, long n, cudaStream_t s);
void identity_launch(int n, double **dB, int lddb, long batch, cudaStream_t s);
// getri.cu: single-launch out-of-place inverse from the factors, n <= 64; -100 = not covered
magma_int_t getri_fused_launch(int n, double **dA, int ldda, int **dipiv, double **dinvA, int lddia, long batch,
                               cudaStream_t s);
// vbatched statistics: out[0..15] = {max_m, max_n, max_minmn, max_mxn(clamped), first_bad_arg,
// count(max(m,n) <= 32), count_nonempty, 0, count(<= 64), count(<= 96), count(<= 128), ...}
void vbatched_stats_launch(const int *m, const int *n, const int *ldda, long batch, int *out16,
                           cudaStream_t s);
// builds seven index lists by size class (<= 32, <= 64, <= 96, <= mid_max, <= 256, <= 384, rest), list c at lists + c*batch
void vbatched_partition_launch(const int *m, const int *n, long batch, int *lists, int *counts,
                               int mid_max, cudaStream_t s);
void dlarnv_launch(uint64_t seed48, int64_t n, double *dx, cudaStream_t s);
double fp64_peak_run(int kind, cudaStream_t s);
double hbm_copy_run(size_t bytes, cudaStream_t s);

}  // namespace mb200
