// Tier F: whole LU of one matrix (m, n <= 128) in ONE launch, one CTA per matrix, two CTAs per SM.
//
// Replaces, for 32 < max(m,n) <= 128, the reference's host recursion (src/zgetrf_batched.cpp:149-203:
// 42 launches at n = 128, 3.7x the algorithmic HBM traffic) and this library's own 8-launch blocked flow.
//
// Data flow (HBM sees each element once in, once out):
//   * columns [0, 64) ("left part") arrive by 1-D TMA bulk copies (cp.async.bulk, one per column, one mbarrier)
//     and stay in shared memory, column-major with a padded leading dimension, for the whole kernel;
//   * columns [64, n) ("right part") arrive by the same route in a 32-column staging buffer (prefetched at
//     kernel start / L2-prefetched), are picked up through the row permutation of the left part's pivots
//     straight into FP64 tensor-core accumulator fragments, receive U12 = L11^-1 A12 and A22 -= L21 U12
//     with the left part as the A operand, and the trailing block A22 is then factored in shared memory too;
//   * everything goes back with TMA bulk stores (U12 directly from the solving lanes).
// Factorisation of a resident block ("view"): 8-column sub-panels. ONE warp holds the sub-panel's rows in
// registers (lane = rows lane, lane+32, ...), so the per-column pivot chain -- search (REDUX on the high words,
// exact cascade on ties), winner broadcast by shuffles, reciprocal prepared speculatively by every lane, scale,
// rank-1 update -- runs without a single block barrier; interchanges are logical inside the sub-panel and become
// physical when it is written back. The other columns then get the sub-panel's net row permutation (<= 16 moves,
// all columns in parallel), the block row is solved against the 8x8 unit-lower block, and the rows below receive
// the rank-8 update as two DMMA.8x8x4 per 8x8 tile, C read-modify-written in shared memory.
// Every element receives fma(-l(i,k), u(k,j), a(i,j)) with k increasing (DMMA = chain of four FMAs, k increasing:
// tools/dmma_probe.cu), l = a * (1/pivot): bit-identical to oracle/lu_oracle.c.
#include "lu_chain.cuh"

namespace mb200 {

namespace {

constexpr int FT = 256;    // threads per CTA
// FLD = 130 (lu_chain.cuh): leading dimension of the left part: C fragments (col 2q, row g) conflict-free (130 = 2 mod 16)
constexpr int FLD2 = 66;   // leading dimension of the trailing block A22 (66 = 2 mod 16)
constexpr int FLDU = 68;   // U block rows: B fragments (k = 4s+q, col 8c+g) conflict-free (68 = 4 mod 16)
constexpr int FLDR = 136;  // staging of the right part, 32 columns: 32*136 = 64*68 doubles >= 64*66

struct FusedSmem {
    static constexpr int LDU = FLDU;
    double M[64 * FLD];    // left part, column c at M + c*FLD (rows 0..m-1)
    double R[32 * FLDR];   // right part staging (32 columns per round); then -U12 (64 x FLDU); then A22 (64 x FLD2)
    double Un[8 * FLDU];   // -U block row of the current sub-panel step
    double Us[8 * FLDU];   // block row handed to the solving lanes (right part)
    double L11[64];        // unit-lower 8x8 block of the current sub-panel, row i = multipliers of pivot row i
    unsigned long long bar[2];
    int ipiv[128];         // 0-based global pivot rows
    int nmoves;
    int info;
    unsigned char mdst[16], msrc[16];  // net row permutation of the last sub-panel: new[dst] = old[src]
    unsigned char perm[128];           // row at position r of the left part now = original row perm[r]
};
static_assert(sizeof(FusedSmem) <= 113 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(FT, 2)
lu_fused_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo, long batch,
                const int *__restrict__ index_list)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem &S = *reinterpret_cast<FusedSmem *>(smem_raw);

    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    if (m <= 0 || n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, q = lane & 3;
    double *__restrict__ A = dA[b];
    const int nl = n < 64 ? n : 64;  // columns of the left part
    const int nr = n - nl;           // columns of the right part
    const int nr0 = nr < 32 ? nr : 32;
    const bool bulk = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ld & 1) == 0) && ((m & 1) == 0);
    const unsigned colbytes = (unsigned)m * 8u;

    if (tid == 0) {
        f_mbar_init(&S.bar[0], 1);
        f_mbar_init(&S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.info = 0;
        S.nmoves = 0;
    }
    if (tid < 128) S.perm[tid] = (unsigned char)tid;
    __syncthreads();

    // ---- loads -------------------------------------------------------------------------------------------------
    if (bulk) {
        if (w == 0) {
            if (lane == 0) {
                f_mbar_expect(&S.bar[0], colbytes * (unsigned)nl);
                if (nr0 > 0) f_mbar_expect(&S.bar[1], colbytes * (unsigned)nr0);
            }
            __syncwarp();
            for (int c = lane; c < nl; c += 32) f_bulk_load(&S.M[c * FLD], A + (size_t)c * ld, colbytes, &S.bar[0]);
            for (int c = lane; c < nr0; c += 32) f_bulk_load(&S.R[c * FLDR], A + (size_t)(64 + c) * ld, colbytes, &S.bar[1]);
            for (int c = 32 + lane; c < nr; c += 32) f_bulk_prefetch_l2(A + (size_t)(64 + c) * ld, colbytes);
        }
        f_mbar_wait(&S.bar[0], 0);
    } else {
        for (int c = w; c < nl; c += FT / 32)
            for (int r = lane; r < m; r += 32) S.M[c * FLD + r] = A[r + (size_t)c * ld];
        for (int c = w; c < nr0; c += FT / 32)
            for (int r = lane; r < m; r += 32) S.R[c * FLDR + r] = A[r + (size_t)(64 + c) * ld];
        __syncthreads();
    }

    // ---- left part ---------------------------------------------------------------------------------------------
    factor_view<FLD, true, 7>(S, S.M, m, nl, 0, nullptr, 0, tid, lane, w);
    __syncthreads();

    // ---- right part --------------------------------------------------------------------------------------------
    if (nr > 0) {
        const int k0n = m < 64 ? m : 64;   // pivots of the left part = rows of U12
        const int NK = (k0n + 7) >> 3;
        const int ctn = (nr + 7) >> 3;
        double acc[2][8][2];
        // pick-up through the row permutation: tile rows t = w, w + 8; lane (g, q) holds C(8t+g, 8ct+2q), C(8t+g, 8ct+2q+1)
        int prow[2];
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
            const int row = 8 * (w + 8 * a2) + g;
            prow[a2] = (row < m) ? (int)S.perm[row] : 0;
        }
        if (bulk) f_mbar_wait(&S.bar[1], 0);
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
                acc[a2][ct][0] = S.R[(8 * ct + 2 * q) * FLDR + prow[a2]];
                acc[a2][ct][1] = S.R[(8 * ct + 2 * q + 1) * FLDR + prow[a2]];
            }
        __syncthreads();  // staging buffer free
        if (nr > 32) {
            const int nr1 = nr - 32;
            if (bulk) {
                if (w == 0) {
                    if (lane == 0) f_mbar_expect(&S.bar[1], colbytes * (unsigned)nr1);
                    __syncwarp();
                    for (int c = lane; c < nr1; c += 32)
                        f_bulk_load(&S.R[c * FLDR], A + (size_t)(96 + c) * ld, colbytes, &S.bar[1]);
                }
                f_mbar_wait(&S.bar[1], 1);
            } else {
                for (int c = w; c < nr1; c += FT / 32)
                    for (int r = lane; r < m; r += 32) S.R[c * FLDR + r] = A[r + (size_t)(96 + c) * ld];
                __syncthreads();
            }
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int ct = 0; ct < 4; ++ct) {
                    acc[a2][4 + ct][0] = S.R[(8 * ct + 2 * q) * FLDR + prow[a2]];
                    acc[a2][4 + ct][1] = S.R[(8 * ct + 2 * q + 1) * FLDR + prow[a2]];
                }
            __syncthreads();  // staging buffer free: it now takes -U12
        } else {
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int ct = 4; ct < 8; ++ct) acc[a2][ct][0] = acc[a2][ct][1] = 0.0;
        }
        double *Un64 = S.R;  // [64][FLDU]

        // block rows of U12 one after the other; every tile below gets its k = 8 update at once
#pragma unroll 1
        for (int K = 0; K < NK; ++K) {
            const int kb = (k0n - 8 * K) < 8 ? (k0n - 8 * K) : 8;
            if (w == K) {
                // this warp's slot-0 tiles ARE block row K: hand them to the solving lanes
#pragma unroll
                for (int ct = 0; ct < 8; ++ct)
                    *reinterpret_cast<double2 *>(&S.Us[g * FLDU + 8 * ct + 2 * q]) = make_double2(acc[0][ct][0], acc[0][ct][1]);
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = lane + 32 * h;
                    if (c < nr) {
                        double x[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = (i < kb) ? S.Us[i * FLDU + c] : 0.0;
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
#pragma unroll
                            for (int i = k + 1; i < 8; ++i)
                                if (i < kb) x[i] = fma(-S.M[(8 * K + k) * FLD + 8 * K + i], x[k], x[i]);
                        }
                        double *gcol = A + (size_t)(8 * K) + (size_t)(64 + c) * ld;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (i < kb) {
                                Un64[(8 * K + i) * FLDU + c] = -x[i];
                                gcol[i] = x[i];  // final U12
                            }
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2) {
                const int t = w + 8 * a2;
                if (t > K && 8 * t < m && kb == 8) {
                    const double af0 = S.M[(8 * K + q) * FLD + 8 * t + g];
                    const double af1 = S.M[(8 * K + 4 + q) * FLD + 8 * t + g];
#pragma unroll
                    for (int ct = 0; ct < 8; ++ct) {
                        if (ct < ctn) {
                            const double bf0 = Un64[(8 * K + q) * FLDU + 8 * ct + g];
                            const double bf1 = Un64[(8 * K + 4 + q) * FLDU + 8 * ct + g];
                            f_dmma(acc[a2][ct][0], acc[a2][ct][1], af0, bf0);
                            f_dmma(acc[a2][ct][0], acc[a2][ct][1], af1, bf1);
                        }
                    }
                }
            }
        }
        if (m > 64) {
            __syncthreads();  // every read of -U12 done: the buffer takes A22 (view rows 8w+g)
            double *A22 = S.R;
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
                A22[(8 * ct + 2 * q) * FLD2 + 8 * w + g] = acc[1][ct][0];
                A22[(8 * ct + 2 * q + 1) * FLD2 + 8 * w + g] = acc[1][ct][1];
            }
            __syncthreads();
            factor_view<FLD2, false, 7>(S, A22, m - 64, nr, 64, S.M + 64, 64, tid, lane, w);
            __syncthreads();
        }
    }

    // ---- stores ------------------------------------------------------------------------------------------------
    const int mn = m < n ? m : n;
    if (bulk) {
        f_fence_async_smem();  // generic-proxy writes of shared memory -> visible to the bulk-copy engine
        __syncthreads();
        if (w == 0) {
            for (int c = lane; c < nl; c += 32) f_bulk_store(A + (size_t)c * ld, &S.M[c * FLD], colbytes);
            if (m > 64)
                for (int c = lane; c < nr; c += 32)
                    f_bulk_store(A + 64 + (size_t)(64 + c) * ld, &S.R[c * FLD2], colbytes - 512u);
        }
        for (int i = tid; i < mn; i += FT) dipiv[b][i] = S.ipiv[i] + 1;
        if (tid == 0) dinfo[b] = S.info;
        if (w == 0) f_bulk_commit_wait();
    } else {
        __syncthreads();
        for (int c = w; c < nl; c += FT / 32)
            for (int r = lane; r < m; r += 32) A[r + (size_t)c * ld] = S.M[c * FLD + r];
        if (m > 64)
            for (int c = w; c < nr; c += FT / 32)
                for (int r = lane; r < m - 64; r += 32) A[64 + r + (size_t)(64 + c) * ld] = S.R[c * FLD2 + r];
        for (int i = tid; i < mn; i += FT) dipiv[b][i] = S.ipiv[i] + 1;
        if (tid == 0) dinfo[b] = S.info;
    }
}

// ---- 32-column panel of the left-looking slab driver, panels of at most 128 rows ------------------------------------
// Replaces panel_kernel (lu_blocked.cu: one thread per row, two block barriers and two arg-max cascades per column,
// ~21k warp-instructions per 128 x 32 panel, 54% issue-slot utilisation) for short panels: the slab's rows j.. arrive
// by TMA in shared memory, ONE warp runs the pivot chains of the four 8-column sub-panels, a second warp applies each
// finished sub-panel to the rest of the slab (permutation, block-row solve, rank-8 DMMA update; next sub-panel first).
// Two warps and LD*256 bytes per CTA: 6 (128 rows) .. 16 (32 rows) pivot chains per SM.
// Outputs as panel_kernel's: factors in final row order, global pivots, info, and the step permutation record sinv.
template <int LD>
struct PanelSmem {
    static constexpr int LDU = 36;  // 32 columns (+4: B fragments conflict-free)
    double V[32 * LD];
    double Un[8 * LDU];
    double L11[64];
    unsigned long long bar[1];
    int ipiv[128];
    int nmoves;
    int info;
    unsigned char mdst[16], msrc[16];
    unsigned char perm[128];  // row at panel position p now = panel row perm[p] before the panel
};

template <int LD, int MINB>
__global__ void __launch_bounds__(64, MINB)
panel_chain_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo, int j, long batch,
                   const int *__restrict__ index_list, unsigned short *__restrict__ sinv, int sinv_rows, int sinv_blocks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PanelSmem<LD> &S = *reinterpret_cast<PanelSmem<LD> *>(smem_raw);
    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < 32 ? (mn - j) : 32;
    const int mp = m - j;  // <= LD - 2
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    double *__restrict__ A = dA[b] + (size_t)j + (size_t)j * ld;  // panel origin
    const bool bulk = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ld & 1) == 0) && ((mp & 1) == 0);
    const unsigned colbytes = (unsigned)mp * 8u;
    if (tid == 0) {
        f_mbar_init(&S.bar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.info = 0;
        S.nmoves = 0;
    }
    for (int i = tid; i < 128; i += 64) S.perm[i] = (unsigned char)i;
    __syncthreads();
    if (bulk) {
        if (w == 0) {
            if (lane == 0) f_mbar_expect(&S.bar[0], colbytes * (unsigned)jb);
            __syncwarp();
            if (lane < jb) f_bulk_load(&S.V[lane * LD], A + (size_t)lane * ld, colbytes, &S.bar[0]);
        }
        f_mbar_wait(&S.bar[0], 0);
    } else {
        // any alignment: 8-byte cp.asyncs, all in flight at once (plain loads here were a chain of dependent
        // load -> store pairs per thread: the n = 127 call spent a third of its time in them)
        for (int c = w; c < jb; c += 2)
            for (int r = lane; r < mp; r += 32) cp_async8(&S.V[c * LD + r], A + r + (size_t)c * ld, true);
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    // S.ipiv / S.info are panel-relative here (row_off = 0): j is added on the way out
    factor_view<LD, true, 1>(S, S.V, mp, jb, 0, nullptr, 0, tid, lane, w);
    __syncthreads();
    if (bulk) {
        f_fence_async_smem();
        __syncthreads();
        if (w == 0 && lane < jb) f_bulk_store(A + (size_t)lane * ld, &S.V[lane * LD], colbytes);
    } else {
        for (int c = w; c < jb; c += 2)
            for (int r = lane; r < mp; r += 32) A[r + (size_t)c * ld] = S.V[c * LD + r];
    }
    if (tid < jb) dipiv[b][j + tid] = j + S.ipiv[tid] + 1;
    if (sinv) {
        unsigned short *sv = sinv + ((size_t)slot * sinv_blocks + (j >> 5)) * sinv_rows + j;
        for (int p = tid; p < mp; p += 64) sv[p] = (unsigned short)(j + S.perm[p]);
    }
    if (tid == 0) {
        const int info = S.info ? j + S.info : 0;
        if (j == 0) dinfo[b] = info;
        else if (info && dinfo[b] == 0) dinfo[b] = info;
    }
    if (bulk && w == 0) f_bulk_commit_wait();
}

__global__ void rcp_selftest_kernel(long n, unsigned long long seed, unsigned long long *bad)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // splitmix64 -> sign, exponent in [-990, 990], 52 random mantissa bits
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const unsigned long long mant = z & 0xFFFFFFFFFFFFFull;
    const unsigned long long ex = 1023ull - 990ull + ((z >> 52) & 0x7FFull) % 1981ull;
    const unsigned long long bits = ((z >> 63) << 63) | (ex << 52) | mant;
    const double x = __longlong_as_double((long long)bits);
    const double a = rcp_fast_f64(x), b = 1.0 / x;
    if (__double_as_longlong(a) != __double_as_longlong(b)) atomicAdd(bad, 1ull);
}

}  // namespace

// Number of inputs (of n pseudo-random normal doubles with exponents in [-990, 990]) on which the chain's inline
// reciprocal differs from the IEEE division 1.0/x. Expected: 0.
long rcp_selftest_run(long n, cudaStream_t s)
{
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return -1;
    cudaMemsetAsync(d, 0, sizeof(h), s);
    rcp_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, 0x1234567ull, d);
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(d);
    return cudaGetLastError() == cudaSuccess ? (long)h : -1;
}

// 32-column panel at (j, j) of matrices with 97..128 rows below j; T = roundup(rows, 32). -100: not covered.
magma_int_t panel_chain_launch(const Dims &d, double **dA, int **dipiv, int *dinfo, int j, int T, long batch, const int *il,
                               cudaStream_t s, unsigned short *sinv, int sinv_rows, int sinv_blocks)
{
    if (T > 128 || batch <= 0) return -100;
#define MB200_PC(LDV, MINB)                                                                                                   \
    do {                                                                                                                      \
        static DevOnce once;                                                                                                  \
        smem_optin(once, panel_chain_kernel<LDV, MINB>, sizeof(PanelSmem<LDV>));                                              \
        panel_chain_kernel<LDV, MINB><<<(unsigned)batch, 64, sizeof(PanelSmem<LDV>), s>>>(d, dA, dipiv, dinfo, j, batch, il, \
                                                                                        sinv, sinv_rows, sinv_blocks);     \
    } while (0)
    if (T > 96) MB200_PC(130, 6);
    else if (T > 64) MB200_PC(98, 8);
    else if (T > 32) MB200_PC(66, 12);
    else MB200_PC(34, 16);
#undef MB200_PC
    count_launch();
    MB200_CHECK_LAUNCH("panel_chain_kernel");
    return 0;
}

magma_int_t lu_fused_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, long batch,
                            const int *index_list, cudaStream_t s)
{
    if (max_m > 128 || max_n > 128) return -100;
    if (batch <= 0) return 0;
    static DevOnce once;
    smem_optin(once, lu_fused_kernel, sizeof(FusedSmem));
    lu_fused_kernel<<<(unsigned)batch, FT, sizeof(FusedSmem), s>>>(d, dA, dipiv, dinfo, batch, index_list);
    count_launch();
    MB200_CHECK_LAUNCH("lu_fused_kernel");
    return 0;
}

}  // namespace mb200
