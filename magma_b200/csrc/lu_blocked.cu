// Blocked LU for any m x n, two drivers over the same register panel kernel.
//
// Replaces the reference's host recursion (src/zgetrf_batched.cpp:149-203 ->
// src/zgetrf_panel_batched.cpp:101-196 -> src/zgetf2_batched.cpp:95-287), which issues 42/105/231
// launches per call at n = 128/256/512: fused panel, setup_pivinfo, adjust_ipiv, 2x laswp,
// recursive trsm, cublasDgemmBatched (+3 pointer-displacement kernels each).
//
//   panel_kernel  : one CTA per matrix, panel rows in registers (thread = R rows x W columns),
//                   CREDUX + shared-memory two-level pivot search, lazy row interchanges; writes
//                   the factored panel in final row order, ipiv (global indices), the step's
//                   net row permutation as a 512-byte record (right-looking driver) and as a 16-bit
//                   position table (left-looking driver) -- the threads know both from their
//                   final positions, there is no setup_pivinfo / adjust_ipiv pass. Instantiated
//                   per panel height with an 80-register budget (2..12 CTAs per SM).
//
// LEFT-LOOKING driver (at most 512 rows: the default): per 32-column slab
//   left_update_kernel<NW> : one CTA owns the slab for its whole history -- reads it once through
//                   the composed row permutation, keeps it as DMMA accumulator fragments, and for
//                   every earlier panel solves the block row and subtracts L(:,K) U(K,J) on the
//                   FP64 tensor pipe; L streams through a CTA-wide cp.async ring handed over by
//                   mbarriers. No physical interchange touches the trailing matrix.
//   laswp_left_sinv_kernel : the deferred interchanges of the L columns, one pass at the end.
//
// RIGHT-LOOKING driver (more rows, or magma_b200_set_tier(6)): per 64 columns
//   swap_trsm_kernel : one CTA per (matrix, 64-column tile): applies the permutation to the tile
//                   (left tiles: interchanges only), and for tiles right of the panel solves the
//                   W x 64 block row against the unit-lower L11 and stores U12;
//   gemm_dmma_kernel / gemm_kernel : C(r,c) = fma(-L21(r,k), U12(k,c), C(r,c)), k increasing, one
//                   128x64 tile per CTA on the tensor pipe (DFMA fallback: tier 4);
//   update_strip_kernel, laswp_left_kernel : short trailing strips in shared memory; deferred
//                   interchanges replayed from ipiv.
// All arithmetic is in the canonical order of oracle/lu_oracle.c: bit-identical factors.
#include "lu_chain.cuh"
#include <map>
#include <mutex>
#include <utility>

namespace mb200 {

namespace {

constexpr int NOPOS = NOPOS_I;
constexpr int KB = 32;  // widest panel

// Per-matrix, per-step record written by the panel kernel, read by swap_trsm_kernel.
struct PivRec {
    int top_src[KB];   // original (panel-relative) row that ends at top position p
    int down_dst[KB];  // positions >= jb that receive an original top row ...
    int down_src[KB];  // ... and which one
    int n_down;
    int pad[31];
};
static_assert(sizeof(PivRec) == 512, "PivRec layout");

// -------------------------------------------------------------------------------------------
// Panel factorisation, registers. Thread t owns panel rows t, t+T, ... (R of them), W columns.
// -------------------------------------------------------------------------------------------
template <int R, int W, int MAXT = 512, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
panel_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo,
             PivRec *__restrict__ recs, int j, long batch, const int *__restrict__ index_list,
             unsigned short *__restrict__ sinv = nullptr, int sinv_rows = 0, int sinv_blocks = 0, int nopiv = 0)
{
    // nopiv (R == 1 only): the diagonal element is the pivot whatever its size (magma_dgetrf_nopiv_batched,
    // magmablas/zgetf2_nopiv_kernels.cu:22-72); a zero diagonal sets info and leaves its column unscaled, the
    // rank-1 update still runs (LAPACK's dgetf2 convention; the reference stops factoring at that point).
    // sinv (left-looking driver): sinv[panel][p] = position BEFORE this panel of the row that is at position p
    // after it (absolute rows, entries >= j only)
    __shared__ unsigned long long cbits[2][32];
    __shared__ int cpos[2][32];
    __shared__ __align__(16) double prow[2][W + 2];  // [W] = 1/pivot
    __shared__ int sipiv[W];
    __shared__ int s_ndown;

    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;  // unused tail of a vbatched index list
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;
    const int mp = m - j;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    double *__restrict__ A = dA[b] + (size_t)j + (size_t)j * ld;  // panel origin
    PivRec &rec = recs[slot];

    double a[R][W];
    int pos[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int r = tid + k * T;
        pos[k] = (r < mp) ? r : NOPOS;
#pragma unroll
        for (int c = 0; c < W; ++c) a[k][c] = (r < mp && c < jb) ? A[r + (size_t)c * ld] : 1.0;
    }
    if (tid == 0) s_ndown = 0;
    int info = 0;

#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i < jb) {
            // local best over this thread's rows
            unsigned long long lb = 0;
            int lp = NOPOS;
            int lk = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool act = (pos[k] >= i) && (pos[k] != NOPOS);
                const unsigned long long v =
                    (unsigned long long)__double_as_longlong(a[k][i]) & 0x7fffffffffffffffull;
                if (act && (v > lb || lp == NOPOS || (v == lb && pos[k] < lp))) {
                    lb = v;
                    lp = pos[k];
                    lk = k;
                }
            }
            // every thread prepares the reciprocal of its own candidate while the search runs
            double cv = a[0][i];
#pragma unroll
            for (int k = 1; k < R; ++k)
                if (lk == k) cv = a[k][i];
            const double rinv = 1.0 / cv;

            int wp = i;  // nopiv: the diagonal
            if (!nopiv) {  // kernel-uniform
                // the winning lane hands over its own candidate (no shuffle of the winner's values to every lane)
                const int wl = warp_argmax_lane(lb, lp);
                if (nw == 1) {  // single-warp panel: decided
                    wp = __shfl_sync(0xffffffffu, lp, wl);
                } else {
                    if (lane == wl) {
                        cbits[i & 1][wid] = lb;
                        cpos[i & 1][wid] = lp;
                    }
                    __syncthreads();
                    const unsigned long long eb = (lane < nw) ? cbits[i & 1][lane] : 0ull;
                    const int ep = (lane < nw) ? cpos[i & 1][lane] : NOPOS;
                    wp = __shfl_sync(0xffffffffu, ep, warp_argmax_lane(eb, ep));
                }
            }
            const int ppos = wp;  // panel-relative position of the pivot row (>= i)
            if (tid == 0) sipiv[i] = ppos;
            if (lp == ppos) {
                // this thread owns the pivot row (its local winner): publish row and reciprocal
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (lk == k) {
#pragma unroll
                        for (int c = 0; c < W; ++c)
                            if (c >= i) prow[i & 1][c] = a[k][c];
                    }
                }
                prow[i & 1][W] = rinv;
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (pos[k] == ppos) pos[k] = i;
                else if (pos[k] == i) pos[k] = ppos;
            }
            __syncthreads();
            const double piv = prow[i & 1][i];
            if (piv == 0.0 && info == 0) info = i + 1;
            if (piv != 0.0 || nopiv) {
                const double r = (piv != 0.0) ? prow[i & 1][W] : 1.0;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (pos[k] > i && pos[k] != NOPOS) {
                        const double l = a[k][i] * r;
                        a[k][i] = l;
#pragma unroll
                        for (int c = 0; c < W; ++c)
                            if (c > i) a[k][c] = fma(-l, prow[i & 1][c], a[k][c]);
                    }
                }
            }
        }
    }

    // factored panel, rows in final order; permutation record
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (pos[k] != NOPOS) {
#pragma unroll
            for (int c = 0; c < W; ++c)
                if (c < jb) A[pos[k] + (size_t)c * ld] = a[k][c];
            const int orig = tid + k * T;
            if (sinv)
                sinv[((size_t)slot * sinv_blocks + (j >> 5)) * sinv_rows + j + pos[k]] = (unsigned short)(j + orig);
            if (pos[k] < jb) {
                rec.top_src[pos[k]] = orig;
            } else if (pos[k] != orig) {  // an original top row that went down
                const int e = atomicAdd(&s_ndown, 1);
                rec.down_dst[e] = pos[k];
                rec.down_src[e] = orig;
            }
        }
    }
    if (!nopiv && tid < jb) dipiv[b][j + tid] = j + sipiv[tid] + 1;
    __syncthreads();
    if (tid == 0) {
        rec.n_down = s_ndown;
        if (j == 0) dinfo[b] = info;
        else if (info && dinfo[b] == 0) dinfo[b] = j + info;
    }
}

// -------------------------------------------------------------------------------------------
// Tall panels of the left-looking driver (129..512 rows) in two 16-column halves.
// panel_kernel keeps a row's 32 columns in registers: 80 registers x 512 threads leave ONE pivot chain per SM at 512 rows
// (two at 384, three at 256), and a tall panel's column costs ~1400 cycles (two CTA barriers) whatever else the SM could do:
// 21% of the n = 512 call and of vbatched config 4. Here a thread holds 16 columns at a time:
//   A  columns 0..15 factored as panel_kernel does (lazy positions, row published through shared memory), the results PARKED
//      in shared memory (16 doubles per row: 64 KB at 512 rows; final places are not known before B);
//   U  the 16 pivot rows publish their multipliers (L11) and their raw second halves; eight warps solve U12 = L11^-1 A12
//      (two columns per warp, lane = row of the block, one shuffle per step); every other row applies the rank-16 update
//      with its parked multipliers and U12 broadcast from shared memory -- k increasing per element, zero-pivot steps
//      skipped: the canonical order;
//   B  columns 16..31 factored the same way; both halves are then stored at the final places.
// Half the registers per thread: two chains per SM at 512 rows, three at 384, four at 256. Outputs as panel_kernel's in the
// left-looking driver (factors in final row order, pivots, step permutation record, info).
// -------------------------------------------------------------------------------------------
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
panel2_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo, int j, long batch,
              const int *__restrict__ index_list, unsigned short *__restrict__ sinv, int sinv_rows, int sinv_blocks)
{
    constexpr int H = 16;
    __shared__ unsigned long long cbits[2][32];
    __shared__ int cpos[2][32];
    __shared__ __align__(16) double prow[2][H + 2];  // [H] = 1/pivot
    __shared__ int sipiv[2 * H];
    __shared__ __align__(16) double U12[H][H];  // second halves of the 16 pivot rows, raw, then solved in place
    __shared__ double L11[H][H + 1];             // L11[k][p], p < k: multipliers of the row that became pivot row k
    __shared__ unsigned zmask_s;                 // bit i: step i met an exactly zero pivot (no scaling, no update)
    extern __shared__ __align__(16) unsigned char p2_raw[];
    double *park = reinterpret_cast<double *>(p2_raw);  // park[c * T + row]: the first half's results, rows in ORIGINAL order

    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < 32 ? (mn - j) : 32;
    const int jb1 = jb < H ? jb : H, jb2 = jb - jb1;
    const int T = (int)blockDim.x;
    const int mp = m - j;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (int)blockDim.x >> 5;
    double *__restrict__ A = dA[b] + (size_t)j + (size_t)j * ld;  // panel origin
    const bool mine = tid < mp;
    int pos = mine ? tid : NOPOS;  // current position of this thread's row (lazy interchanges)
    double a[H];
#pragma unroll
    for (int c = 0; c < H; ++c) a[c] = (mine && c < jb1) ? A[tid + (size_t)c * ld] : 1.0;
    if (tid == 0) zmask_s = 0;
    int info = 0;

    // cnt column steps on a[0..cnt): step i = ibase + ii works on a[ii]
    auto chain = [&](const int ibase, const int cnt) {
#pragma unroll
        for (int ii = 0; ii < H; ++ii) {
            if (ii < cnt) {
                const int i = ibase + ii;
                const bool act = (pos >= i) && (pos != NOPOS);
                const unsigned long long lb = act ? ((unsigned long long)__double_as_longlong(a[ii]) & 0x7fffffffffffffffull) : 0ull;
                const int lp = act ? pos : NOPOS;
                const double rinv = rcp_fast_f64(a[ii]);  // every thread inverts its own candidate while the search runs
                const int wl = warp_argmax_lane(lb, lp);
                if (lane == wl) {
                    cbits[i & 1][wid] = lb;
                    cpos[i & 1][wid] = lp;
                }
                __syncthreads();
                const unsigned long long eb = (lane < nw) ? cbits[i & 1][lane] : 0ull;
                const int ep = (lane < nw) ? cpos[i & 1][lane] : NOPOS;
                const int ppos = __shfl_sync(0xffffffffu, ep, warp_argmax_lane(eb, ep));  // position of the pivot row (>= i)
                if (tid == 0) sipiv[i] = ppos;
                if (lp == ppos) {
#pragma unroll
                    for (int c = 0; c < H; ++c)
                        if (c >= ii) prow[i & 1][c] = a[c];
                    // the inline reciprocal equals IEEE 1/x for |x| in [2^-999, 2^993); outside (one thread, rare) divide
                    prow[i & 1][H] = rcp_fast_ok((unsigned)(lb >> 32)) ? rinv : 1.0 / a[ii];
                }
                if (pos == ppos) pos = i;
                else if (pos == i) pos = ppos;
                __syncthreads();
                const double piv = prow[i & 1][ii];
                if (piv == 0.0) {
                    if (info == 0) info = i + 1;
                    if (tid == 0) zmask_s |= 1u << i;
                } else if (pos > i && pos != NOPOS) {
                    const double l = a[ii] * prow[i & 1][H];
                    a[ii] = l;
#pragma unroll
                    for (int c = 0; c < H; ++c)
                        if (c > ii) a[c] = fma(-l, prow[i & 1][c], a[c]);
                }
            }
        }
    };

    chain(0, jb1);
    if (jb2 == 0) {
        if (mine) {
#pragma unroll
            for (int c = 0; c < H; ++c)
                if (c < jb1) A[pos + (size_t)c * ld] = a[c];
        }
    } else {
        // jb1 == 16 here. Park the first half; the pivot rows publish their multipliers.
        if (mine) {
#pragma unroll
            for (int c = 0; c < H; ++c) park[c * T + tid] = a[c];
            if (pos < H) {
#pragma unroll
                for (int p = 0; p < H; ++p)
                    if (p < pos) L11[pos][p] = a[p];
            }
        }
#pragma unroll
        for (int c = 0; c < H; ++c) a[c] = (mine && c < jb2) ? A[tid + (size_t)(H + c) * ld] : 1.0;
        if (mine && pos < H) {
#pragma unroll
            for (int c = 0; c < H; ++c) U12[pos][c] = a[c];
        }
        __syncthreads();
        const unsigned zm = zmask_s;
        for (int cp = wid; cp < 8; cp += nw) {  // U12 = L11^-1 A12 in place: a warp takes columns 2 cp, 2 cp + 1; lane = (column, row k)
            const int k = lane & 15, c = 2 * cp + (lane >> 4);
            double x = U12[k][c];
#pragma unroll
            for (int p = 0; p < H - 1; ++p) {
                const double xp = __shfl_sync(0xffffffffu, x, (lane & 16) | p);  // x(p) is final: every step q < p has been applied
                if (k > p && !((zm >> p) & 1u)) x = fma(-L11[k][p], xp, x);
            }
            U12[k][c] = x;
        }
        __syncthreads();
        if (mine) {
            if (pos < H) {
#pragma unroll
                for (int c = 0; c < H; ++c) a[c] = U12[pos][c];
            } else {
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    if (!((zm >> k) & 1u)) {
                        const double l = park[k * T + tid];  // this row's parked multiplier of step k
#pragma unroll
                        for (int c = 0; c < H; ++c) a[c] = fma(-l, U12[k][c], a[c]);
                    }
                }
            }
        }
        chain(H, jb2);
        if (mine) {
#pragma unroll
            for (int c = 0; c < H; ++c)
                if (c < jb2) A[pos + (size_t)(H + c) * ld] = a[c];
#pragma unroll
            for (int c = 0; c < H; ++c) A[pos + (size_t)c * ld] = park[c * T + tid];
        }
    }
    if (mine) sinv[((size_t)slot * sinv_blocks + (j >> 5)) * sinv_rows + j + pos] = (unsigned short)(j + tid);
    if (tid < jb) dipiv[b][j + tid] = j + sipiv[tid] + 1;
    if (tid == 0) {
        if (j == 0) dinfo[b] = info;
        else if (info && dinfo[b] == 0) dinfo[b] = j + info;
    }
}

// -------------------------------------------------------------------------------------------
// Panel factorisation straight on global memory: correctness fallback for panels taller than
// the register kernel covers (m - j > 8192). One CTA per matrix, W columns. Interchanges are
// physical here, so the record lists are rebuilt from the pivots by thread 0.
// -------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
panel_global_kernel(Dims d, double **dA, int **dipiv, int *dinfo, PivRec *recs, int j, long batch,
                    const int *index_list)
{
    __shared__ unsigned long long cbits[8];
    __shared__ int cpos[8];
    __shared__ int spiv;
    __shared__ int lp_[W];
    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;
    const int mp = m - j;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *A = dA[b] + (size_t)j + (size_t)j * ld;
    int info = 0;
    for (int i = 0; i < jb; ++i) {
        unsigned long long lb = 0;
        int lp = NOPOS;
        for (int r = i + tid; r < mp; r += 256) {
            const unsigned long long v =
                (unsigned long long)__double_as_longlong(A[r + (size_t)i * ld]) & 0x7fffffffffffffffull;
            if (v > lb || lp == NOPOS) {  // rows visited in increasing order: strict > keeps the first
                lb = v;
                lp = r;
            }
        }
        unsigned long long wb;
        int wp;
        warp_argmax(lb, lp, wb, wp);
        if (lane == 0) {
            cbits[wid] = wb;
            cpos[wid] = wp;
        }
        __syncthreads();
        if (wid == 0) {
            unsigned long long eb = (lane < 8) ? cbits[lane] : 0ull;
            int ep = (lane < 8) ? cpos[lane] : NOPOS;
            warp_argmax(eb, ep, wb, wp);
            if (lane == 0) spiv = wp;
        }
        __syncthreads();
        const int p = spiv;
        if (tid == 0) {
            dipiv[b][j + i] = j + p + 1;
            lp_[i] = p;
        }
        if (p != i && tid < jb) {
            const double t0 = A[i + (size_t)tid * ld];
            A[i + (size_t)tid * ld] = A[p + (size_t)tid * ld];
            A[p + (size_t)tid * ld] = t0;
        }
        __syncthreads();
        const double piv = A[i + (size_t)i * ld];
        if (piv != 0.0) {
            const double r = 1.0 / piv;
            for (int rr = i + 1 + tid; rr < mp; rr += 256) {
                const double l = A[rr + (size_t)i * ld] * r;
                A[rr + (size_t)i * ld] = l;
                for (int c = i + 1; c < jb; ++c)
                    A[rr + (size_t)c * ld] = fma(-l, A[i + (size_t)c * ld], A[rr + (size_t)c * ld]);
            }
        } else if (info == 0) {
            info = i + 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        // net permutation of the jb sequential interchanges (tiny: jb <= 8 here)
        PivRec &rec = recs[slot];
        int top[W], epos[W], econt[W], ne = 0;
        for (int k = 0; k < jb; ++k) top[k] = k;
        for (int i = 0; i < jb; ++i) {
            const int p = lp_[i];
            if (p == i) continue;
            if (p < jb) {
                const int t = top[i];
                top[i] = top[p];
                top[p] = t;
                continue;
            }
            int e = -1;
            for (int q = 0; q < ne; ++q)
                if (epos[q] == p) e = q;
            if (e < 0) {
                e = ne++;
                epos[e] = p;
                econt[e] = p;
            }
            const int t = top[i];
            top[i] = econt[e];
            econt[e] = t;
        }
        for (int k = 0; k < jb; ++k) rec.top_src[k] = top[k];
        for (int e = 0; e < ne; ++e) {
            rec.down_dst[e] = epos[e];
            rec.down_src[e] = econt[e];
        }
        rec.n_down = ne;
        if (j == 0) dinfo[b] = info;
        else if (info && dinfo[b] == 0) dinfo[b] = j + info;
    }
}

// -------------------------------------------------------------------------------------------
// Interchanges on a 64-column tile (+ U12 = L11^-1 * block row for tiles right of the panel).
// -------------------------------------------------------------------------------------------
constexpr int TN = 64;
constexpr int ST_THREADS = 128;
constexpr int TNP = TN + 1;

template <int W>
__global__ void __launch_bounds__(ST_THREADS)
swap_trsm_kernel(Dims d, double **__restrict__ dA, const PivRec *__restrict__ recs, int j, int right_tiles,
                 int left_tiles, int left_begin, int pre_k0, long batch, const int *__restrict__ index_list)
{
    // left tiles cover columns [left_begin, j). pre_k0 >= 0: this step's trailing update by the 32 columns
    // at pre_k0 was postponed (64-wide pairing, lu_blocked_launch): the block row gets it here, after the
    // interchanges and before the triangular solve, so every element still sees k increasing.
    __shared__ double T0[KB * TNP];       // original top rows   [k*TNP + c]
    __shared__ double Bs[KB * TNP];       // permuted top block  [p*TNP + c]
    __shared__ double Ls[KB * (KB + 1)];  // L11                 [i*(KB+1) + k]
    __shared__ int s_top[KB], s_ddst[KB], s_dsrc[KB];

    const int tiles = right_tiles + left_tiles;
    const long slot = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;
    int c0, c1;
    bool right;
    if (tile < right_tiles) {
        right = true;
        c0 = j + jb + tile * TN;
        c1 = c0 + TN < n ? c0 + TN : n;
    } else {
        right = false;
        c0 = left_begin + (tile - right_tiles) * TN;
        c1 = c0 + TN < j ? c0 + TN : j;
    }
    if (c0 >= c1) return;
    const int wt = c1 - c0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *__restrict__ A = dA[b];
    double *Ap = A + (size_t)j;  // row origin of the panel
    const PivRec &rec = recs[slot];
    if (tid < KB) {
        s_top[tid] = rec.top_src[tid];
        s_ddst[tid] = rec.down_dst[tid];
        s_dsrc[tid] = rec.down_src[tid];
    }
    const int nd = rec.n_down;
    // original top rows (coalesced: k is the fast index)
#pragma unroll 4
    for (int idx = tid; idx < KB * TN; idx += ST_THREADS) {
        const int k = idx & 31, c = idx >> 5;
        if (k < jb && c < wt) T0[k * TNP + c] = Ap[k + (size_t)(c0 + c) * ld];
    }
    if (right) {
        for (int idx = tid; idx < KB * KB; idx += ST_THREADS) {
            const int i = idx & 31, k = idx >> 5;
            if (i < jb && k < jb) Ls[i * (KB + 1) + k] = Ap[i + (size_t)(j + k) * ld];
        }
    }
    __syncthreads();
    // permuted top block: rows that come up from below are read before their slots are overwritten
#pragma unroll 4
    for (int idx = tid; idx < KB * TN; idx += ST_THREADS) {
        const int c = idx & 63, p = idx >> 6;
        if (p < jb && c < wt) {
            const int src = s_top[p];
            Bs[p * TNP + c] = (src < jb) ? T0[src * TNP + c] : Ap[src + (size_t)(c0 + c) * ld];
        }
    }
    __syncthreads();
    if (lane < nd) {
        const int dst = s_ddst[lane], src = s_dsrc[lane];
        for (int c = wid; c < wt; c += ST_THREADS / 32) Ap[dst + (size_t)(c0 + c) * ld] = T0[src * TNP + c];
    }
    if (!right) {
#pragma unroll 4
        for (int idx = tid; idx < KB * TN; idx += ST_THREADS) {
            const int k = idx & 31, c = idx >> 5;
            if (k < jb && c < wt && s_top[k] != k) Ap[k + (size_t)(c0 + c) * ld] = Bs[k * TNP + c];
        }
        return;
    }
    if (pre_k0 >= 0) {
        // postponed update of the block row: Bs(i, c) -= sum_k A(j+i, pre_k0+k) * A(pre_k0+k, c0+c), k increasing.
        // T0 is free now: it takes U(pre_k0.., tile) (k fast); Lp goes where L11 will be loaded afterwards.
        // (Prefetching all of this with cp.async into separate buffers at kernel start was slower: 67 KB of
        // shared memory per CTA cost a quarter of the occupancy.)
        __syncthreads();
        double *Lp = Ls;  // [i*(KB+1) + k]
        for (int idx = tid; idx < KB * KB; idx += ST_THREADS) {
            const int i = idx & 31, k = idx >> 5;
            Lp[i * (KB + 1) + k] = (i < jb) ? Ap[i + (size_t)(pre_k0 + k) * ld] : 0.0;
        }
#pragma unroll 4
        for (int idx = tid; idx < KB * TN; idx += ST_THREADS) {
            const int k = idx & 31, c = idx >> 5;
            if (c < wt) T0[k * TNP + c] = A[(size_t)(pre_k0 + k) + (size_t)(c0 + c) * ld];
        }
        __syncthreads();
        {
            const int col = tid & (TN - 1), half = tid >> 6;  // 64 columns x 2 halves of the 32 rows
            if (col < wt) {
#pragma unroll 1
                for (int i = half * 16; i < half * 16 + 16; i += 4) {
                    double x0 = Bs[i * TNP + col], x1 = Bs[(i + 1) * TNP + col], x2 = Bs[(i + 2) * TNP + col],
                           x3 = Bs[(i + 3) * TNP + col];
#pragma unroll 8
                    for (int k = 0; k < KB; ++k) {
                        const double u = T0[k * TNP + col];
                        x0 = fma(-Lp[i * (KB + 1) + k], u, x0);
                        x1 = fma(-Lp[(i + 1) * (KB + 1) + k], u, x1);
                        x2 = fma(-Lp[(i + 2) * (KB + 1) + k], u, x2);
                        x3 = fma(-Lp[(i + 3) * (KB + 1) + k], u, x3);
                    }
                    if (i < jb) Bs[i * TNP + col] = x0;
                    if (i + 1 < jb) Bs[(i + 1) * TNP + col] = x1;
                    if (i + 2 < jb) Bs[(i + 2) * TNP + col] = x2;
                    if (i + 3 < jb) Bs[(i + 3) * TNP + col] = x3;
                }
            }
        }
        __syncthreads();
        // now L11 of this step
        for (int idx = tid; idx < KB * KB; idx += ST_THREADS) {
            const int i = idx & 31, k = idx >> 5;
            if (i < jb && k < jb) Ls[i * (KB + 1) + k] = Ap[i + (size_t)(j + k) * ld];
        }
        __syncthreads();
    }
    // U12 = L11^-1 * Bs: one thread per column, canonical order
    if (tid < wt) {
        double x[W];
#pragma unroll
        for (int i = 0; i < W; ++i) x[i] = (i < jb) ? Bs[i * TNP + tid] : 0.0;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (k < jb) {
#pragma unroll
                for (int i = 0; i < W; ++i)
                    if (i > k && i < jb) x[i] = fma(-Ls[i * (KB + 1) + k], x[k], x[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i)
            if (i < jb) Bs[i * TNP + tid] = x[i];
    }
    __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < KB * TN; idx += ST_THREADS) {
        const int k = idx & 31, c = idx >> 5;
        if (k < jb && c < wt) Ap[k + (size_t)(c0 + c) * ld] = Bs[k * TNP + c];
    }
}

// -------------------------------------------------------------------------------------------
// Trailing update: one 128 x 64 tile of C per CTA.
// -------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 64, GEMM_THREADS = 256;

template <int W>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_kernel(Dims d, double **__restrict__ dA, int j, int tiles_m, int tiles_n, long batch,
            const int *__restrict__ index_list)
{
    __shared__ __align__(16) double As[W * GM];  // [k*GM + r]
    __shared__ __align__(16) double Bs[W * GN];  // [k*GN + c]

    const int tiles = tiles_m * tiles_n;
    const long slot = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;
    const int r0 = j + jb + (t % tiles_m) * GM;
    const int c0 = j + jb + (t / tiles_m) * GN;
    if (r0 >= m || c0 >= n) return;
    const int rows = (m - r0) < GM ? (m - r0) : GM;
    const int cols = (n - c0) < GN ? (n - c0) : GN;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *__restrict__ A = dA[b];
    const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

    // warp grid 4 x 2 over the tile; lane grid 8 x 4; thread tile 4 rows x 8 cols:
    // rows {2lr, 2lr+1, 16+2lr, 17+2lr}, cols {8q + 2lc, 8q + 2lc + 1 : q = 0..3} of the warp tile
    const int wr = wid & 3, wc = wid >> 2;
    const int lr = lane & 7, lc = lane >> 3;
    const int trow = wr * 32 + 2 * lr;
    const int tcol = wc * 32 + 2 * lc;

    // C first (longest latency), then the operands
    double acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = tcol + 8 * q + cc;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = trow + 16 * h;
                double v0 = 0.0, v1 = 0.0;
                if (c < cols && r < rows) {
                    const double *src = A + (size_t)(r0 + r) + (size_t)(c0 + c) * ld;
                    if (vec_ok && ((r0 + r) & 1) == 0 && r + 1 < rows) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(src);
                        v0 = t2.x;
                        v1 = t2.y;
                    } else {
                        v0 = src[0];
                        if (r + 1 < rows) v1 = src[1];
                    }
                }
                acc[2 * h][2 * q + cc] = v0;
                acc[2 * h + 1][2 * q + cc] = v1;
            }
        }
    // operands: asynchronous copies, all in flight together with the C loads above
    const double *__restrict__ L21 = A + (size_t)j * ld;  // column j, absolute rows
    const double *__restrict__ U12 = A + (size_t)j;       // row j, absolute columns
    if (vec_ok && (r0 & 1) == 0) {
#pragma unroll
        for (int idx = tid; idx < W * GM / 2; idx += GEMM_THREADS) {
            const int r = (idx & (GM / 2 - 1)) * 2, k = idx / (GM / 2);
            const double *src = &L21[(size_t)(r0 + r) + (size_t)k * ld];
            if (r + 1 < rows && k < jb) {
                cp_async16(&As[k * GM + r], src, true);
            } else {  // ragged tail: last valid row alone, everything else zero-filled
                const bool ok1 = (r < rows) && (k < jb);
                cp_async8(&As[k * GM + r], ok1 ? src : A, ok1);
                cp_async8(&As[k * GM + r + 1], A, false);
            }
        }
    } else {
#pragma unroll
        for (int idx = tid; idx < W * GM; idx += GEMM_THREADS) {
            const int r = idx & (GM - 1), k = idx / GM;
            const bool ok = (r < rows) && (k < jb);
            cp_async8(&As[idx], ok ? &L21[(size_t)(r0 + r) + (size_t)k * ld] : A, ok);
        }
    }
#pragma unroll
    for (int idx = tid; idx < W * GN; idx += GEMM_THREADS) {
        const int k = idx % W, c = idx / W;  // k fast: contiguous in global memory
        const bool ok = (c < cols) && (k < jb);
        cp_async8(&Bs[k * GN + c], ok ? &U12[(size_t)k + (size_t)(c0 + c) * ld] : A, ok);
    }
    cp_async_wait_all();
    __syncthreads();

#pragma unroll 4
    for (int k = 0; k < W; ++k) {
        if (k < jb) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&As[k * GM + trow]);
            const double2 a23 = *reinterpret_cast<const double2 *>(&As[k * GM + trow + 16]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 bq = *reinterpret_cast<const double2 *>(&Bs[k * GN + tcol + 8 * q]);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    acc[rr][2 * q] = fma(-av[rr], bq.x, acc[rr][2 * q]);
                    acc[rr][2 * q + 1] = fma(-av[rr], bq.y, acc[rr][2 * q + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = tcol + 8 * q + cc;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = trow + 16 * h;
                if (c < cols && r < rows) {
                    double *dst = A + (size_t)(r0 + r) + (size_t)(c0 + c) * ld;
                    if (vec_ok && ((r0 + r) & 1) == 0 && r + 1 < rows) {
                        *reinterpret_cast<double2 *>(dst) =
                            make_double2(acc[2 * h][2 * q + cc], acc[2 * h + 1][2 * q + cc]);
                    } else {
                        dst[0] = acc[2 * h][2 * q + cc];
                        if (r + 1 < rows) dst[1] = acc[2 * h + 1][2 * q + cc];
                    }
                }
            }
        }
}


// -------------------------------------------------------------------------------------------
// Trailing update on the FP64 tensor pipe: one 128 x 64 tile of C per CTA, C -= L21 * U12 with
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4). tools/dmma_probe.cu showed on the B200 that this
// instruction equals the chain c = fma(a0,b0,c); c = fma(a1,b1,c); ... with k increasing, bit for
// bit, so with A negated the tile receives exactly the canonical updates of oracle/lu_oracle.c.
// Relative to the DFMA kernel above: 16 DMMA per warp and k-step feed on 8 conflict-free LDS.64
// (the 4x8 DFMA register tile needs 6 LDS.128 per 32 DFMA), so neither issue slots nor the LSU pipe
// limit the FP64 pipe. Operands are staged once per tile with cp.async; zero padding beyond the
// panel width is exact (fma(-0, 0, c) = c).
// Warp grid 4 x 2, warp tile 32 x 32 = 4 x 4 fragments; lane (g = lane/4, q = lane%4) holds
// C(8i+g, 8jt+2q) and C(8i+g, 8jt+2q+1).
// -------------------------------------------------------------------------------------------
constexpr int DM = 128, DN = 64, DMMA_THREADS = 256;

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Region form: C[rs.., cs..climit) -= A[rs.., k0..k0+kw) * A[k0..k0+kw, cs..climit), kw = min(KW, min(m,n) - k0).
template <int KW>  // widest k range staged per tile (32 or 64)
__global__ void __launch_bounds__(DMMA_THREADS, 2)
gemm_dmma_kernel(Dims d, double **__restrict__ dA, int k0, int rs, int cs, int climit, int tiles_m, int tiles_n,
                 long batch, const int *__restrict__ index_list)
{
    constexpr int LDA = DM + 4;  // As[k*LDA + r]: (k, r) -> banks 8(k%4) + 2(r%8) + ..: conflict free
    constexpr int LDB = KW + 4;  // Bs[c*LDB + k]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);
    double *Bs = As + KW * LDA;

    const int tiles = tiles_m * tiles_n;
    const long slot = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    const int j = k0;
    if (j >= mn) return;
    const int jb = (mn - j) < KW ? (mn - j) : KW;
    const int r0 = rs + (t % tiles_m) * DM;
    const int c0 = cs + (t / tiles_m) * DN;
    const int nlim = n < climit ? n : climit;
    if (r0 >= m || c0 >= nlim) return;
    const int rows = (m - r0) < DM ? (m - r0) : DM;
    const int cols = (nlim - c0) < DN ? (nlim - c0) : DN;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wr = wid & 3, wc = wid >> 2;
    double *__restrict__ A = dA[b];
    const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

    // C fragments first (longest latency). Interior tiles take the unguarded path.
    double acc[4][4][2];
    const bool full = (rows == DM) && (cols == DN);
    double *Cb = A + (size_t)(r0 + wr * 32 + g) + (size_t)(c0 + wc * 32 + 2 * q) * ld;  // fragment (0,0)
    if (full) {
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) {
            const double *cp0 = Cb + (size_t)(8 * jt) * ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][jt][0] = cp0[8 * i];
                acc[i][jt][1] = cp0[8 * i + ld];
            }
        }
    } else {
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) {
            const int c = wc * 32 + 8 * jt + 2 * q;
            const double *cp0 = Cb + (size_t)(8 * jt) * ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool rok = (wr * 32 + 8 * i + g) < rows;
                acc[i][jt][0] = (rok && c < cols) ? cp0[8 * i] : 0.0;
                acc[i][jt][1] = (rok && c + 1 < cols) ? cp0[8 * i + ld] : 0.0;
            }
        }
    }
    // operands: L21 (rows x jb) -> As[k][r], U12 (jb x cols) -> Bs[c][k]; zero fill outside
    const double *__restrict__ L21 = A + (size_t)j * ld;  // column j, absolute rows
    const double *__restrict__ U12 = A + (size_t)j;       // row j, absolute columns
    if (vec_ok && (r0 & 1) == 0) {
        for (int idx = tid; idx < KW * DM / 2; idx += DMMA_THREADS) {
            const int r = (idx % (DM / 2)) * 2, k = idx / (DM / 2);
            const double *src = &L21[(size_t)(r0 + r) + (size_t)k * ld];
            if (r + 1 < rows && k < jb) {
                cp_async16(&As[k * LDA + r], src, true);
            } else {
                const bool ok1 = (r < rows) && (k < jb);
                cp_async8(&As[k * LDA + r], ok1 ? src : A, ok1);
                cp_async8(&As[k * LDA + r + 1], A, false);
            }
        }
    } else {
        for (int idx = tid; idx < KW * DM; idx += DMMA_THREADS) {
            const int r = idx % DM, k = idx / DM;
            const bool ok = (r < rows) && (k < jb);
            cp_async8(&As[k * LDA + r], ok ? &L21[(size_t)(r0 + r) + (size_t)k * ld] : A, ok);
        }
    }
    if (vec_ok && (j & 1) == 0) {
        for (int idx = tid; idx < KW * DN / 2; idx += DMMA_THREADS) {
            const int k = (idx % (KW / 2)) * 2, c = idx / (KW / 2);
            const double *src = &U12[(size_t)k + (size_t)(c0 + c) * ld];
            if (k + 1 < jb && c < cols) {
                cp_async16(&Bs[c * LDB + k], src, true);
            } else {
                const bool ok1 = (k < jb) && (c < cols);
                cp_async8(&Bs[c * LDB + k], ok1 ? src : A, ok1);
                cp_async8(&Bs[c * LDB + k + 1], A, false);
            }
        }
    } else {
        for (int idx = tid; idx < KW * DN; idx += DMMA_THREADS) {
            const int k = idx % KW, c = idx / KW;
            const bool ok = (k < jb) && (c < cols);
            cp_async8(&Bs[c * LDB + k], ok ? &U12[(size_t)k + (size_t)(c0 + c) * ld] : A, ok);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int ksteps = (jb + 3) / 4;
    const double *Ap = As + q * LDA + wr * 32 + g;          // + ks*4*LDA + 8*i
    const double *Bp = Bs + (wc * 32 + g) * LDB + q;        // + 8*jt*LDB + ks*4
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = -Ap[ks * 4 * LDA + 8 * i];
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) bf[jt] = Bp[8 * jt * LDB + ks * 4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jt = 0; jt < 4; ++jt) dmma_884(acc[i][jt][0], acc[i][jt][1], af[i], bf[jt]);
    }

    if (full) {
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) {
            double *cp0 = Cb + (size_t)(8 * jt) * ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cp0[8 * i] = acc[i][jt][0];
                cp0[8 * i + ld] = acc[i][jt][1];
            }
        }
    } else {
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) {
            const int c = wc * 32 + 8 * jt + 2 * q;
            double *cp0 = Cb + (size_t)(8 * jt) * ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool rok = (wr * 32 + 8 * i + g) < rows;
                if (rok && c < cols) cp0[8 * i] = acc[i][jt][0];
                if (rok && c + 1 < cols) cp0[8 * i + ld] = acc[i][jt][1];
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// Strip-resident update (panels at most 128 rows tall): one CTA per (matrix, 64-column strip).
// The whole strip -- block row and every row below it -- is brought into shared memory with
// coalesced full-column loads, the step's interchanges are applied there (no 8-byte scattered
// global accesses, no partial-sector traffic), the block row is solved against L11, the rows
// below get their rank-W update from shared memory, and the strip goes back with coalesced stores:
// HBM sees exactly one read and one write of the trailing matrix per step.
// -------------------------------------------------------------------------------------------
constexpr int SW = 64;            // strip width
constexpr int SROWS = 128;        // tallest strip
constexpr int LDC = SROWS + 2;    // padded column stride of the strip buffer (even: 16-byte rows pairs)
constexpr int STRIP_THREADS = 256;

struct StripSmem {
    double Cs[SW * LDC];        // strip, column-major: Cs[c*LDC + r], r = 0 is panel row j
    double As[KB * SROWS];      // L21, k-major: As[k*SROWS + r], r = 0 is row j + jb
    double Ls[KB * (KB + 1)];   // L11
    int top[KB], ddst[KB], dsrc[KB];
};

template <int W>
__global__ void __launch_bounds__(STRIP_THREADS, 2)
update_strip_kernel(Dims d, double **__restrict__ dA, const PivRec *__restrict__ recs, int j, int strips,
                    long batch, const int *__restrict__ index_list)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StripSmem &S = *reinterpret_cast<StripSmem *>(smem_raw);

    const long slot = blockIdx.x / strips;
    const int strip = blockIdx.x % strips;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;
    const int c0 = j + jb + strip * SW;
    if (c0 >= n) return;
    const int wt = (n - c0) < SW ? (n - c0) : SW;
    const int mp = m - j;        // strip rows (<= SROWS)
    const int mr = mp - jb;      // rows below the block row
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *__restrict__ A = dA[b];
    const PivRec &rec = recs[slot];
    if (tid < KB) {
        S.top[tid] = rec.top_src[tid];
        S.ddst[tid] = rec.down_dst[tid];
        S.dsrc[tid] = rec.down_src[tid];
    }
    const int nd = rec.n_down;

    // ---- loads: strip (full columns), L11, L21 ------------------------------------------------
    for (int c = wid; c < wt; c += STRIP_THREADS / 32) {
        const double *col = A + (size_t)j + (size_t)(c0 + c) * ld;
#pragma unroll
        for (int r = lane; r < SROWS; r += 32)
            if (r < mp) S.Cs[c * LDC + r] = col[r];
    }
    for (int idx = tid; idx < KB * KB; idx += STRIP_THREADS) {
        const int i = idx & 31, k = idx >> 5;
        if (i < jb && k < jb) S.Ls[i * (KB + 1) + k] = A[(size_t)(j + i) + (size_t)(j + k) * ld];
    }
#pragma unroll 4
    for (int idx = tid; idx < KB * SROWS; idx += STRIP_THREADS) {
        const int r = idx & (SROWS - 1), k = idx >> 7;
        S.As[idx] = (r < mr && k < jb) ? A[(size_t)(j + jb + r) + (size_t)(j + k) * ld] : 0.0;
    }
    __syncthreads();

    // ---- interchanges in shared memory: gather into registers, barrier, scatter ------------------
    {
        // items: (p, c) for the jb top positions, then (e, c) for the nd rows that go down
        const int items = (jb + nd) * SW;
        double v[(2 * KB * SW) / STRIP_THREADS];
#pragma unroll
        for (int u = 0; u < (2 * KB * SW) / STRIP_THREADS; ++u) {
            const int idx = tid + u * STRIP_THREADS;
            const int c = idx & (SW - 1), q = idx >> 6;
            v[u] = 0.0;
            if (idx < items && c < wt) {
                const int src = (q < jb) ? S.top[q] : S.dsrc[q - jb];
                v[u] = S.Cs[c * LDC + src];
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < (2 * KB * SW) / STRIP_THREADS; ++u) {
            const int idx = tid + u * STRIP_THREADS;
            const int c = idx & (SW - 1), q = idx >> 6;
            if (idx < items && c < wt) {
                const int dst = (q < jb) ? q : S.ddst[q - jb];
                S.Cs[c * LDC + dst] = v[u];
            }
        }
    }
    __syncthreads();

    // ---- U12 = L11^-1 * block row: one thread per column, canonical order -------------------------
    if (tid < wt) {
        double x[W];
#pragma unroll
        for (int i = 0; i < W; ++i) x[i] = (i < jb) ? S.Cs[tid * LDC + i] : 0.0;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (k < jb) {
#pragma unroll
                for (int i = 0; i < W; ++i)
                    if (i > k && i < jb) x[i] = fma(-S.Ls[i * (KB + 1) + k], x[k], x[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i)
            if (i < jb) S.Cs[tid * LDC + i] = x[i];
    }
    __syncthreads();

    // ---- rows below: C(r,c) = fma(-L21(r,k), U12(k,c), C(r,c)), k increasing ------------------------
    if (mr > 0) {
        const int wr = wid & 3, wc = wid >> 2;
        const int lr = lane & 7, lc = lane >> 3;
        const int trow = wr * 32 + 2 * lr;   // + {0,1,16,17}, relative to row j + jb
        const int tcol = wc * 32 + 2 * lc;   // + 8q + {0,1}
        if (wr * 32 < mr && wc * 32 < wt) {
            double acc[4][8];
            // jb may be odd on a matrix's last panel: C rows start at Cs row jb
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const double *colp = &S.Cs[(tcol + 8 * q + cc) * LDC + jb];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = trow + 16 * h;
                        acc[2 * h][2 * q + cc] = (r < mr) ? colp[r] : 0.0;
                        acc[2 * h + 1][2 * q + cc] = (r + 1 < mr) ? colp[r + 1] : 0.0;
                    }
                }
#pragma unroll 4
            for (int k = 0; k < W; ++k) {
                if (k < jb) {
                    const double2 a01 = *reinterpret_cast<const double2 *>(&S.As[k * SROWS + trow]);
                    const double2 a23 = *reinterpret_cast<const double2 *>(&S.As[k * SROWS + trow + 16]);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // U12(k, c) sits in the strip's block row: Cs[c*LDC + k]
                        const double bx = S.Cs[(tcol + 8 * q) * LDC + k];
                        const double by = S.Cs[(tcol + 8 * q + 1) * LDC + k];
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            acc[rr][2 * q] = fma(-av[rr], bx, acc[rr][2 * q]);
                            acc[rr][2 * q + 1] = fma(-av[rr], by, acc[rr][2 * q + 1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    double *colp = &S.Cs[(tcol + 8 * q + cc) * LDC + jb];
                    const bool cok = (tcol + 8 * q + cc) < wt;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = trow + 16 * h;
                        if (cok && r < mr) colp[r] = acc[2 * h][2 * q + cc];
                        if (cok && r + 1 < mr) colp[r + 1] = acc[2 * h + 1][2 * q + cc];
                    }
                }
        }
        __syncthreads();
    }

    // ---- store the strip ---------------------------------------------------------------------------
    for (int c = wid; c < wt; c += STRIP_THREADS / 32) {
        double *col = A + (size_t)j + (size_t)(c0 + c) * ld;
#pragma unroll
        for (int r = lane; r < SROWS; r += 32)
            if (r < mp) col[r] = S.Cs[c * LDC + r];
    }
}

// -------------------------------------------------------------------------------------------
// Deferred interchanges of the L part: column block J (32 columns) receives, in ONE pass at the
// end, every interchange of the later panels (LAPACK's dlaswp with k1 = 32(J+1)+1 .. min(m,n)).
// The reference applies them step by step with 8-byte scattered accesses
// (magmablas/zlaswp_batched.cu:31-43, "swap left" at src/zgetrf_batched.cpp:167-172); one pass
// reads and writes each column of L once, coalesced, permuting through shared memory.
// -------------------------------------------------------------------------------------------
constexpr int LSWP_THREADS = 256;
constexpr int LSWP_COLS = 4;  // columns staged per round

__global__ void __launch_bounds__(LSWP_THREADS)
laswp_left_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int blocks, int max_rows, int step,
                  int pair_end_block, long batch, const int *__restrict__ index_list)
{
    // blocks J < pair_end_block with J even were factored as the first half of a 64-wide pair: the second
    // half's interchanges were applied to them on the spot, the deferred pass starts one panel later
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *perm = reinterpret_cast<int *>(smem_raw);                                   // [max_rows]
    int *spiv = perm + max_rows;                                                     // [max_rows]
    double *buf = reinterpret_cast<double *>(spiv + max_rows + (max_rows & 1) * 0);  // [LSWP_COLS][max_rows]
    buf = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(buf) + 15) & ~(uintptr_t)15);

    const long slot = blockIdx.x / blocks;
    const int J = blockIdx.x % blocks;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    const bool paired = (J < pair_end_block) && ((J & 1) == 0);
    const int r0 = (J + (paired ? 2 : 1)) * step;  // first row touched
    if (r0 >= mn) return;                          // no later panel
    const int rows = m - r0;
    const int tid = threadIdx.x;
    double *__restrict__ A = dA[b];
    const int *__restrict__ ipiv = dipiv[b];
    for (int i = tid; i < rows; i += LSWP_THREADS) {
        perm[i] = i;
        spiv[i] = (r0 + i < mn) ? (ipiv[r0 + i] - 1 - r0) : i;
    }
    __syncthreads();
    if (tid == 0) {
        const int kmax = mn - r0;
        bool any = false;
        for (int i = 0; i < kmax; ++i) {
            const int p = spiv[i];
            if (p != i) {
                const int t = perm[i];
                perm[i] = perm[p];
                perm[p] = t;
                any = true;
            }
        }
        spiv[0] = any ? 1 : 0;
    }
    __syncthreads();
    if (!spiv[0]) return;
    const int cbeg = J * step, cend = (J + 1) * step;  // the block's own columns
    for (int cb = cbeg; cb < cend; cb += LSWP_COLS) {
        const int nc = (cend - cb) < LSWP_COLS ? (cend - cb) : LSWP_COLS;
        for (int c = 0; c < nc; ++c) {
            const double *col = A + (size_t)r0 + (size_t)(cb + c) * ld;
            for (int i = tid; i < rows; i += LSWP_THREADS) buf[c * max_rows + i] = col[i];
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            double *col = A + (size_t)r0 + (size_t)(cb + c) * ld;
            for (int i = tid; i < rows; i += LSWP_THREADS) {
                const int src = perm[i];
                if (src != i) col[i] = buf[c * max_rows + src];
            }
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------
// Left-looking column-slab update (matrices of at most 512 rows). One CTA owns the 32-column slab
// J of one matrix for the whole of its history: it reads the slab ONCE, through the composed row
// permutation of every earlier panel, keeps it in DMMA accumulator fragments, and for each earlier
// panel K (increasing) solves the 32 x 32 block row against L_KK and subtracts L(:, K) * U(K, J)
// from every row below -- the canonical k-increasing order per element, so the result is
// bit-identical to the right-looking flow. The slab goes back once, rows in current order, ready
// for its own panel factorisation.
// Why: the right-looking flow re-reads and re-writes the whole trailing matrix at every step and
// moves pivot rows with 8-byte scattered accesses (32-byte sectors, read-modify-write): 120 GB of
// HBM traffic at n = 512 x 4000 against 16.8 GB of matrix, i.e. the FP64 pipe waits on HBM. Here a
// slab costs one read + one write + one read of the L columns to its left (gathered by row, but the
// permutation is the identity except for <= 32 rows per panel): ~45 GB in total.
// L blocks stay in the row order of their own panel (the interchanges of later panels are applied
// to them in one pass at the end, laswp_left_kernel); rmap[K][i] walks row i back to that order
// through the per-panel step permutations the panel kernel records (sinv).
// Warp w owns the 8-row tiles w, w + NW, w + 2NW, w + 3NW (cyclic: the shrinking active region stays
// balanced); lane (g = lane/4, q = lane%4) holds C(8t+g, 8c+2q) and C(8t+g, 8c+2q+1).
// -------------------------------------------------------------------------------------------
constexpr int LL_LDU = 36;  // Us[k*LL_LDU + c]: B fragments (k = 4s+q, c = 8ct+g) hit 32 distinct banks

// mbarrier helpers of the L chunk pipeline
__device__ __forceinline__ void ll_mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ll_mbar_arrive(void *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void ll_mbar_arrive_cp_async(void *bar)  // arrives when this thread's earlier cp.asyncs have landed
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void ll_mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "LLWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@!p bra LLWAIT_%=;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void ll_mbar_expect_tx(void *bar, unsigned bytes)  // transaction bytes only, no arrival
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ll_bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar)  // 1-D TMA copy
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int NW>
struct LeftSmem {
    static constexpr int RING = (NW == 4) ? 3 : 4;  // chunks in the ring
    static constexpr int LDR = NW * 32 + 4;         // As[kk*LDR + row]: A fragments (k = 4s+q, row 8t+g) hit 32 distinct banks
    unsigned short rmap[NW][NW * 32];       // row i of the current order -> row in the order of panel K
    unsigned short src[NW * 32];            // row i of the current order -> original row (slab load)
    unsigned short sinv_s[NW][NW * 32];     // the recorded step permutations while the maps are built
    unsigned long long full[RING], empty[RING], lsbar, stbar;
    // ---- from here on one contiguous region: at kernel start it receives the whole slab (32 columns, stride LDR) by
    // TMA while the maps are being built; afterwards its parts take their own roles
    alignas(16) double Us[32 * LL_LDU];     // block row K of the slab as the owners left it
    double Un[32 * LL_LDU];                 // -U(K, J): written by the solve of step K (between the step's two
                                            // barriers), read by its ring pass; every warp is past that pass when
                                            // the next solve starts, so one buffer is enough
    alignas(16) double Ls[32 * 33];         // L_KK [k*32 + slot(i)]: read by the solve, refilled (cp.async) during the ring pass
    alignas(16) double ring[RING * 8 * LDR];  // L chunks, CTA-wide, rows in the order of their own panel
};

// scratch of the fused tail (factor_view's bookkeeping), overlaid on Us once the update loop is over
struct TailSmem {
    static constexpr int LDU = 36;
    double Un[8 * 36];
    double L11[64];
    int ipiv[128];
    int nmoves, info;
    unsigned char mdst[16], msrc[16], perm[128];
};

// ANY = false: the kernel as tuned for 16-byte aligned matrices (TMA staging only where every column starts on a 16-byte
// boundary and has an even number of rows, cp.async otherwise). ANY = true: TMA staging for any alignment through
// parity-shifted shared-memory columns (below); chosen by the host for variable-size batches and for fixed sizes with an odd
// m or ldda. Two instantiations because both paths in one kernel cost the aligned case 1.3-2.6% (n = 128, n = 512).
template <int NW, bool ANY>
__global__ void __launch_bounds__(NW * 32, NW == 16 ? 1 : (NW == 8 ? 2 : 4))
left_update_kernel(Dims d, double **__restrict__ dA, const unsigned short *__restrict__ sinv_g, int sinv_rows,
                   int sinv_blocks, int J, int finish, int ahead, int use_bulk, long batch, const int *__restrict__ index_list,
                   int tail, int **__restrict__ dipiv, int *__restrict__ dinfo, unsigned short *__restrict__ sinv_w)
{
    // tail = 1 (four-warp CTAs, at most 96 rows below the slab's block row): the slab's own panel is factored here, from
    // the accumulators, instead of going to HBM and back through a panel kernel (see the end of the kernel).
    // ahead: chunks in flight (1 .. RING-1); a ring slot is refilled RING - ahead chunks after its last use.
    // finish = 1: second visit of the slab that holds the LAST, narrower panel of a wide matrix
    // (32J < min(m,n) < min(n, 32J+32)): its columns right of the panel still need that panel's step.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeftSmem<NW> &S = *reinterpret_cast<LeftSmem<NW> *>(smem_raw);
    constexpr int T = NW * 32;
    constexpr int RING = LeftSmem<NW>::RING;
    static_assert(sizeof(unsigned short) * NW * T <= sizeof(double) * RING * 8 * LeftSmem<NW>::LDR, "sinv overlay");
    const unsigned FULLM = 0xffffffffu;

    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    const int c0 = 32 * J;
    if (c0 >= n) return;
    const int nc = (n - c0) < 32 ? (n - c0) : 32;
    if (finish && !(c0 < mn && mn < c0 + nc)) return;
    const int cb = finish ? mn - c0 : 0;                  // first column of the slab handled here
    const int kend = finish ? mn : (c0 < mn ? c0 : mn);   // pivots taken so far = rows of the slab that become U
    const int nk = (kend + 31) >> 5;                      // panels to the left ...
    const int kfirst = finish ? J : 0;                    // ... of which [kfirst, nk) are still to be applied
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(FULLM, tid >> 5, 0);
    const int g = lane >> 2, q = lane & 3;
    double *__restrict__ A = dA[b];

    constexpr int LDR = LeftSmem<NW>::LDR;
    static_assert(sizeof(double) * 32 * LDR <= sizeof(double) * (2 * 32 * LL_LDU + 32 * 33 + RING * 8 * LDR), "slab staging region");
    const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool aligned_ok = vec_ok && ((m & 1) == 0);  // 16-byte aligned columns of a multiple of 16 bytes (TMA stores)
    // TMA staging works for ANY alignment: a column whose first wanted element sits at an odd double index is copied from
    // one element earlier (always inside the matrix or, for column 0 of a matrix at 8 mod 16, inside its allocation, whose
    // start is 256-byte aligned) and lands one slot lower in shared memory, so that source and destination are both
    // 16-byte aligned; readers add the column's parity to the row index. An odd remainder row goes by a plain copy.
    // (Half of a vbatched batch with ldda = n has an odd leading dimension, half of the rest an 8-mod-16 base: BASELINE
    // config 4 ran three quarters of its matrices through the 8-byte cp.async paths, 58.3 ms against 49.7 ms all-aligned.)
    const bool bulk_ok = ANY ? ((use_bulk & 1) != 0) : (aligned_ok && use_bulk != 0);
    // parity of the double index of element (0, col); from A and ld each time, so that nothing extra stays live in the
    // update loop (the aligned case must keep the register allocation it had: C3 / C5 are all-aligned)
    auto colpar = [&](int col) { return (int)(((reinterpret_cast<uintptr_t>(A) >> 3) + (uintptr_t)(col & ld)) & 1u); };
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            ll_mbar_init(&S.full[i], T);
            ll_mbar_init(&S.empty[i], T);
        }
        ll_mbar_init(&S.lsbar, T);
        ll_mbar_init(&S.stbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- the slab: one TMA bulk copy per column into the staging region, in flight while the maps are built ------
    double *const stage = S.Us;
    if (!ANY && bulk_ok && tid < 32) {
        if (tid == 0) {
            int ncopy = 0;
            for (int cc = 0; cc < 32; ++cc) ncopy += (cc >= cb && cc < nc) ? 1 : 0;
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&S.stbar)),
                         "r"((unsigned)ncopy * (unsigned)m * 8u)
                         : "memory");
        }
        if (tid >= cb && tid < nc) ll_bulk_load(stage + tid * LDR, A + (size_t)(c0 + tid) * ld, (unsigned)m * 8u, &S.stbar);
    }
    if (ANY && bulk_ok && tid < 32) {
        // column tid of the slab: rows [-p, m) -> slots [0, m + p), p = its parity; an even number of them by TMA
        const bool mine = tid >= cb && tid < nc;
        const int p = colpar(c0 + tid);
        const unsigned ce = mine ? (unsigned)((m + p) & ~1) : 0u;
        const unsigned total = __reduce_add_sync(FULLM, ce);
        if (tid == 0)
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&S.stbar)),
                         "r"(total * 8u)
                         : "memory");
        if (mine) {
            const double *colp = A + (size_t)(c0 + tid) * ld;
            if (ce) ll_bulk_load(stage + tid * LDR, colp - p, ce * 8u, &S.stbar);
            if ((unsigned)(m + p) > ce) stage[tid * LDR + m - 1 + p] = colp[m - 1];  // visible after the barriers below
        }
    }
    // ---- row maps ----------------------------------------------------------------------------------------
    unsigned short(*sinv_s)[T] = S.sinv_s;
    {
        const unsigned short *sg = sinv_g + (size_t)slot * sinv_blocks * sinv_rows;
        for (int K = kfirst; K < nk; ++K) sinv_s[K][tid] = (tid < m) ? sg[(size_t)K * sinv_rows + tid] : (unsigned short)tid;
    }
    __syncthreads();
    {
        int r = tid;
        for (int K = nk - 1; K >= kfirst; --K) {
            S.rmap[K][tid] = (unsigned short)r;
            if (r >= 32 * K) r = sinv_s[K][r];
        }
        S.src[tid] = (unsigned short)r;
    }
    __syncthreads();  // maps complete

    // ---- L chunk pipeline: chunk 4K+ch = columns 32K+8ch .. +8 of L, every row below block row K, copied in the
    // row order of panel K itself: contiguous 16-byte cp.asyncs (the first version gathered each warp's rows through
    // rmap with 8-byte copies: 11 tag requests and 14 sectors per warp instruction instead of 4 and 16 for twice the
    // bytes, 63% of all LSU wavefronts of the kernel). The permutation is applied when the A fragments are READ from
    // shared memory (rmap, word granular). Slots are handed over with mbarriers: full[] counts the threads'
    // cp.async completions, empty[] the threads that are done reading.
    auto issue = [&](int nrel) {  // nrel: chunk number counted from the first step
        const int K = kfirst + (nrel >> 2), ch = nrel & 3;
        const int slot_r = nrel % RING;
        if (nrel >= RING) ll_mbar_wait(&S.empty[slot_r], (unsigned)((nrel / RING - 1) & 1));
        if (K < nk) {
            const int kbK = (kend - 32 * K) < 32 ? (kend - 32 * K) : 32;
            const int rlo = 32 * (K + 1);                     // even
            double *dst = S.ring + (size_t)slot_r * 8 * LDR;
            const double *colbase = A + (size_t)(32 * K + 8 * ch) * ld;
            if (rlo < m && bulk_ok && kbK == 32) {
                // TMA: one bulk copy per column of the chunk, issued by one thread; the slot's barrier takes the bytes
                // as a transaction count on top of the T plain arrivals below (LDGSTS needed up to four copies plus
                // address arithmetic from every thread: 20% of the kernel's stall samples sat in this block)
                // (one copy per LANE: issued from a loop on one thread, the copies of a step cost that warp ~2k cycles)
                // Column kk: rows [rlo - p, m) -> slots [rlo, m + p), p = the column's parity (rlo is even).
                if (tid < 8) {
                    if (!ANY) {  // every column starts on a 16-byte boundary and has an even number of rows
                        const unsigned bytes = (unsigned)(m - rlo) * 8u;
                        if (tid == 0) ll_mbar_expect_tx(&S.full[slot_r], 8u * bytes);
                        ll_bulk_load(dst + tid * LDR + rlo, colbase + (size_t)tid * ld + rlo, bytes, &S.full[slot_r]);
                    } else {
                        // column kk = tid: rows [rlo - p, m) -> slots [rlo, m + p), p = the column's parity (32 K + 8 ch is even)
                        const int p = colpar(tid);
                        const unsigned ce = (unsigned)((m - rlo + p) & ~1);
                        if (tid == 0) {
                            const unsigned c0e = (unsigned)((m - rlo + colpar(0)) & ~1), c1e = (unsigned)((m - rlo + colpar(1)) & ~1);
                            ll_mbar_expect_tx(&S.full[slot_r], 4u * (c0e + c1e) * 8u);  // the eight columns alternate (or share) a parity
                        }
                        const double *colp = colbase + (size_t)tid * ld;
                        if (ce) ll_bulk_load(dst + tid * LDR + rlo, colp + rlo - p, ce * 8u, &S.full[slot_r]);
                        if ((unsigned)(m - rlo + p) > ce) cp_async8(dst + tid * LDR + m - 1 + p, colp + m - 1, true);
                    }
                }
            } else if (rlo < m) {
                if (vec_ok) {
                    const int npairs = (m - rlo + 1) >> 1;
                    const unsigned magic = 0xFFFFFFFFu / (unsigned)npairs + 1u;  // u / npairs == umulhi(u, magic) for npairs > 1 (u * npairs < 2^32)
                    // at most 8 * (NW*16) / T = 4 copies per thread: unrolled, so that every copy has its own address
                    // registers (LDGSTS holds them until it has read them: a rolled loop stalled on that hazard)
                    const int total = 8 * npairs;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int u = tid + it * T;
                        if (u < total) {
                            const int kk = npairs > 1 ? (int)__umulhi((unsigned)u, magic) : u, r = rlo + 2 * (u - kk * npairs);
                            const bool kok = (8 * ch + kk) < kbK;
                            const int bytes = kok ? ((r + 1 < m) ? 16 : 8) : 0;
                            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + kk * LDR + r);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa),
                                         "l"(kok ? colbase + (size_t)kk * ld + r : A), "r"(bytes)
                                         : "memory");
                        }
                    }
                } else {
                    const int nr = m - rlo;
                    const unsigned magic = 0xFFFFFFFFu / (unsigned)nr + 1u;
                    for (int u = tid; u < 8 * nr; u += T) {
                        const int kk = nr > 1 ? (int)__umulhi((unsigned)u, magic) : u, r = rlo + (u - kk * nr);
                        const bool kok = (8 * ch + kk) < kbK;
                        cp_async8(dst + kk * LDR + r, kok ? colbase + (size_t)kk * ld + r : A, kok);
                    }
                }
            }
        }
        ll_mbar_arrive_cp_async(&S.full[slot_r]);
    };
    int ls_uses = 0;  // completed uses of lsbar
    auto stage_lkk = [&](int K) {  // L_KK -> Ls, zero outside the panel width
        const int kbK = (kend - 32 * K) < 32 ? (kend - 32 * K) : 32;
        const double *LKK = A + (size_t)(32 * K) + (size_t)(32 * K) * ld;
        if (false && bulk_ok && (use_bulk & 2) && kbK == 32) {  // TMA variant (plain column layout): measured slower, retired
            if (tid < 32) {
                if (tid == 0) ll_mbar_expect_tx(&S.lsbar, 32u * 256u);
                ll_bulk_load(&S.Ls[tid * 32], LKK + (size_t)tid * ld, 256u, &S.lsbar);
            }
        } else {
            for (int idx = tid; idx < 1024; idx += T) {
                const int i = idx & 31, k = idx >> 5;
                const bool ok = (i < kbK && k < kbK);
                // NW = 4: per column k the 32 multipliers sit in the order the solving lanes read them (a lane's rows
                // rg + 4 i8 at rg*8 + i8: LDS.128). Wider CTAs keep the plain order (the same change cost them 1.5%).
                const int slot_i = (NW == 4) ? ((i & 3) * 8 + (i >> 2)) : i;
                cp_async8(&S.Ls[k * 32 + slot_i], ok ? LKK + i + (size_t)k * ld : A, ok);
            }
        }
        ll_mbar_arrive_cp_async(&S.lsbar);
    };

    // ---- the slab, rows in current order ---------------------------------------------------------------------
    // Copied with contiguous 16-byte cp.asyncs (original row order) into the ring, which the L chunks do not need
    // yet, and picked up through src[] from shared memory: the direct gather cost 26 tag requests per 8-byte warp load
    // and a fifth of the kernel's time on the middle slabs.
    constexpr int SCOLS = (RING * 8 >= 32) ? 32 : 16;  // slab columns the ring can hold at once (LDGSTS path)
    double acc[4][4][2];
    if (bulk_ok) {
        ll_mbar_wait(&S.stbar, 0u);  // all 32 columns have landed
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int row = 8 * (w + NW * a) + g;
            const bool rok = row < m;
            const double *sp = stage + (rok ? (int)S.src[row] : 0) + (2 * q) * LDR;
            const int p0 = ANY ? colpar(0) : 0, p1 = ANY ? colpar(1) : 0;  // even / odd columns of the slab (c0 is even)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[a][c][0] = rok ? sp[(8 * c) * LDR + p0] : 0.0;
                acc[a][c][1] = rok ? sp[(8 * c + 1) * LDR + p1] : 0.0;
            }
        }
        __syncthreads();  // the staging region is free: its parts take their own roles
    } else {
#pragma unroll
    for (int p0 = 0; p0 < 32; p0 += SCOLS) {
        if (vec_ok) {
            const int npairs = (m + 1) >> 1;
            const unsigned magic = 0xFFFFFFFFu / (unsigned)npairs + 1u;
            for (int u = tid; u < SCOLS * npairs; u += T) {
                const int cc = npairs > 1 ? (int)__umulhi((unsigned)u, magic) : u, r = 2 * (u - cc * npairs);
                const int col = p0 + cc;
                const bool ok = (col >= cb && col < nc);
                const int bytes = ok ? ((r + 1 < m) ? 16 : 8) : 0;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(S.ring + cc * LDR + r);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa),
                             "l"(ok ? A + r + (size_t)(c0 + col) * ld : A), "r"(bytes)
                             : "memory");
            }
        } else {
            const unsigned magic = 0xFFFFFFFFu / (unsigned)m + 1u;
            for (int u = tid; u < SCOLS * m; u += T) {
                const int cc = m > 1 ? (int)__umulhi((unsigned)u, magic) : u, r = u - cc * m;
                const int col = p0 + cc;
                const bool ok = (col >= cb && col < nc);
                cp_async8(S.ring + cc * LDR + r, ok ? A + r + (size_t)(c0 + col) * ld : A, ok);
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int row = 8 * (w + NW * a) + g;
            const bool rok = row < m;
            const double *sp = S.ring + (rok ? (int)S.src[row] : 0) + (2 * q - p0) * LDR;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (8 * c >= p0 && 8 * c < p0 + SCOLS) {  // static
                    acc[a][c][0] = rok ? sp[(8 * c) * LDR] : 0.0;
                    acc[a][c][1] = rok ? sp[(8 * c + 1) * LDR] : 0.0;
                }
            }
        }
        __syncthreads();  // the ring is free again; every row of these columns is in registers
    }
    }
    stage_lkk(kfirst);
#pragma unroll 1
    for (int pch = 0; pch < ahead; ++pch) issue(pch);

    int nrel = 0;  // chunk consumed next, counted from the first step
#pragma unroll 1
    for (int K = kfirst; K < nk; ++K) {
        const int kb = (kend - 32 * K) < 32 ? (kend - 32 * K) : 32;
        // ---- block row K -> Us -----------------------------------------------------------------------------
        {
            const int aK = (4 * K) / NW;            // accumulator slot of the block row's tiles
            const int tt = w - (4 * K) % NW;        // which of its four tiles this warp holds, if any
            if (tt >= 0 && tt < 4) {
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a == aK) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<double2 *>(&S.Us[(8 * tt + g) * LL_LDU + 8 * c + 2 * q]) =
                                make_double2(acc[a][c][0], acc[a][c][1]);
                    }
                }
            }
        }
        __syncthreads();
        // ---- U(K, J) = L_KK^-1 * block row ----------------------------------------------------------------------
        if (NW < 8 || w < 8) ll_mbar_wait(&S.lsbar, (unsigned)(ls_uses & 1));  // L_KK has landed
        ++ls_uses;
        if (NW >= 8) {
            // eight warps: column 4w+q, rows g, g+8, g+16, g+24 per lane
            if (w < 8) {
                const int cc = 4 * w + q;
                const double *Lk = S.Ls;
                double x[4];
#pragma unroll
                for (int sblk = 0; sblk < 4; ++sblk) x[sblk] = S.Us[(g + 8 * sblk) * LL_LDU + cc];
#pragma unroll
                for (int k = 0; k < 31; ++k) {
                    const double u = __shfl_sync(FULLM, x[k >> 3], ((k & 7) << 2) | q);
#pragma unroll
                    for (int sblk = 0; sblk < 4; ++sblk) {
                        if (8 * sblk + 7 > k) {  // static: this row group still has rows below k
                            const int i = g + 8 * sblk;
                            const double l = Lk[k * 32 + i];
                            if (i > k) x[sblk] = fma(-l, u, x[sblk]);
                        }
                    }
                }
                const bool cok = (cc >= cb && cc < nc);
                double *Un = S.Un;
#pragma unroll
                for (int sblk = 0; sblk < 4; ++sblk) {
                    const int i = g + 8 * sblk;
                    Un[i * LL_LDU + cc] = -x[sblk];
                    if (cok && i < kb) A[(size_t)(32 * K + i) + (size_t)(c0 + cc) * ld] = x[sblk];  // final rows of U
                }
            }
        } else {
            // four warps (at most 128 rows): column 8w + lane%8, rows lane/8 + 4i (i < 8) per lane
            const int cl = lane & 7, rg = lane >> 3;
            const int cc = 8 * w + cl;
            const unsigned lk = (unsigned)__cvta_generic_to_shared(S.Ls) + (unsigned)(rg * 8) * 8u;
            double x[8];
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) x[i8] = S.Us[(rg + 4 * i8) * LL_LDU + cc];
            // the multipliers of step k + 1 are fetched (LDS.128 into their own registers) before the DFMAs of step k:
            // written as `l = Ls[..]` per use, ptxas recycled ONE register pair for every load, so each DFMA waited out
            // a full shared-memory latency (ncu: 8 exposed LDS per step). n = 128: 10.47 -> 9.87 ms on the same box.
            double lb[2][8];
            auto load_l = [&](double (&dst)[8], int k) {  // only the pairs that still have rows below k (static)
#pragma unroll
                for (int p2 = 0; p2 < 4; ++p2)
                    if (4 * (2 * p2 + 1) + 3 > k)
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(dst[2 * p2]), "=d"(dst[2 * p2 + 1])
                                     : "r"(lk + (unsigned)(k * 32 + 2 * p2) * 8u) : "memory");
            };
            load_l(lb[0], 0);
#pragma unroll
            for (int k = 0; k < 31; ++k) {
                if (k + 1 < 31) load_l(lb[(k + 1) & 1], k + 1);
                const double u = __shfl_sync(FULLM, x[k >> 2], ((k & 3) << 3) | cl);
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    if (4 * i8 + 3 > k) {
                        const int i = rg + 4 * i8;
                        if (i > k) x[i8] = fma(-lb[k & 1][i8], u, x[i8]);
                    }
                }
            }
            const bool cok = (cc >= cb && cc < nc);
            double *Un = S.Un;
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
                const int i = rg + 4 * i8;
                Un[i * LL_LDU + cc] = -x[i8];
                if (cok && i < kb) A[(size_t)(32 * K + i) + (size_t)(c0 + cc) * ld] = x[i8];
            }
        }
        __syncthreads();
        // ---- rows below: C -= L(:, K) * U(K, J) on the tensor pipe, L streamed through the ring ---------------
        // A warp's tiles leave the active region in order (a = 0 first): the pass is specialised on the number
        // of tiles already out. (With a per-tile `if` ptxas predicates the DMMAs instead of branching, and a
        // predicated-off DMMA still occupies the tensor pipe: half of its cycles in the late slabs.)
        const double *Un = S.Un;
        // this lane's four tile rows in the row order of panel K (where the chunk's rows sit in shared memory)
        int roff[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int row = 8 * (w + NW * a) + g;
            roff[a] = (row < m) ? (int)S.rmap[K][row] : 32 * (K + 1);  // rows past m: any row of the chunk
        }
        // chunks that came in by TMA sit one slot lower in the columns of odd parity; this lane reads columns 4 s2 + q
        if (ANY && bulk_ok && kb == 32) {
            const int pq = colpar(q);
#pragma unroll
            for (int a = 0; a < 4; ++a) roff[a] += pq;
        }
        auto ring_pass = [&](auto amin_c, auto aend_c) {
            constexpr int AMIN = decltype(amin_c)::value;   // first active tile slot
            constexpr int AEND = decltype(aend_c)::value;   // one past the last tile slot that has rows (< m)
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                issue(nrel + ahead);
                if (ch == 0 && K + 1 < nk) stage_lkk(K + 1);
                const int slot_r = nrel % RING;
                ll_mbar_wait(&S.full[slot_r], (unsigned)((nrel / RING) & 1));
                if (AMIN < AEND) {
                    const double *As = S.ring + (size_t)slot_r * 8 * LDR + q * LDR;
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        double bf[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) bf[c] = Un[(8 * ch + 4 * s2 + q) * LL_LDU + 8 * c + g];
#pragma unroll
                        for (int a = AMIN; a < AEND; ++a) {
                            const double af = As[4 * s2 * LDR + roff[a]];
#pragma unroll
                            for (int c = 0; c < 4; ++c) dmma_884(acc[a][c][0], acc[a][c][1], af, bf[c]);
                        }
                    }
                }
                ll_mbar_arrive(&S.empty[slot_r]);  // after the DMMAs that consumed the fragments
                ++nrel;
            }
        };
        {
            // first tile slot a with w + NW a > 4K + 3
            const int need = 4 * K + 4 - w;  // tiles t >= 4K+4  <=>  NW a >= need
            const int amin = need <= 0 ? 0 : (need + NW - 1) / NW;
            // tile slots of this warp that hold rows of the matrix: a < aend (a matrix shorter than the CTA's reach,
            // e.g. 384 rows under 16 warps or any vbatched class member, must not issue DMMAs for the missing tiles:
            // predicated off they would still occupy the tensor pipe)
            const int tiles_m = (m + 7) >> 3;
            const int aend = tiles_m <= w ? 0 : ((tiles_m - w + NW - 1) / NW > 4 ? 4 : (tiles_m - w + NW - 1) / NW);
            const int a0 = amin < aend ? amin : aend;  // nothing to do: (aend, aend)
#define MB200_RP(A0, A1) ring_pass(std::integral_constant<int, A0>{}, std::integral_constant<int, A1>{})
            switch (a0 * 5 + aend) {
                case 0 * 5 + 4: MB200_RP(0, 4); break;
                case 1 * 5 + 4: MB200_RP(1, 4); break;
                case 2 * 5 + 4: MB200_RP(2, 4); break;
                case 3 * 5 + 4: MB200_RP(3, 4); break;
                case 0 * 5 + 3: MB200_RP(0, 3); break;
                case 1 * 5 + 3: MB200_RP(1, 3); break;
                case 2 * 5 + 3: MB200_RP(2, 3); break;
                case 0 * 5 + 2: MB200_RP(0, 2); break;
                case 1 * 5 + 2: MB200_RP(1, 2); break;
                case 0 * 5 + 1: MB200_RP(0, 1); break;
                default: MB200_RP(0, 0); break;  // no tile: keep the pipeline protocol running
            }
#undef MB200_RP
        }
    }

    // ---- fused tail: this slab's panel, factored where it already is -------------------------------------------------
    // The updated rows below the block row ARE the next panel. Instead of storing them and launching a panel kernel that
    // loads them again, the CTA writes them into a shared-memory image (the ring is free now) and runs the resident-block
    // factorisation of lu_chain.cuh on it: warp 0 = single-warp pivot chains, warps 1..3 = permutation, block-row solve and
    // rank-8 DMMA updates. Outputs as panel_chain_kernel's (factors in final row order, pivots, step permutation record,
    // info). Three launches and three slab round trips fewer at n = 128 -- and MEASURED SLOWER (same box, 50000 x 128^2:
    // 8.99 ms off, 9.69 ms for the 33..96-row panels, 10.28 ms with the last panel too; n = 96: 4.98 / 5.39 / 5.98): this
    // kernel keeps four CTAs, i.e. four pivot chains, per SM where the panel kernels keep 8..12, and chain throughput is
    // chains in flight / ~850 cycles per column; the overlap with the other CTAs' DMMA phases does not make up for it.
    // Kept behind magma_b200_set_fused_tail (default 0) with its parity tests.
    if (NW == 4 && tail && !finish && c0 < mn) {
        constexpr int LDV = 98;  // at most 96 rows (the driver's condition)
        static_assert(sizeof(TailSmem) <= sizeof(double) * 32 * LL_LDU, "tail scratch fits Us");
        static_assert(NW != 4 || 32 * LDV <= RING * 8 * LDR, "panel image fits the ring");
        __syncthreads();  // every warp is past its last ring pass and solve: the staging region is free
        TailSmem &P = *reinterpret_cast<TailSmem *>(S.Us);
        double *V = S.ring;
        const int jb = (mn - c0) < 32 ? (mn - c0) : 32;
        const int mp = m - c0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int row = 8 * (w + NW * a) + g;
            if (row >= c0 && row < m) {
                double *vp = V + (2 * q) * LDV + (row - c0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = 8 * c + 2 * q;
                    if (col < jb) vp[(8 * c) * LDV] = acc[a][c][0];
                    if (col + 1 < jb) vp[(8 * c + 1) * LDV] = acc[a][c][1];
                }
                if (jb < nc) {  // wide last panel: the columns right of it go back as they are (the finish pass takes them)
                    double *dst = A + row + (size_t)(c0 + 2 * q) * ld;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = 8 * c + 2 * q;
                        if (col >= jb && col < nc) dst[(size_t)(8 * c) * ld] = acc[a][c][0];
                        if (col + 1 >= jb && col + 1 < nc) dst[(size_t)(8 * c + 1) * ld] = acc[a][c][1];
                    }
                }
            }
        }
        if (tid == 0) {
            P.info = 0;
            P.nmoves = 0;
        }
        P.perm[tid] = (unsigned char)tid;  // T = 128
        __syncthreads();
        factor_view<LDV, true, 3>(P, V, mp, jb, 0, nullptr, 0, tid, lane, w);
        __syncthreads();
        double *Ap = A + c0 + (size_t)c0 * ld;
        if (aligned_ok) {
            f_fence_async_smem();
            __syncthreads();
            if (w == 0 && lane < jb) f_bulk_store(Ap + (size_t)lane * ld, &V[lane * LDV], (unsigned)mp * 8u);
        } else {
            for (int c = w; c < jb; c += NW)
                for (int r = lane; r < mp; r += 32) Ap[r + (size_t)c * ld] = V[c * LDV + r];
        }
        if (tid < jb) dipiv[b][c0 + tid] = c0 + P.ipiv[tid] + 1;
        {
            unsigned short *sv = sinv_w + ((size_t)slot * sinv_blocks + J) * sinv_rows + c0;
            for (int p = tid; p < mp; p += T) sv[p] = (unsigned short)(c0 + P.perm[p]);
        }
        if (tid == 0 && P.info && dinfo[b] == 0) dinfo[b] = c0 + P.info;  // J > 0: earlier panels have priority
        if (aligned_ok && w == 0) f_bulk_commit_wait();
        return;
    }
    // ---- rows that are not U yet go back, current order ---------------------------------------------------------
    const int cfirst = cb;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int row = 8 * (w + NW * a) + g;
        if (row >= kend && row < m) {
            double *dst = A + row + (size_t)(c0 + 2 * q) * ld;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = 8 * c + 2 * q;
                if (col >= cfirst && col < nc) dst[(size_t)(8 * c) * ld] = acc[a][c][0];
                if (col + 1 >= cfirst && col + 1 < nc) dst[(size_t)(8 * c + 1) * ld] = acc[a][c][1];
            }
        }
    }
}

template <int NW>
magma_int_t launch_left_update(const Dims &d, double **dA, unsigned short *sinv, int sinv_rows, int sinv_blocks,
                               int J, int finish, long batch, const int *il, cudaStream_t s, int tail = 0, int **dipiv = nullptr,
                               int *dinfo = nullptr)
{
    const size_t smem = sizeof(LeftSmem<NW>);
    static DevOnce once, once_any;
    smem_optin(once, left_update_kernel<NW, false>, smem);
    smem_optin(once_any, left_update_kernel<NW, true>, smem);
    // variable sizes, or a fixed size with an odd m or ldda: the any-alignment instantiation (see the kernel)
    const bool any = d.vm != nullptr || (d.m & 1) || (d.ldda & 1);
    static int ahead = -1;
    if (ahead < 0) {
        const char *e = getenv("MB200_LL_AHEAD");  // tuning sweeps
        // measured with the TMA chunks (ahead = 1 / 2 / 3): n = 512 25.58 / 24.76 / 25.01 ms, n = 256 16.27 / 16.12 / 15.92
        ahead = e ? atoi(e) : (NW == 8 ? 3 : 2);
        if (ahead < 1) ahead = 1;
        if (ahead > LeftSmem<NW>::RING - 1) ahead = LeftSmem<NW>::RING - 1;
    }
    static int use_bulk = -1;
    if (use_bulk < 0) {
        const char *e = getenv("MB200_LL_BULK");  // 0: LDGSTS staging instead of TMA bulk copies (A/B runs)
        use_bulk = e ? atoi(e) : 1;  // bit 0: L chunks and slab (n = 512: 30.8 -> 27.2 ms), bit 1: L_KK too (32 copies of 256 B: 28.7 ms, off)
    }
    if (any)
        left_update_kernel<NW, true><<<(unsigned)batch, NW * 32, smem, s>>>(d, dA, sinv, sinv_rows, sinv_blocks, J, finish, ahead, use_bulk,
                                                                            batch, il, NW == 4 ? tail : 0, dipiv, dinfo, sinv);
    else
        left_update_kernel<NW, false><<<(unsigned)batch, NW * 32, smem, s>>>(d, dA, sinv, sinv_rows, sinv_blocks, J, finish, ahead, use_bulk,
                                                                             batch, il, NW == 4 ? tail : 0, dipiv, dinfo, sinv);
    count_launch();
    MB200_CHECK_LAUNCH("left_update_kernel");
    return 0;
}

// -------------------------------------------------------------------------------------------
// Deferred interchanges of the L part when the step permutations are on record (left-looking driver):
// one CTA per (matrix, column block J), thread = row, 8 columns at a time. The source row of every
// final position comes from walking the recorded permutations back (nk - J - 1 table look-ups in
// shared memory, all rows in parallel) instead of replaying up to 480 interchanges on one thread;
// each thread holds its 8 values in registers across ONE barrier (all reads of a column precede all
// writes), rows that did not move are neither read nor written.
// -------------------------------------------------------------------------------------------
constexpr int LSP_COLS = 8;

__global__ void __launch_bounds__(512)
laswp_left_sinv_kernel(Dims d, double **__restrict__ dA, const unsigned short *__restrict__ sinv_g, int sinv_rows,
                       int sinv_blocks, int blocks, long batch, const int *__restrict__ index_list)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *tab = reinterpret_cast<unsigned short *>(smem_raw);  // [later panels][sinv_rows]

    constexpr int GROUPS = 32 / LSP_COLS;
    const int J = blockIdx.x % blocks;
    const long slot = blockIdx.x / blocks;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    const int nk = (mn + 31) >> 5;  // panels of this matrix
    if (J + 1 >= nk) return;        // no later panel: nothing to apply
    const int r0 = 32 * (J + 1);
    const int rows = m - r0;
    const int tid = threadIdx.x;
    const unsigned short *sg = sinv_g + (size_t)slot * sinv_blocks * sinv_rows;
    for (int K = J + 1; K < nk; ++K)
        for (int i = tid; i < m; i += blockDim.x) tab[(K - J - 1) * sinv_rows + i] = sg[(size_t)K * sinv_rows + i];
    __syncthreads();
    int r = r0 + tid;
    const bool live = tid < rows;
    if (live) {
        for (int K = nk - 1; K > J; --K)
            if (r >= 32 * K) r = tab[(K - J - 1) * sinv_rows + r];
    }
    const bool moved = live && (r != r0 + tid);
    // the block's 32 columns in groups of 8 (one table and one walk per block; groups touch disjoint columns, so
    // one barrier per group -- all reads of its columns before any write -- is enough)
#pragma unroll 1
    for (int grp = 0; grp < GROUPS; ++grp) {
        double *__restrict__ A = dA[b] + (size_t)(32 * J + LSP_COLS * grp) * ld;
        double v[LSP_COLS];
#pragma unroll
        for (int c = 0; c < LSP_COLS; ++c) v[c] = moved ? A[r + (size_t)c * ld] : 0.0;
        __syncthreads();
        if (moved) {
#pragma unroll
            for (int c = 0; c < LSP_COLS; ++c) A[r0 + tid + (size_t)c * ld] = v[c];
        }
    }
}

// C[rs.., cs..climit) -= A[:, k0..k0+KW) * A[k0..k0+KW, :) over at most rows_max x cols_max per matrix
template <int KW>
magma_int_t launch_gemm_dmma(const Dims &d, double **dA, int k0, int rs, int cs, int climit, int rows_max, int cols_max,
                             long batch, const int *il, cudaStream_t s)
{
    if (rows_max <= 0 || cols_max <= 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)KW * (DM + 4) + (size_t)DN * (KW + 4));
    static DevOnce once;
    smem_optin(once, gemm_dmma_kernel<KW>, smem);
    const int tiles_m = (rows_max + DM - 1) / DM, tiles_n = (cols_max + DN - 1) / DN;
    const long grid = (long)tiles_m * tiles_n * batch;
    if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
    gemm_dmma_kernel<KW><<<(unsigned)grid, DMMA_THREADS, smem, s>>>(d, dA, k0, rs, cs, climit, tiles_m, tiles_n, batch, il);
    count_launch();
    MB200_CHECK_LAUNCH("gemm_dmma_kernel");
    return 0;
}

// 32-column register panel with T threads (one row each). Short panels get a register budget (80) that lets two
// (<= 384 rows), three (<= 256), four (<= 192), six (<= 128) ... CTAs share an SM: the per-column pivot chain (~900 cycles) is latency, and
// an SM fills it only with other panels.
inline void launch_panel32(const Dims &d, double **dA, int **dipiv, int *dinfo, PivRec *recs, int j, int T, long batch,
                           const int *il, cudaStream_t s, unsigned short *sinv = nullptr, int sinv_rows = 0,
                           int sinv_blocks = 0, int nopiv = 0)
{
    if (T <= 64)
        panel_kernel<1, 32, 64, 12><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else if (T <= 96)
        panel_kernel<1, 32, 96, 8><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else if (T <= 128)
        panel_kernel<1, 32, 128, 6><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else if (T <= 192)
        panel_kernel<1, 32, 192, 4><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else if (T <= 256)
        panel_kernel<1, 32, 256, 3><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else if (T <= 384)
        panel_kernel<1, 32, 384, 2><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    else
        panel_kernel<1, 32><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il, sinv, sinv_rows, sinv_blocks, nopiv);
    count_launch();
}

// Tall panels (129..512 rows) of the left-looking driver: two 16-column halves per thread (panel2_kernel)
inline void launch_panel2(const Dims &d, double **dA, int **dipiv, int *dinfo, int j, int T, long batch, const int *il,
                          cudaStream_t s, unsigned short *sinv, int sinv_rows, int sinv_blocks)
{
    const size_t smem = (size_t)T * 16 * sizeof(double);  // the parked half
    static DevOnce o192, o256, o384, o512;
    if (T <= 192) {
        smem_optin(o192, panel2_kernel<192, 6>, 192 * 16 * sizeof(double));
        panel2_kernel<192, 6><<<(unsigned)batch, T, smem, s>>>(d, dA, dipiv, dinfo, j, batch, il, sinv, sinv_rows, sinv_blocks);
    } else if (T <= 256) {
        smem_optin(o256, panel2_kernel<256, 4>, 256 * 16 * sizeof(double));
        panel2_kernel<256, 4><<<(unsigned)batch, T, smem, s>>>(d, dA, dipiv, dinfo, j, batch, il, sinv, sinv_rows, sinv_blocks);
    } else if (T <= 384) {
        smem_optin(o384, panel2_kernel<384, 3>, 384 * 16 * sizeof(double));
        panel2_kernel<384, 3><<<(unsigned)batch, T, smem, s>>>(d, dA, dipiv, dinfo, j, batch, il, sinv, sinv_rows, sinv_blocks);
    } else {
        smem_optin(o512, panel2_kernel<512, 2>, 512 * 16 * sizeof(double));
        panel2_kernel<512, 2><<<(unsigned)batch, T, smem, s>>>(d, dA, dipiv, dinfo, j, batch, il, sinv, sinv_rows, sinv_blocks);
    }
    count_launch();
}

// Two 32-column panels as one 64-wide step (panels <= 512 rows, both in the tiled regime). The trailing
// matrix is read and written ONCE per 64 columns instead of once per 32 (the k = 32 update was HBM bound:
// 87 GB per n = 512 call):
//   panel(j); interchanges + U12a for every column right of it; k = 32 update of the NEXT panel's columns only;
//   panel(j+32); its interchanges applied to the first panel's columns on the spot (so L21 of both panels is in
//   the same row order); interchanges + postponed k = j..j+31 update of the block row + U12b for the columns
//   right of j+64; one k = 64 update of everything below and right of j+64.
// Per element the updates still arrive with k increasing: bit-identical to the 32-wide flow.
magma_int_t run_pair(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, PivRec *recs, int j,
                     long batch, const int *il, cudaStream_t s)
{
    const int j2 = j + 32;
    auto panel = [&](int jj) -> magma_int_t {
        int T = ((max_m - jj + 31) / 32) * 32;
        if (T > 512) return MAGMA_ERR_NOT_SUPPORTED;
        launch_panel32(d, dA, dipiv, dinfo, recs, jj, T, batch, il, s);
        MB200_CHECK_LAUNCH("panel_kernel");
        return 0;
    };
    auto swap_trsm = [&](int jj, int right_tiles, int left_tiles, int left_begin, int pre_k0) -> magma_int_t {
        const long grid = (long)(right_tiles + left_tiles) * batch;
        if (grid <= 0) return 0;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        swap_trsm_kernel<32><<<(unsigned)grid, ST_THREADS, 0, s>>>(d, dA, recs, jj, right_tiles, left_tiles, left_begin, pre_k0,
                                                                   batch, il);
        count_launch();
        MB200_CHECK_LAUNCH("swap_trsm_kernel");
        return 0;
    };
    magma_int_t rc;
    const int nright1 = d.vm ? (max_n - j - 1) : (max_n - j - 32);
    const int nright2 = d.vm ? (max_n - j2 - 1) : (max_n - j2 - 32);
    if ((rc = panel(j)) != 0) return rc;
    if ((rc = swap_trsm(j, nright1 > 0 ? (nright1 + TN - 1) / TN : 0, 0, 0, -1)) != 0) return rc;
    if ((rc = launch_gemm_dmma<32>(d, dA, j, j2, j2, j2 + 32, max_m - j2, 32, batch, il, s)) != 0) return rc;
    if ((rc = panel(j2)) != 0) return rc;
    if ((rc = swap_trsm(j2, 0, 1, j, -1)) != 0) return rc;
    if ((rc = swap_trsm(j2, nright2 > 0 ? (nright2 + TN - 1) / TN : 0, 0, 0, j)) != 0) return rc;
    return launch_gemm_dmma<64>(d, dA, j, j2 + 32, j2 + 32, 0x7fffffff, max_m - j2 - 32, max_n - j2 - 32, batch, il, s);
}

template <int R, int W>
magma_int_t run_step(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, PivRec *recs,
                     int j, long batch, const int *il, cudaStream_t s, bool global_panel, bool defer_left)
{
    const int mp = max_m - j;
    if (global_panel) {
        panel_global_kernel<W><<<(unsigned)batch, 256, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il);
    } else {
        int T = (mp + R - 1) / R;
        T = ((T + 31) / 32) * 32;
        if (T > 512) return MAGMA_ERR_NOT_SUPPORTED;
        if constexpr (R == 1 && W == 32) {
            launch_panel32(d, dA, dipiv, dinfo, recs, j, T, batch, il, s);
        } else {
            panel_kernel<R, W><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, recs, j, batch, il);
            count_launch();
        }
    }
    if (global_panel) count_launch();
    MB200_CHECK_LAUNCH("panel_kernel");

    // fixed size: the panel width is known here; variable size: a matrix on its last, narrower
    // panel can leave up to max_n-j-1 columns to its right
    const int max_mn = max_m < max_n ? max_m : max_n;
    const int jb_host = (max_mn - j) < W ? (max_mn - j) : W;
    const int nright_max = d.vm ? (max_n - j - 1) : (max_n - j - jb_host);
    const int mbelow_max = d.vm ? (max_m - j - 1) : (max_m - j - jb_host);
    // interchanges of the columns to the left: deferred to laswp_left_kernel when every step is 32
    // wide, otherwise applied now (left tiles of swap_trsm_kernel)
    const int left_tiles = (!defer_left && j > 0) ? (j + TN - 1) / TN : 0;
    if (nright_max <= 0 && left_tiles == 0) return 0;
    if (mp <= SROWS && nright_max > 0) {
        // short panel: the whole trailing strip fits in shared memory
        static DevOnce once;
        smem_optin(once, update_strip_kernel<32>, sizeof(StripSmem));
        const int strips = (nright_max + SW - 1) / SW;
        const long grid = (long)strips * batch;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        update_strip_kernel<32><<<(unsigned)grid, STRIP_THREADS, sizeof(StripSmem), s>>>(d, dA, recs, j, strips, batch, il);
        count_launch();
        MB200_CHECK_LAUNCH("update_strip_kernel");
        if (left_tiles > 0) {
            swap_trsm_kernel<W><<<(unsigned)((long)left_tiles * batch), ST_THREADS, 0, s>>>(d, dA, recs, j, 0, left_tiles, 0, -1, batch, il);
            count_launch();
            MB200_CHECK_LAUNCH("swap_trsm_kernel");
        }
        return 0;
    }
    const int right_tiles = nright_max > 0 ? (nright_max + TN - 1) / TN : 0;
    {
        const long grid = (long)(right_tiles + left_tiles) * batch;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        swap_trsm_kernel<W><<<(unsigned)grid, ST_THREADS, 0, s>>>(d, dA, recs, j, right_tiles, left_tiles, 0, -1, batch, il);
        count_launch();
        MB200_CHECK_LAUNCH("swap_trsm_kernel");
    }
    if (mbelow_max > 0 && nright_max > 0) {
        if (W == 32 && g_tier != 4) {  // FP64 tensor pipe (tier 4 forces the DFMA kernel, for A/B runs)
            // region rows/cols >= j + 32: a matrix on a narrower (last) panel has nothing left below or to the right
            const magma_int_t rc = launch_gemm_dmma<32>(d, dA, j, j + 32, j + 32, 0x7fffffff, max_m - j - 32, max_n - j - 32,
                                                        batch, il, s);
            if (rc != 0) return rc;
        } else {
            const int tiles_m = (mbelow_max + GM - 1) / GM, tiles_n = (nright_max + GN - 1) / GN;
            const long grid = (long)tiles_m * tiles_n * batch;
            if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
            gemm_kernel<W><<<(unsigned)grid, GEMM_THREADS, 0, s>>>(d, dA, j, tiles_m, tiles_n, batch, il);
            count_launch();
            MB200_CHECK_LAUNCH("gemm_kernel");
        }
    }
    return 0;
}

}  // namespace

size_t lu_blocked_workspace_bytes(long batch) { return sizeof(PivRec) * (size_t)(batch > 0 ? batch : 0); }

// Step-permutation records of the left-looking driver: ceil(min(m,n)/32) arrays of roundup(m,32) 16-bit rows
// per matrix; 0 when the shape is outside that driver (more than 512 rows, or a single panel).
size_t lu_blocked_perm_bytes(long batch, int max_m, int max_n, bool any_width)
{
    const int mn = max_m < max_n ? max_m : max_n;
    // every shape of at most 512 rows (tier 6 keeps the right-looking flow)
    if (batch <= 0 || (max_n <= 32 && !any_width) || max_m > 512) return 0;
    const size_t rows = (size_t)((max_m + 31) / 32) * 32, blocks = (size_t)(mn + 31) / 32;
    return sizeof(unsigned short) * rows * blocks * (size_t)batch;
}

namespace {

// Left-looking driver (max_m <= 512): per 32-column slab one update kernel (everything to its left, once)
// and one panel kernel; the interchanges of the L columns in one pass at the end.
magma_int_t run_left_looking_one(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, PivRec *recs,
                                 unsigned short *sinv, long batch, const int *il, cudaStream_t s, int nopiv)
{
    const int max_mn = max_m < max_n ? max_m : max_n;
    const int sinv_rows = ((max_m + 31) / 32) * 32, sinv_blocks = (max_mn + 31) / 32;
    const int slabs = (max_n + 31) / 32;
    auto left = [&](int J, int finish, int tail = 0) -> magma_int_t {
        if (max_m <= 128) return launch_left_update<4>(d, dA, sinv, sinv_rows, sinv_blocks, J, finish, batch, il, s, tail, dipiv, dinfo);
        return (max_m <= 256) ? launch_left_update<8>(d, dA, sinv, sinv_rows, sinv_blocks, J, finish, batch, il, s)
                              : launch_left_update<16>(d, dA, sinv, sinv_rows, sinv_blocks, J, finish, batch, il, s);
    };
    for (int J = 0; J < slabs; ++J) {
        magma_int_t rc = 0;
        const int j = 32 * J;
        // fused tail (left_update_kernel<4>): the panel of slab J is factored by the CTA that has just updated it.
        // g_fused_tail: 0 off (default, see the kernel), 1 panels of 33..96 rows, 2 every panel of at most 96 rows
        const int Tj = ((max_m - j + 31) / 32) * 32;
        const bool tail = J > 0 && j < max_mn && !nopiv && max_m <= 128 && Tj <= 96 && g_fused_tail > 0 && (g_fused_tail >= 2 || Tj > 32);
        if (J > 0) rc = left(J, 0, tail ? 1 : 0);
        if (rc != 0) return rc;
        if (j < max_mn) {
            const int T = Tj;
            // single-warp pivot chains (panel_chain_kernel, lu_fused.cu) for panels of 33..128 rows (g_chain_panel = 3: rows >
            // 128 - 32 * level). Same box, panel_kernel -> chain kernel: n = 128 9.72 -> 9.11 ms, n = 96 5.28 -> 5.00,
            // n = 64 4.36 -> 4.25, n = 48 3.79 -> 3.66. The last <= 32-row panel stays with panel_kernel (level 4 costs 0.2-0.4 ms).
            if (tail) {
                // factored inside left_update_kernel
            } else if (!nopiv && T <= 128 && g_chain_panel > 0 && T > 128 - 32 * g_chain_panel) {
                if ((rc = panel_chain_launch(d, dA, dipiv, dinfo, j, T, batch, il, s, sinv, sinv_rows, sinv_blocks)) != 0) return rc;
            } else if (!nopiv && T > 128 && g_tall_panel2 && sinv) {
                launch_panel2(d, dA, dipiv, dinfo, j, T, batch, il, s, sinv, sinv_rows, sinv_blocks);
                MB200_CHECK_LAUNCH("panel2_kernel");
            } else {
                launch_panel32(d, dA, dipiv, dinfo, recs, j, T, batch, il, s, sinv, sinv_rows, sinv_blocks, nopiv);
                MB200_CHECK_LAUNCH("panel_kernel");
            }
            // a last, narrower panel with columns to its right in the same slab (wide matrices). Variable sizes:
            // any panel may be some matrix's last one, the kernel sorts that out per matrix.
            const bool partial_here = d.vm ? (max_n > j + 1) : (max_mn < j + 32 && max_n > max_mn);
            if (partial_here && (rc = left(J, 1)) != 0) return rc;
        }
    }
    if (max_mn > 32 && !nopiv) {  // (no interchanges to apply without pivoting)
        const int blocks = (max_mn - 1) / 32;  // column blocks that have a later panel
        const int T = ((max_m - 32 + 31) / 32) * 32;
        const size_t smem = sizeof(unsigned short) * (size_t)sinv_rows * (size_t)(sinv_blocks - 1);
        const long grid = (long)blocks * batch;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        laswp_left_sinv_kernel<<<(unsigned)grid, T, smem, s>>>(d, dA, sinv, sinv_rows, sinv_blocks, blocks, batch, il);
        count_launch();
        MB200_CHECK_LAUNCH("laswp_left_sinv_kernel");
    }
    return 0;
}

// Helper streams of the batch split below, one set per (device, caller stream); created on first use, kept for the process.
struct SplitCtx {
    cudaStream_t st[3];
    cudaEvent_t fork, join[3];
};
SplitCtx &split_ctx(cudaStream_t s)
{
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, SplitCtx> ctxs;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = ctxs.find({dev, s});
    if (it == ctxs.end()) {
        SplitCtx c;
        for (int i = 0; i < 3; ++i) {
            // (default priority: helper streams at the highest priority, so that their CTAs take freed slots first and the two
            // kinds of kernels mix on every SM, were measured slower everywhere: n = 128 9.17, n = 256 16.53, n = 512 25.67 ms)
            cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming);
        it = ctxs.emplace(std::make_pair(dev, s), c).first;
    }
    return it->second;
}

// The left-looking flow is a strict chain of kernels per matrix -- panel, slab update, panel, ... -- that alternate between
// two kinds of work: pivot chains (latency-bound, a third of the issue slots used) and DMMA slab updates. Matrices are
// independent, so the batch is cut into g_split parts that run the same chain on their own streams: while one part is in a
// panel kernel another is in a slab update, the SMs hold CTAs of both kinds, and every kernel's tail is filled by the other
// parts' work. Scratch records are addressed by position in the batch, so each part simply takes its slice.
magma_int_t run_left_looking(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, PivRec *recs,
                             unsigned short *sinv, long batch, const int *il, cudaStream_t s, int nopiv)
{
    // Measured (same box): a kernel whose grid exceeds the SMs' CTA slots keeps the block scheduler to itself until its
    // tail, so the overlap is the tails only. Two parts pay where a CTA fills an SM (more than 256 rows: n = 512 25.45 ->
    // 24.76 ms, 27 waves of 148 CTAs per kernel) and cost 1-2% below that (n = 128 8.99 -> 9.14, n = 256 16.12 -> 16.37):
    // g_split = 0 (default) splits in two above 256 rows only; 1..4 force a part count.
    int parts = g_split;
    if (parts <= 0) parts = max_m > 256 ? 2 : 1;
    if (parts > 4) parts = 4;
    while (parts > 1 && batch < (long)parts * 1184) --parts;  // at least eight CTAs per SM and part
    if (parts <= 1) return run_left_looking_one(d, max_m, max_n, dA, dipiv, dinfo, recs, sinv, batch, il, s, nopiv);
    const int max_mn = max_m < max_n ? max_m : max_n;
    const size_t sinv_per = (size_t)(((max_m + 31) / 32) * 32) * (size_t)((max_mn + 31) / 32);
    SplitCtx &c = split_ctx(s);
    cudaEventRecord(c.fork, s);
    magma_int_t rc = 0;
    for (int p = 0; p < parts && rc == 0; ++p) {
        const long lo = batch * p / parts, hi = batch * (p + 1) / parts;
        cudaStream_t sp = p == 0 ? s : c.st[p - 1];
        if (p > 0) cudaStreamWaitEvent(sp, c.fork, 0);
        rc = run_left_looking_one(d, max_m, max_n, il ? dA : dA + lo, il ? dipiv : (dipiv ? dipiv + lo : nullptr), il ? dinfo : dinfo + lo,
                                  recs ? recs + lo : nullptr, sinv ? sinv + (size_t)lo * sinv_per : nullptr, hi - lo,
                                  il ? il + lo : nullptr, sp, nopiv);
        if (p > 0) {
            cudaEventRecord(c.join[p - 1], sp);
            cudaStreamWaitEvent(s, c.join[p - 1], 0);
        }
    }
    return rc;
}

}  // namespace

magma_int_t lu_blocked_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                              long batch, const int *index_list, void *workspace, cudaStream_t s, void *perm_workspace,
                              int nopiv)
{
    if (batch <= 0) return 0;
    if (nopiv) {  // no-pivoting LU: the left-looking driver only (register panels of at most 512 rows)
        if (!perm_workspace || max_m > 512) return MAGMA_ERR_NOT_SUPPORTED;
        return run_left_looking(d, max_m, max_n, dA, nullptr, dinfo, reinterpret_cast<PivRec *>(workspace),
                                reinterpret_cast<unsigned short *>(perm_workspace), batch, index_list, s, 1);
    }
    PivRec *recs = reinterpret_cast<PivRec *>(workspace);
    const int max_mn = max_m < max_n ? max_m : max_n;  // upper bound of min(m_b, n_b)
    // tiers 4 (DFMA only), 5 (no pairing), 6 (right-looking) keep the right-looking flow for A/B runs
    if (perm_workspace && max_n > 32 && max_m <= XOVER_LEFT_ROWS && g_tier != 4 && g_tier != 5 &&
        g_tier != 6)
        return run_left_looking(d, max_m, max_n, dA, dipiv, dinfo, recs, reinterpret_cast<unsigned short *>(perm_workspace),
                                batch, index_list, s, 0);
    const bool defer_left = (max_m <= XOVER_LEFT_ROWS);             // every step is 32 wide
    int pair_end_block = 0;                              // column blocks < this were factored in 64-wide pairs
    int j = 0;
    while (j < max_mn) {
        const int mp = max_m - j;
        magma_int_t rc;
        int w;
        // 64-wide pairing: both panels in the register-panel regime, the second one still above the strip
        // regime, tensor-pipe update available, and nothing to the left of j still waiting for pairing
        if (defer_left && g_tier != 4 && g_tier != 5 && mp <= 512 && (mp - 32) > SROWS && (j + 32) < max_mn &&
            j == 32 * pair_end_block) {
            rc = run_pair(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s);
            if (rc != 0) return rc;
            j += 64;
            pair_end_block += 2;
            continue;
        }
        w = panel_width_for_rows(mp);  // the tuning table (common.cuh), also behind magma_get_dgetrf_batched_nbparam
        if (mp > 8192)    rc = run_step<1, 8>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, true, defer_left);
        else if (w == 32) rc = run_step<1, 32>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, false, defer_left);
        else if (w == 16) rc = run_step<2, 16>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, false, defer_left);
        else if (w == 8)  rc = run_step<4, 8>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, false, defer_left);
        else if (w == 4)  rc = run_step<8, 4>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, false, defer_left);
        else              rc = run_step<16, 2>(d, max_m, max_n, dA, dipiv, dinfo, recs, j, batch, index_list, s, false, defer_left);
        if (rc != 0) return rc;
        j += w;
    }
    if (defer_left && max_mn > 32) {
        const int blocks = (max_mn - 1) / 32;  // column blocks that have a later panel
        const int max_rows = max_m - 32;
        const size_t smem = sizeof(int) * 2 * (size_t)max_rows + 16 + sizeof(double) * LSWP_COLS * (size_t)max_rows;
        const long grid = (long)blocks * batch;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        laswp_left_kernel<<<(unsigned)grid, LSWP_THREADS, smem, s>>>(d, dA, dipiv, blocks, max_rows, 32, pair_end_block, batch, index_list);
        count_launch();
        MB200_CHECK_LAUNCH("laswp_left_kernel");
    }
    return 0;
}

}  // namespace mb200
