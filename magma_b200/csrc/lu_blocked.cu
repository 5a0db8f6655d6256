// Blocked right-looking LU for any m x n: two kernels per panel step, no scratch memory.
//
// Replaces the reference's host recursion (src/zgetrf_batched.cpp:149-203 ->
// src/zgetrf_panel_batched.cpp:101-196 -> src/zgetf2_batched.cpp:95-287), which issues 42/105/231
// launches per call at n = 128/256/512: fused panel, setup_pivinfo, adjust_ipiv, 2x laswp,
// recursive trsm, cublasDgemmBatched (+3 pointer-displacement kernels each). Here a step is
//   panel_kernel  : one CTA per matrix, panel rows in registers (thread = R rows x W columns),
//                   CREDUX + shared-memory two-level pivot search, lazy row interchanges,
//                   writes the factored panel in final row order and ipiv (global indices);
//   update_kernel : one CTA per (matrix, 64-column tile). Rebuilds the step's net row
//                   permutation from ipiv in one warp, applies it to the tile (left tiles: swap
//                   only), solves the W x 64 block row against the unit-lower L11, and applies
//                   the rank-W update to every row below, accumulating in place with one fma
//                   per k in increasing k (bit-identical to oracle/lu_oracle.c).
// so n = 128/256/512 take 8/16/32 launches, and pivinfo/adjust/displace kernels do not exist.
#include "common.cuh"

namespace mb200 {

namespace {

// -------------------------------------------------------------------------------------------
// (bits, pos) arg-max over a warp: larger |x| bit pattern wins, ties go to the smaller pos.
// Every lane returns the winner's values.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_argmax(unsigned long long bits, int pos,
                                            unsigned long long &wbits, int &wpos)
{
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)(bits >> 32);
    const unsigned mx = __reduce_max_sync(full, hi);
    bool cand = (hi == mx);
    unsigned bal = __ballot_sync(full, cand);
    if (__popc(bal) != 1) {
        const unsigned lo = cand ? (unsigned)bits : 0u;
        const unsigned mx2 = __reduce_max_sync(full, lo);
        cand = cand && (lo == mx2);
        bal = __ballot_sync(full, cand);
        if (__popc(bal) != 1) {
            const unsigned kp = cand ? (unsigned)pos : 0xffffffffu;
            const unsigned mp = __reduce_min_sync(full, kp);
            cand = cand && ((unsigned)pos == mp);
            bal = __ballot_sync(full, cand);
        }
    }
    const int wl = __ffs(bal) - 1;
    wbits = __shfl_sync(full, bits, wl);
    wpos = __shfl_sync(full, pos, wl);
}

constexpr int NOPOS = 0x7fffffff;

// -------------------------------------------------------------------------------------------
// Panel factorisation, registers. Thread t owns panel rows t, t+T, ... (R of them), W columns.
// -------------------------------------------------------------------------------------------
template <int R, int W>
__global__ void __launch_bounds__(512)
panel_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo, int j,
             long batch, const int *__restrict__ index_list)
{
    __shared__ unsigned long long cbits[2][32];
    __shared__ int cpos[2][32];
    __shared__ __align__(16) double prow[2][W];
    __shared__ int sipiv[W];

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

    double a[R][W];
    int pos[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int r = tid + k * T;
        pos[k] = (r < mp) ? r : NOPOS;
#pragma unroll
        for (int c = 0; c < W; ++c) a[k][c] = (r < mp && c < jb) ? A[r + (size_t)c * ld] : 0.0;
    }
    int info = 0;

#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i < jb) {
            // local best over this thread's rows
            unsigned long long lb = 0;
            int lp = NOPOS;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool act = (pos[k] >= i) && (pos[k] != NOPOS);
                const unsigned long long v =
                    (unsigned long long)__double_as_longlong(a[k][i]) & 0x7fffffffffffffffull;
                if (act && (v > lb || (v == lb && pos[k] < lp))) {
                    lb = v;
                    lp = pos[k];
                }
            }
            unsigned long long wb;
            int wp;
            warp_argmax(lb, lp, wb, wp);
            if (lane == 0) {
                cbits[i & 1][wid] = wb;
                cpos[i & 1][wid] = wp;
            }
            __syncthreads();
            {
                unsigned long long eb = (lane < nw) ? cbits[i & 1][lane] : 0ull;
                int ep = (lane < nw) ? cpos[i & 1][lane] : NOPOS;
                warp_argmax(eb, ep, wb, wp);
            }
            const int ppos = wp;  // panel-relative position of the pivot row (>= i)
            if (tid == 0) sipiv[i] = ppos;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (pos[k] == ppos) {
                    pos[k] = i;
#pragma unroll
                    for (int c = 0; c < W; ++c)
                        if (c >= i) prow[i & 1][c] = a[k][c];
                } else if (pos[k] == i) {
                    pos[k] = ppos;
                }
            }
            __syncthreads();
            const double piv = prow[i & 1][i];
            if (piv != 0.0) {
                const double r = 1.0 / piv;
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
            } else if (info == 0) {
                info = i + 1;
            }
        }
    }

    // factored panel, rows in final order
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (pos[k] != NOPOS) {
#pragma unroll
            for (int c = 0; c < W; ++c)
                if (c < jb) A[pos[k] + (size_t)c * ld] = a[k][c];
        }
    }
    if (tid < jb) dipiv[b][j + tid] = j + sipiv[tid] + 1;
    if (tid == 0) {
        if (j == 0) dinfo[b] = info ? info : 0;
        else if (info && dinfo[b] == 0) dinfo[b] = j + info;
    }
}

// -------------------------------------------------------------------------------------------
// Panel factorisation straight on global memory: correctness fallback for panels taller than
// the register kernel covers (m - j > 8192). One CTA per matrix, W columns.
// -------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
panel_global_kernel(Dims d, double **dA, int **dipiv, int *dinfo, int j, long batch, const int *index_list)
{
    __shared__ unsigned long long cbits[8];
    __shared__ int cpos[8];
    __shared__ int spiv;
    const long slot = blockIdx.x;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;  // unused tail of a vbatched index list
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
        if (tid == 0) dipiv[b][j + i] = j + p + 1;
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
        if (j == 0) dinfo[b] = info;
        else if (info && dinfo[b] == 0) dinfo[b] = j + info;
    }
}

// -------------------------------------------------------------------------------------------
// Update kernel.
// -------------------------------------------------------------------------------------------
constexpr int TN = 64;    // columns per CTA
constexpr int TM = 128;   // rows per GEMM chunk
constexpr int KB = 32;    // max panel width
constexpr int UPD_THREADS = 256;

struct UpdSmem {
    double As[KB * TM];       // L21 chunk, k-major: As[k*TM + r]      (aliases T0/E0 staging)
    double Bs[KB * TN];       // U12 tile,  k-major: Bs[k*TN + c]
    double Ls[KB * (KB + 1)]; // L11, Ls[i*(KB+1) + k]
    int top_src[KB];          // original (panel-relative) row now at top position k
    int top_ext[KB];          // if top_src[k] >= jb: index of that row in the extra list
    int ext_pos[KB];          // extra positions (panel-relative, >= jb)
    int ext_src[KB];          // original top row (< jb) that ends at ext_pos[e]
    int n_ext;
};

template <int W>
__global__ void __launch_bounds__(UPD_THREADS, 2)
update_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int j, int right_tiles,
              int left_tiles, long batch, const int *__restrict__ index_list)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UpdSmem &S = *reinterpret_cast<UpdSmem *>(smem_raw);

    const int tiles = right_tiles + left_tiles;
    const long slot = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;  // unused tail of a vbatched index list
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    if (j >= mn) return;
    const int jb = (mn - j) < W ? (mn - j) : W;

    // column range of this tile
    int c0, c1;
    bool right;
    if (tile < right_tiles) {
        right = true;
        c0 = j + jb + tile * TN;
        c1 = c0 + TN < n ? c0 + TN : n;
    } else {
        right = false;
        c0 = (tile - right_tiles) * TN;
        c1 = c0 + TN < j ? c0 + TN : j;
    }
    if (c0 >= c1) return;
    const int wt = c1 - c0;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *__restrict__ A = dA[b];
    const int *__restrict__ ipiv = dipiv[b] + j;

    // ---- 1. net permutation of this step's interchanges (warp 0) ------------------------------
    if (wid == 0) {
        const unsigned full = 0xffffffffu;
        int cur_top = lane;           // content of top position `lane`
        int epos = -1, econt = -1;    // extra list lives in lanes 0..ne-1
        int ne = 0;
        const int myp = (lane < jb) ? (ipiv[lane] - 1 - j) : 0;
        for (int i = 0; i < jb; ++i) {
            const int p = __shfl_sync(full, myp, i);
            if (p == i) continue;
            const int ci = __shfl_sync(full, cur_top, i);
            if (p < jb) {
                const int cp = __shfl_sync(full, cur_top, p);
                if (lane == i) cur_top = cp;
                if (lane == p) cur_top = ci;
            } else {
                const unsigned hit = __ballot_sync(full, lane < ne && epos == p);
                if (hit) {
                    const int e = __ffs(hit) - 1;
                    const int ce = __shfl_sync(full, econt, e);
                    if (lane == i) cur_top = ce;
                    if (lane == e) econt = ci;
                } else {
                    if (lane == i) cur_top = p;
                    if (lane == ne) { epos = p; econt = ci; }
                    ++ne;
                }
            }
        }
        // map top rows that came from below to their slot in the extra list
        int te = -1;
        for (int e = 0; e < ne; ++e) {
            const int pe = __shfl_sync(full, epos, e);
            if (cur_top == pe) te = e;
        }
        if (lane < KB) {
            S.top_src[lane] = cur_top;
            S.top_ext[lane] = te;
            S.ext_pos[lane] = epos;
            S.ext_src[lane] = econt;
        }
        if (lane == 0) S.n_ext = ne;
    }
    __syncthreads();
    const int ne = S.n_ext;

    // ---- 2. apply it to this tile's columns ---------------------------------------------------
    // T0 = original top rows, E0 = original extra rows (staged in As)
    double *T0 = S.As;            // [k*TN + c]
    double *E0 = S.As + KB * TN;  // [e*TN + c]
    double *Ap = A + (size_t)j;   // row origin of the panel
    for (int idx = tid; idx < jb * wt; idx += UPD_THREADS) {
        const int k = idx % jb, c = idx / jb;
        T0[k * TN + c] = Ap[k + (size_t)(c0 + c) * ld];
    }
    for (int idx = tid; idx < ne * wt; idx += UPD_THREADS) {
        const int e = idx % ne, c = idx / ne;
        E0[e * TN + c] = Ap[S.ext_pos[e] + (size_t)(c0 + c) * ld];
    }
    __syncthreads();
    for (int idx = tid; idx < ne * wt; idx += UPD_THREADS) {
        const int e = idx % ne, c = idx / ne;
        Ap[S.ext_pos[e] + (size_t)(c0 + c) * ld] = T0[S.ext_src[e] * TN + c];
    }
    if (!right) {
        for (int idx = tid; idx < jb * wt; idx += UPD_THREADS) {
            const int k = idx % jb, c = idx / jb;
            const int src = S.top_src[k];
            if (src != k) {
                const double v = (src < jb) ? T0[src * TN + c] : E0[S.top_ext[k] * TN + c];
                Ap[k + (size_t)(c0 + c) * ld] = v;
            }
        }
        return;
    }
    for (int idx = tid; idx < KB * TN; idx += UPD_THREADS) {
        const int k = idx / TN, c = idx % TN;
        double v = 0.0;
        if (k < jb && c < wt) {
            const int src = S.top_src[k];
            v = (src < jb) ? T0[src * TN + c] : E0[S.top_ext[k] * TN + c];
        }
        S.Bs[k * TN + c] = v;
    }
    // L11 (unit lower) -> shared
    for (int idx = tid; idx < jb * jb; idx += UPD_THREADS) {
        const int i = idx % jb, k = idx / jb;
        S.Ls[i * (KB + 1) + k] = Ap[i + (size_t)(j + k) * ld];
    }
    __syncthreads();

    // ---- 3. U12 = L11^-1 * top block: one thread per column, canonical order --------------------
    if (tid < wt) {
        double x[W];
#pragma unroll
        for (int i = 0; i < W; ++i) x[i] = S.Bs[i * TN + tid];
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (k < jb) {
#pragma unroll
                for (int i = 0; i < W; ++i)
                    if (i > k && i < jb) x[i] = fma(-S.Ls[i * (KB + 1) + k], x[k], x[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (i < jb) {
                S.Bs[i * TN + tid] = x[i];
                Ap[i + (size_t)(c0 + tid) * ld] = x[i];
            }
        }
    }
    __syncthreads();

    // ---- 4. rows below: C(r, c) = fma(-L21(r,k), U12(k,c), C(r,c)), k increasing ------------------
    const int r_begin = j + jb;
    if (r_begin >= m) return;
    // warp grid 4 x 2 over a 128 x 64 CTA tile; lane grid 8 x 4; thread tile 4 rows x 8 cols:
    // rows {2lr, 2lr+1, 16+2lr, 17+2lr}, cols {8q + 2lc, 8q + 2lc + 1 : q = 0..3} of the warp tile.
    const int wr = wid & 3, wc = wid >> 2;
    const int lr = lane & 7, lc = lane >> 3;
    const int trow = wr * 32 + 2 * lr;   // + {0,1,16,17}
    const int tcol = wc * 32 + 2 * lc;   // + 8q + {0,1}
    const double *__restrict__ L21 = A + (size_t)j * ld;  // column j, absolute rows
    const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

    for (int r0 = r_begin; r0 < m; r0 += TM) {
        __syncthreads();  // previous chunk's As (or the T0/E0 staging) no longer needed
        const int rows = (m - r0) < TM ? (m - r0) : TM;
        for (int idx = tid; idx < KB * TM; idx += UPD_THREADS) {
            const int r = idx % TM, k = idx / TM;
            S.As[k * TM + r] = (r < rows && k < jb) ? L21[(size_t)(r0 + r) + (size_t)k * ld] : 0.0;
        }
        __syncthreads();

        double acc[4][8];
        // load C
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = tcol + 8 * q + cc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = trow + 16 * h;
                    double v0 = 0.0, v1 = 0.0;
                    if (c < wt && r < rows) {
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
        }
#pragma unroll 4
        for (int k = 0; k < W; ++k) {
            if (k < jb) {
                const double2 a01 = *reinterpret_cast<const double2 *>(&S.As[k * TM + trow]);
                const double2 a23 = *reinterpret_cast<const double2 *>(&S.As[k * TM + trow + 16]);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 bq = *reinterpret_cast<const double2 *>(&S.Bs[k * TN + tcol + 8 * q]);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        acc[rr][2 * q] = fma(-av[rr], bq.x, acc[rr][2 * q]);
                        acc[rr][2 * q + 1] = fma(-av[rr], bq.y, acc[rr][2 * q + 1]);
                    }
                }
            }
        }
        // store C
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = tcol + 8 * q + cc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = trow + 16 * h;
                    if (c < wt && r < rows) {
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
    }
}

template <int R, int W>
magma_int_t run_step(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, int j,
                     long batch, const int *il, cudaStream_t s, bool global_panel)
{
    const int mp = max_m - j;
    if (global_panel) {
        panel_global_kernel<W><<<(unsigned)batch, 256, 0, s>>>(d, dA, dipiv, dinfo, j, batch, il);
    } else {
        int T = (mp + R - 1) / R;
        T = ((T + 31) / 32) * 32;
        if (T > 512) return MAGMA_ERR_NOT_SUPPORTED;
        panel_kernel<R, W><<<(unsigned)batch, T, 0, s>>>(d, dA, dipiv, dinfo, j, batch, il);
    }
    count_launch();
    MB200_CHECK_LAUNCH("panel_kernel");

    // tiles: right part starts at j + jb (jb = W except on a matrix's last, narrower panel)
    // fixed size: the panel width is known here; variable size: a matrix on its last, narrower
    // panel can leave up to max_n-j-1 columns to its right
    const int max_mn = max_m < max_n ? max_m : max_n;
    const int jb_host = (max_mn - j) < W ? (max_mn - j) : W;
    const int nright_max = d.vm ? (max_n - j - 1) : (max_n - j - jb_host);
    const int right_tiles = nright_max > 0 ? (nright_max + TN - 1) / TN : 0;
    const int left_tiles = j > 0 ? (j + TN - 1) / TN : 0;
    const int tiles = right_tiles + left_tiles;
    if (tiles > 0) {
        static bool attr_set = false;
        const size_t smem = sizeof(UpdSmem);
        if (!attr_set) {
            cudaFuncSetAttribute(update_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        const long grid = (long)tiles * batch;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        update_kernel<W><<<(unsigned)grid, UPD_THREADS, smem, s>>>(d, dA, dipiv, j, right_tiles, left_tiles,
                                                                  batch, il);
        count_launch();
        MB200_CHECK_LAUNCH("update_kernel");
    }
    return 0;
}

}  // namespace

magma_int_t lu_blocked_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                              long batch, const int *index_list, cudaStream_t s)
{
    if (batch <= 0) return 0;
    const int max_mn = max_m < max_n ? max_m : max_n;  // upper bound of min(m_b, n_b)
    int j = 0;
    while (j < max_mn) {
        const int mp = max_m - j;
        magma_int_t rc;
        int w;
        if (mp <= 512)       { w = 32; rc = run_step<1, 32>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, false); }
        else if (mp <= 1024) { w = 16; rc = run_step<2, 16>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, false); }
        else if (mp <= 2048) { w = 8;  rc = run_step<4, 8>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, false); }
        else if (mp <= 4096) { w = 4;  rc = run_step<8, 4>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, false); }
        else if (mp <= 8192) { w = 2;  rc = run_step<16, 2>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, false); }
        else                 { w = 8;  rc = run_step<1, 8>(d, max_m, max_n, dA, dipiv, dinfo, j, batch, index_list, s, true); }
        if (rc != 0) return rc;
        j += w;
    }
    return 0;
}

}  // namespace mb200
