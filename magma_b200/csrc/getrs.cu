// Batched triangular solves with LU factors: getrs, and the standalone laswp / trsm / gemm entry
// points the reference declares publicly.
//
// Replaces src/zgetrs_batched.cpp:118-178 -- row-serial laswp kernel (nrhs threads per CTA), two
// recursive trsm drivers (magmablas/ztrsm_batched_core.cpp:33-350: for n=512 each is 16 leaf
// kernels + 15 cuBLAS gemms, every gemm preceded by three pointer-displacement kernels) or the
// trsv+gemv recursion (magmablas/ztrsv_batched.cu:65-195) -- with ONE kernel per call:
// a CTA owns a tile of right-hand sides in shared memory, applies the interchanges while
// loading it, streams L then U once from HBM/L2 in 32-column blocks (diagonal block solved inside
// a warp with shuffles, rows below/above updated thread-per-row), and stores X.
// Arithmetic is the canonical order of oracle/lu_oracle.c (multiply by the inverted diagonal, as
// the reference's trsm does: magmablas/trsm_template_device.cuh:54-58).
#include "common.cuh"

namespace mb200 {

namespace {

constexpr int SOLVE_THREADS = 256;
constexpr int SB = 32;  // block size along the triangle

// X tile lives in smem as Bs[c*ldb_s + i], i < n, c < tr (column per right-hand side: thread-per-row
// accesses are conflict free). Xs[k*TRW + c] holds the block of unknowns just solved, rhs-contiguous,
// so the update reads it with broadcast 128-bit loads.
constexpr int TRW = 16;  // right-hand sides per CTA tile

// One routine for the four triangular operators op(T):
//   FWD = op(T) is lower triangular (L, or U^T): forward substitution, k increasing;
//   else  op(T) is upper triangular (U, or L^T): backward substitution, k decreasing.
//   TRANS: op(T)(i,k) = A[k + i*ld], else A[i + k*ld].
// Per unknown the update order is the canonical one of oracle/lu_oracle.c.
template <bool UNIT, bool FWD, bool TRANS>
__device__ void solve_tri(int n, int tr, const double *__restrict__ A, int ld, double *Bs, int ldb_s, double *Xs)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = SOLVE_THREADS / 32;
    const int nblk = (n + SB - 1) / SB;
    for (int bi = 0; bi < nblk; ++bi) {
        const int blk = FWD ? bi : nblk - 1 - bi;
        const int kb = blk * SB;
        const int kw = (n - kb) < SB ? (n - kb) : SB;
        // ---- diagonal block: warp per rhs column, lane = row ---------------------------------
        if (wid < tr) {
            double tc[SB];  // tc[k] = op(T)(kb+lane, kb+k)
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const bool need = (k < kw) && (lane < kw) && (FWD ? (lane >= k) : (lane <= k));
                const size_t off = TRANS ? ((size_t)(kb + k) + (size_t)(kb + lane) * ld)
                                         : ((size_t)(kb + lane) + (size_t)(kb + k) * ld);
                tc[k] = need ? A[off] : 0.0;
            }
            const double dinv = (!UNIT && lane < kw) ? 1.0 / A[(size_t)(kb + lane) * (ld + 1)] : 1.0;
            for (int c = wid; c < tr; c += nw) {
                double x = (lane < kw) ? Bs[c * ldb_s + kb + lane] : 0.0;
#pragma unroll
                for (int kk = 0; kk < SB; ++kk) {
                    const int k = FWD ? kk : SB - 1 - kk;
                    if (k < kw) {
                        if (!UNIT && lane == k) x = x * dinv;
                        const double xk = __shfl_sync(0xffffffffu, x, k);
                        if (FWD ? (lane > k) : (lane < k)) x = fma(-tc[k], xk, x);
                    }
                }
                if (lane < kw) {
                    Bs[c * ldb_s + kb + lane] = x;
                    Xs[lane * TRW + c] = x;
                }
            }
        }
        __syncthreads();
        // ---- remaining rows: thread = 2 rows x TRW rhs, k in canonical order --------------------
        const int lo = FWD ? kb + kw : 0;
        const int hi = FWD ? n : kb;
        for (int i0 = lo + tid; i0 < hi; i0 += 2 * SOLVE_THREADS) {
            const int i1 = i0 + SOLVE_THREADS;
            const bool has1 = i1 < hi;
            double acc0[TRW], acc1[TRW];
#pragma unroll
            for (int c = 0; c < TRW; ++c) {
                acc0[c] = (c < tr) ? Bs[c * ldb_s + i0] : 0.0;
                acc1[c] = (c < tr && has1) ? Bs[c * ldb_s + i1] : 0.0;
            }
#pragma unroll
            for (int k8 = 0; k8 < SB; k8 += 8) {
                const int kbase = FWD ? k8 : SB - 8 - k8;
                double t0[8], t1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = kbase + u;
                    const bool ok = k < kw;
                    const size_t o0 = TRANS ? ((size_t)(kb + k) + (size_t)i0 * ld) : ((size_t)i0 + (size_t)(kb + k) * ld);
                    const size_t o1 = TRANS ? ((size_t)(kb + k) + (size_t)i1 * ld) : ((size_t)i1 + (size_t)(kb + k) * ld);
                    t0[u] = ok ? A[o0] : 0.0;
                    t1[u] = (ok && has1) ? A[o1] : 0.0;
                }
#pragma unroll
                for (int uu = 0; uu < 8; ++uu) {
                    const int u = FWD ? uu : 7 - uu;
                    const int k = kbase + u;
                    if (k < kw) {
#pragma unroll
                        for (int c = 0; c < TRW; c += 2) {
                            const double2 x = *reinterpret_cast<const double2 *>(&Xs[k * TRW + c]);
                            acc0[c] = fma(-t0[u], x.x, acc0[c]);
                            acc0[c + 1] = fma(-t0[u], x.y, acc0[c + 1]);
                            acc1[c] = fma(-t1[u], x.x, acc1[c]);
                            acc1[c + 1] = fma(-t1[u], x.y, acc1[c + 1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < TRW; ++c) {
                if (c < tr) {
                    Bs[c * ldb_s + i0] = acc0[c];
                    if (has1) Bs[c * ldb_s + i1] = acc1[c];
                }
            }
        }
        __syncthreads();
    }
}

// perm[i] = original row that LAPACK's forward interchanges ipiv[0..n) leave at position i
__device__ void build_perm(int n, const int *__restrict__ ipiv, int *perm, int *sipiv)
{
    for (int i = threadIdx.x; i < n; i += SOLVE_THREADS) {
        perm[i] = i;
        sipiv[i] = ipiv ? ipiv[i] - 1 : i;  // null: no interchanges (nopiv solves)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) {
            const int p = sipiv[i];
            if (p != i) {
                const int t = perm[i];
                perm[i] = perm[p];
                perm[p] = t;
            }
        }
    }
    __syncthreads();
}

// mode: 0 getrs NoTrans, 1 getrs Trans, 2 laswp only (k1..k2), 3.. trsm variants
__global__ void __launch_bounds__(SOLVE_THREADS, 2)
getrs_kernel(int trans, int n, int nrhs, int tr_max, double **__restrict__ dA, int ldda, int **__restrict__ dipiv,
             double **__restrict__ dB, int lddb, int rhs_tiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long b = blockIdx.x / rhs_tiles;
    const int tile = blockIdx.x % rhs_tiles;
    const int c0 = tile * tr_max;
    const int tr = (nrhs - c0) < tr_max ? (nrhs - c0) : tr_max;
    const int ldb_s = n | 1;  // odd stride: the per-column diagonal-block accesses spread over banks
    double *Bs = reinterpret_cast<double *>(smem_raw);
    double *Xs = Bs + (((size_t)tr_max * ldb_s + 1) & ~(size_t)1);  // 16-byte aligned
    int *perm = reinterpret_cast<int *>(Xs + SB * TRW);
    int *sipiv = perm + n;
    for (int i = threadIdx.x; i < SB * TRW; i += SOLVE_THREADS) Xs[i] = 0.0;
    const double *__restrict__ A = dA[b];
    double *__restrict__ B = dB[b] + (size_t)c0 * lddb;

    build_perm(n, dipiv ? dipiv[b] : nullptr, perm, sipiv);
    if (trans == MagmaNoTrans) {
        for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
            const int i = idx % n, c = idx / n;
            Bs[c * ldb_s + i] = B[perm[i] + (size_t)c * lddb];
        }
        __syncthreads();
        solve_tri<true, true, false>(n, tr, A, ldda, Bs, ldb_s, Xs);    // L y = P b (unit)
        solve_tri<false, false, false>(n, tr, A, ldda, Bs, ldb_s, Xs);  // U x = y
        for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
            const int i = idx % n, c = idx / n;
            B[i + (size_t)c * lddb] = Bs[c * ldb_s + i];
        }
    } else {
        for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
            const int i = idx % n, c = idx / n;
            Bs[c * ldb_s + i] = B[i + (size_t)c * lddb];
        }
        __syncthreads();
        solve_tri<false, true, true>(n, tr, A, ldda, Bs, ldb_s, Xs);   // U^T y = b
        solve_tri<true, false, true>(n, tr, A, ldda, Bs, ldb_s, Xs);   // L^T x = y (unit)
        // inverse of the forward interchanges: x[perm[i]] = y[i]
        for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
            const int i = idx % n, c = idx / n;
            B[perm[i] + (size_t)c * lddb] = Bs[c * ldb_s + i];
        }
    }
}

// -------------------------------------------------------------------------------------------
// getrs, NoTrans, on the FP64 tensor pipe: the right-hand-side tile (n x 16) lives in REGISTERS for
// the whole call, as DMMA accumulator fragments (warp w owns rows [64w, 64w+64): 8 row tiles x 2
// column tiles). Per 32-row block: the block's rows go to shared memory, are solved against the
// diagonal block (staged with cp.async one block ahead; lane = row, one shuffle per unknown), and
// every other row tile receives C -= A_tile * X_blk as mma.sync.m8n8k4.f64 with the A fragments
// loaded straight from global memory (each element of L and U is read exactly once, full sectors).
// DMMA equals the FMA chain in fragment order (tools/dmma_probe.cu), and the backward sweep feeds
// the fragments in reversed k, so every unknown sees the canonical update order of
// oracle/lu_oracle.c: bit-identical results. The DFMA kernel above re-read its tile from shared
// memory per block and waited on dependent global loads inside the update loop (7.4 ms for 4000
// systems of n = 512, nrhs = 16; reference 5.1 ms).
// -------------------------------------------------------------------------------------------
constexpr int XW = 20;  // Xs[k*XW + c]: B-fragment reads (4 rows x 8 columns) are conflict free

__device__ __forceinline__ void dmma_884s(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8_s(void *smem, const void *gmem, bool pred)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}

// diagonal block kb -> Ts[k*32 + i] = A(kb+i, kb+k), zero outside the matrix
__device__ __forceinline__ void stage_diag(double *Ts, const double *__restrict__ A, int ld, int n, int kb)
{
    for (int idx = threadIdx.x; idx < 1024; idx += SOLVE_THREADS) {
        const int i = idx & 31, k = idx >> 5;
        const bool ok = (kb + i < n) && (kb + k < n);
        cp_async8_s(&Ts[idx], ok ? &A[(size_t)(kb + i) + (size_t)(kb + k) * ld] : A, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// perm[i] = original row that LAPACK's forward interchanges leave at position i, computed by
// tracing every position backwards through the interchanges (n independent traces, no serial pass)
__device__ void build_perm_par(int n, const int *__restrict__ ipiv, int *perm, int *sipiv)
{
    for (int i = threadIdx.x; i < n; i += SOLVE_THREADS) sipiv[i] = ipiv ? ipiv[i] - 1 : i;
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < n; i0 += 2 * SOLVE_THREADS) {
        int c0 = i0, c1 = i0 + SOLVE_THREADS;
        for (int k = n - 1; k >= 0; --k) {
            const int p = sipiv[k];
            c0 = (c0 == k) ? p : ((c0 == p) ? k : c0);
            c1 = (c1 == k) ? p : ((c1 == p) ? k : c1);
        }
        perm[i0] = c0;
        if (i0 + SOLVE_THREADS < n) perm[i0 + SOLVE_THREADS] = c1;
    }
    __syncthreads();
}

// Row tiles (8 rows) are dealt to the warps round robin: tile t = 8 i + w is local tile i of warp w, so
// the rows still to be updated are spread evenly over the warps at every block step. Block blk (32 rows) =
// tiles 4 blk .. 4 blk + 3 = local tile blk / 2 of warps 4 (blk & 1) .. 4 (blk & 1) + 3.
template <bool fwd, int RT>
__device__ __forceinline__ void getrs_sweep(double (&acc)[RT][2][2], double *Xs, double *Ts, const double *__restrict__ A,
                                            int ld, int n, int nblk, int w, int lane, int g, int q)
{
#pragma unroll 1
    for (int bi = 0; bi < nblk; ++bi) {
        const int blk = fwd ? bi : (nblk - 1 - bi);
        const int ib = blk >> 1, par = blk & 1;  // the block's rows: local tile ib of warps 4 par .. 4 par + 3
        const int kb = 32 * blk;
        const int seq = fwd ? blk : (2 * nblk - 1 - blk);  // position in the overall block sequence
        double *Tc = Ts + (seq & 1) * 1024;
        const bool owner = (w >> 2) == par;
        double *xrow = Xs + (8 * (w & 3) + g) * XW + 2 * q;
        // ---- the block's rows -> shared memory (acc is indexed statically: uniform switch) ---------------
        if (owner) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
                if (i == ib) {
#pragma unroll
                    for (int jt = 0; jt < 2; ++jt) {
                        xrow[8 * jt] = acc[i][jt][0];
                        xrow[8 * jt + 1] = acc[i][jt][1];
                    }
                }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // next diagonal block in flight while this one is used (the turn reuses the last block)
        {
            const int nseq = seq + 1;
            if (nseq < 2 * nblk) {
                const int nb2 = nseq < nblk ? nseq : (2 * nblk - 1 - nseq);
                stage_diag(Ts + (nseq & 1) * 1024, A, ld, n, 32 * nb2);
            }
        }
        // ---- diagonal block: warp per two right-hand sides, lane = row -----------------------------
        {
            const int ca = 2 * w, cb2 = 2 * w + 1;
            double xa = Xs[lane * XW + ca], xb = Xs[lane * XW + cb2];
            if (fwd) {
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const double t = (lane > k) ? Tc[k * 32 + lane] : 0.0;
                    const double ka = __shfl_sync(0xffffffffu, xa, k), kb2 = __shfl_sync(0xffffffffu, xb, k);
                    xa = fma(-t, ka, xa);
                    xb = fma(-t, kb2, xb);
                }
            } else {
                const double dg = Tc[lane * 32 + lane];
                const double dinv = (kb + lane < n) ? 1.0 / dg : 0.0;
#pragma unroll 8
                for (int kk = 0; kk < 32; ++kk) {
                    const int k = 31 - kk;
                    const double t = (lane < k) ? Tc[k * 32 + lane] : 0.0;
                    if (lane == k) {
                        xa = xa * dinv;
                        xb = xb * dinv;
                    }
                    const double ka = __shfl_sync(0xffffffffu, xa, k), kb2 = __shfl_sync(0xffffffffu, xb, k);
                    xa = fma(-t, ka, xa);
                    xb = fma(-t, kb2, xb);
                }
            }
            Xs[lane * XW + ca] = xa;
            Xs[lane * XW + cb2] = xb;
        }
        __syncthreads();
        // ---- solved rows back to their owner ---------------------------------------------------------------
        if (owner) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
                if (i == ib) {
#pragma unroll
                    for (int jt = 0; jt < 2; ++jt) {
                        acc[i][jt][0] = xrow[8 * jt];
                        acc[i][jt][1] = xrow[8 * jt + 1];
                    }
                }
        }
        // ---- every other row of this sweep: C -= A_tile * X_blk. A fragments of local tile i: rows
        //      8 (8 i + w) + g, columns kb + k(ks, q); loaded one tile ahead of their use ------------------------
        double af[2][8];
        auto load_tile = [&](int i, double (&dst)[8]) {
            const int t = 8 * i + w;
            const bool todo = fwd ? (t >= 4 * blk + 4) : (t < 4 * blk);
            const int r = 8 * t + g;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int k = fwd ? (4 * ks + q) : (31 - (4 * ks + q));
                dst[ks] = (todo && r < n && kb + k < n) ? A[(size_t)r + (size_t)(kb + k) * ld] : 0.0;
            }
        };
        load_tile(0, af[0]);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            if (i + 1 < RT) load_tile(i + 1, af[(i + 1) & 1]);
            const int t = 8 * i + w;
            const bool todo = fwd ? (t >= 4 * blk + 4) : (t < 4 * blk);
            if (todo && 8 * t < n) {  // warp-uniform
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int k = fwd ? (4 * ks + q) : (31 - (4 * ks + q));
                    const double na = -af[i & 1][ks];
#pragma unroll
                    for (int jt = 0; jt < 2; ++jt) dmma_884s(acc[i][jt][0], acc[i][jt][1], na, Xs[k * XW + 8 * jt + g]);
                }
            }
        }
        __syncthreads();  // Xs is rewritten by the next block
    }
}

template <int RT>  // row tiles (8 rows) per warp: n <= 64 * RT
__global__ void __launch_bounds__(SOLVE_THREADS, 2)
getrs_dmma_kernel(int n, int nrhs, double **__restrict__ dA, int ldda, int **__restrict__ dipiv,
                  double **__restrict__ dB, int lddb, int rhs_tiles, int mode, double alpha)
{
    // mode 0: getrs (both sweeps); 1: forward sweep only = trsm Left / Lower / NoTrans / Unit; 2: backward sweep only = trsm
    // Left / Upper / NoTrans / NonUnit (multiply by the inverted diagonal, as the reference's trsm). alpha scales B on the way
    // in (BLAS dtrsm order); the getrs path passes exactly 1.0.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Xs = reinterpret_cast<double *>(smem_raw);  // [32][XW]
    double *Ts = Xs + 32 * XW;                           // [2][32*32]
    int *perm = reinterpret_cast<int *>(Ts + 2 * 1024);
    int *sipiv = perm + n;

    const long b = blockIdx.x / rhs_tiles;
    const int tile = blockIdx.x % rhs_tiles;
    const int c0 = tile * 16;
    const int tr = (nrhs - c0) < 16 ? (nrhs - c0) : 16;
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, q = lane & 3;
    const double *__restrict__ A = dA[b];
    double *__restrict__ B = dB[b] + (size_t)c0 * lddb;
    const int ld = ldda;
    const int nblk = (n + 31) / 32;

    // the first diagonal block of the sequence, where the sweep that runs first expects it (getrs_sweep: Ts + (seq & 1) * 1024)
    if (mode == 2) stage_diag(Ts + (nblk & 1) * 1024, A, ld, n, 32 * (nblk - 1));
    else stage_diag(Ts, A, ld, n, 0);
    build_perm_par(n, dipiv ? dipiv[b] : nullptr, perm, sipiv);

    // right-hand sides -> accumulator fragments, interchanges applied on the way in
    double acc[RT][2][2];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = 8 * (8 * i + w) + g;
        const int pr = (r < n) ? perm[r] : 0;
#pragma unroll
        for (int jt = 0; jt < 2; ++jt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * jt + 2 * q + e;
                acc[i][jt][e] = (r < n && c < tr) ? alpha * B[pr + (size_t)c * lddb] : 0.0;
            }
    }

    // L y = P b (unit lower, blocks ascending), then U x = y (blocks descending)
    if (mode != 2) getrs_sweep<true, RT>(acc, Xs, Ts, A, ld, n, nblk, w, lane, g, q);
    if (mode != 1) getrs_sweep<false, RT>(acc, Xs, Ts, A, ld, n, nblk, w, lane, g, q);
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // mode 1 leaves the turn block's prefetch in flight

#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = 8 * (8 * i + w) + g;
#pragma unroll
        for (int jt = 0; jt < 2; ++jt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * jt + 2 * q + e;
                if (r < n && c < tr) B[r + (size_t)c * lddb] = acc[i][jt][e];
            }
    }
}

// standalone trsm, side = Left.  B <- alpha * op(A)^-1 B
__global__ void __launch_bounds__(SOLVE_THREADS, 2)
trsm_left_kernel(int uplo, int trans, int diag, int n, int nrhs, int tr_max, double alpha,
                 double **__restrict__ dA, int ldda, double **__restrict__ dB, int lddb, int rhs_tiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long b = blockIdx.x / rhs_tiles;
    const int tile = blockIdx.x % rhs_tiles;
    const int c0 = tile * tr_max;
    const int tr = (nrhs - c0) < tr_max ? (nrhs - c0) : tr_max;
    const int ldb_s = n | 1;
    double *Bs = reinterpret_cast<double *>(smem_raw);
    double *Xs = Bs + (((size_t)tr_max * ldb_s + 1) & ~(size_t)1);  // 16-byte aligned
    const double *__restrict__ A = dA[b];
    double *__restrict__ B = dB[b] + (size_t)c0 * lddb;
    for (int i = threadIdx.x; i < SB * TRW; i += SOLVE_THREADS) Xs[i] = 0.0;
    for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
        const int i = idx % n, c = idx / n;
        Bs[c * ldb_s + i] = alpha * B[i + (size_t)c * lddb];
    }
    __syncthreads();
    const bool unit = (diag == MagmaUnit);
    const bool lower = (uplo == MagmaLower);
    const bool nt = (trans == MagmaNoTrans);
    if (lower && nt) { if (unit) solve_tri<true, true, false>(n, tr, A, ldda, Bs, ldb_s, Xs); else solve_tri<false, true, false>(n, tr, A, ldda, Bs, ldb_s, Xs); }
    else if (!lower && nt) { if (unit) solve_tri<true, false, false>(n, tr, A, ldda, Bs, ldb_s, Xs); else solve_tri<false, false, false>(n, tr, A, ldda, Bs, ldb_s, Xs); }
    else if (!lower && !nt) { if (unit) solve_tri<true, true, true>(n, tr, A, ldda, Bs, ldb_s, Xs); else solve_tri<false, true, true>(n, tr, A, ldda, Bs, ldb_s, Xs); }
    else { if (unit) solve_tri<true, false, true>(n, tr, A, ldda, Bs, ldb_s, Xs); else solve_tri<false, false, true>(n, tr, A, ldda, Bs, ldb_s, Xs); }
    for (int idx = threadIdx.x; idx < n * tr; idx += SOLVE_THREADS) {
        const int i = idx % n, c = idx / n;
        B[i + (size_t)c * lddb] = Bs[c * ldb_s + i];
    }
}

// LAPACK-order interchanges k1..k2 (1-based) on n columns; thread per column.
__global__ void laswp_rowserial_kernel(int n, double **__restrict__ dA, int lda, int k1, int k2,
                                       int **__restrict__ dipiv, int col_tiles)
{
    const long b = blockIdx.x / col_tiles;
    const int c = (blockIdx.x % col_tiles) * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double *col = dA[b] + (size_t)c * lda;
    const int *ipiv = dipiv[b];
    for (int i = k1 - 1; i < k2; ++i) {
        const int p = ipiv[i] - 1;
        if (p != i) {
            const double t = col[i];
            col[i] = col[p];
            col[p] = t;
        }
    }
}

// C(Ci.., Cj..) <- alpha A(Ai.., Aj..) B(Bi.., Bj..) + beta C, NoTrans x NoTrans. 32x32 tiles,
// 16x16 threads, 2x2 per thread; k accumulated in increasing order.
__global__ void __launch_bounds__(256)
gemm_nn_kernel(int m, int n, int k, double alpha, double const *const *__restrict__ dA, int Ai, int Aj, int ldda,
               double const *const *__restrict__ dB, int Bi, int Bj, int lddb, double beta,
               double **__restrict__ dC, int Ci, int Cj, int lddc, int mt, int nt)
{
    __shared__ double As[32][33];
    __shared__ double Bsm[32][33];
    const long b = blockIdx.x / (mt * nt);
    const int t = blockIdx.x % (mt * nt);
    const int r0 = (t % mt) * 32, c0 = (t / mt) * 32;
    const double *A = dA[b] + Ai + (size_t)Aj * ldda;
    const double *B = dB[b] + Bi + (size_t)Bj * lddb;
    double *C = dC[b] + Ci + (size_t)Cj * lddc;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < k; k0 += 32) {
        for (int idx = threadIdx.x; idx < 1024; idx += 256) {
            const int i = idx % 32, kk = idx / 32;
            As[kk][i] = (r0 + i < m && k0 + kk < k) ? A[(size_t)(r0 + i) + (size_t)(k0 + kk) * ldda] : 0.0;
            Bsm[i][kk] = (k0 + i < k && c0 + kk < n) ? B[(size_t)(k0 + i) + (size_t)(c0 + kk) * lddb] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double a0 = As[kk][tx], a1 = As[kk][tx + 16];
            const double b0 = Bsm[kk][ty], b1 = Bsm[kk][ty + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int r = r0 + tx + 16 * ii, c = c0 + ty + 16 * jj;
            if (r < m && c < n) {
                double *dst = C + (size_t)r + (size_t)c * lddc;
                *dst = (beta == 0.0) ? alpha * acc[ii][jj] : fma(alpha, acc[ii][jj], beta * (*dst));
            }
        }
}

// rhs columns per CTA such that the tile (+perm, ipiv) fits in shared memory
int pick_tr(int n, int nrhs, size_t &smem)
{
    const size_t cap = 200 * 1024;
    const int ldb_s = n | 1;
    int tr = nrhs < 16 ? nrhs : 16;
    const size_t fixed = (size_t)SB * TRW * 8 + (size_t)n * 8 + 8;  // Xs + perm/ipiv + alignment pad
    while (tr > 1 && (size_t)tr * ldb_s * 8 + fixed > cap) tr >>= 1;
    smem = (size_t)tr * ldb_s * 8 + fixed;
    return smem <= 227 * 1024 ? tr : 0;
}

}  // namespace

magma_int_t getrs_launch(int trans, int n, int nrhs, double **dA, int ldda, int **dipiv, double **dB, int lddb,
                         long batch, cudaStream_t s)
{
    if (trans == MagmaNoTrans && n > 32 && n <= 512 && g_tier != 4) {  // tier 4: DFMA kernels only (A/B runs)
        const int rhs_tiles = (nrhs + 15) / 16;
        const long grid = batch * rhs_tiles;
        if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
        const size_t smem = sizeof(double) * (32 * XW + 2 * 1024) + sizeof(int) * 2 * (size_t)n;
        if (n <= 256)
            getrs_dmma_kernel<4><<<(unsigned)grid, SOLVE_THREADS, smem, s>>>(n, nrhs, dA, ldda, dipiv, dB, lddb, rhs_tiles, 0, 1.0);
        else
            getrs_dmma_kernel<8><<<(unsigned)grid, SOLVE_THREADS, smem, s>>>(n, nrhs, dA, ldda, dipiv, dB, lddb, rhs_tiles, 0, 1.0);
        count_launch();
        MB200_CHECK_LAUNCH("getrs_dmma_kernel");
        return 0;
    }
    size_t smem;
    const int tr = pick_tr(n, nrhs, smem);
    if (tr == 0) return MAGMA_ERR_NOT_SUPPORTED;
    const int rhs_tiles = (nrhs + tr - 1) / tr;
    const long grid = batch * rhs_tiles;
    if (grid > 0x7fffffffL) return MAGMA_ERR_NOT_SUPPORTED;
    static DevOnce once;
    smem_optin(once, getrs_kernel, 227 * 1024);
    getrs_kernel<<<(unsigned)grid, SOLVE_THREADS, smem, s>>>(trans, n, nrhs, tr, dA, ldda, dipiv, dB, lddb, rhs_tiles);
    count_launch();
    MB200_CHECK_LAUNCH("getrs_kernel");
    return 0;
}

void laswp_rowserial_launch(int n, double **dA, int lda, int k1, int k2, int **dipiv, long batch, cudaStream_t s)
{
    if (n <= 0 || batch <= 0 || k2 < k1) return;
    const int threads = n < 128 ? ((n + 31) / 32) * 32 : 128;
    const int col_tiles = (n + threads - 1) / threads;
    const long per = 0x7fffffffL / col_tiles;  // grid.x stays below 2^31: larger batches go in chunks
    for (long off = 0; off < batch; off += per) {
        const long cnt = batch - off < per ? batch - off : per;
        laswp_rowserial_kernel<<<(unsigned)(cnt * col_tiles), threads, 0, s>>>(n, dA + off, lda, k1, k2, dipiv + off, col_tiles);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("laswp_rowserial_kernel");
    }
}

void trsm_left_launch(int uplo, int trans, int diag, int m, int n, double alpha, double **dA, int ldda,
                      double **dB, int lddb, long batch, cudaStream_t s)
{
    if (m <= 0 || n <= 0 || batch <= 0) return;
    // the two solves of an LU -- unit lower forward, non-unit upper backward -- run on the getrs kernel's tensor-pipe sweeps
    // (one sweep each): same canonical order, bit-identical to the shared-memory solver below
    // (m = 128, n = 64, 20000 matrices: 12.7 ms there)
    const bool fwd_unit = uplo == MagmaLower && trans == MagmaNoTrans && diag == MagmaUnit;
    const bool bwd_nonunit = uplo == MagmaUpper && trans == MagmaNoTrans && diag == MagmaNonUnit;
    if ((fwd_unit || bwd_nonunit) && m > 32 && m <= 512 && g_tier != 4) {
        const int rhs_tiles = (n + 15) / 16;
        const size_t smem_d = sizeof(double) * (32 * XW + 2 * 1024) + sizeof(int) * 2 * (size_t)m;
        const long per = 0x7fffffffL / rhs_tiles;
        for (long off = 0; off < batch; off += per) {
            const long cnt = batch - off < per ? batch - off : per;
            if (m <= 256)
                getrs_dmma_kernel<4><<<(unsigned)(cnt * rhs_tiles), SOLVE_THREADS, smem_d, s>>>(m, n, dA + off, ldda, nullptr, dB + off, lddb,
                                                                                             rhs_tiles, fwd_unit ? 1 : 2, alpha);
            else
                getrs_dmma_kernel<8><<<(unsigned)(cnt * rhs_tiles), SOLVE_THREADS, smem_d, s>>>(m, n, dA + off, ldda, nullptr, dB + off, lddb,
                                                                                             rhs_tiles, fwd_unit ? 1 : 2, alpha);
            count_launch();
            MB200_CHECK_LAUNCH_VOID("getrs_dmma_kernel (trsm)");
        }
        return;
    }
    size_t smem;
    const int tr = pick_tr(m, n, smem);
    if (tr == 0) {
        fprintf(stderr, "libmagma_b200: magmablas_dtrsm_batched: m = %d too large for the shared-memory solver\n", m);
        magma_xerbla("magmablas_dtrsm_batched", -MAGMA_ERR_NOT_SUPPORTED);
        return;
    }
    const int rhs_tiles = (n + tr - 1) / tr;
    static DevOnce once;
    smem_optin(once, trsm_left_kernel, 227 * 1024);
    const long per = 0x7fffffffL / rhs_tiles;
    for (long off = 0; off < batch; off += per) {
        const long cnt = batch - off < per ? batch - off : per;
        trsm_left_kernel<<<(unsigned)(cnt * rhs_tiles), SOLVE_THREADS, smem, s>>>(uplo, trans, diag, m, n, tr, alpha, dA + off,
                                                                                ldda, dB + off, lddb, rhs_tiles);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("trsm_left_kernel");
    }
}

void gemm_nn_launch(int m, int n, int k, double alpha, double const *const *dA, int Ai, int Aj, int ldda,
                    double const *const *dB, int Bi, int Bj, int lddb, double beta, double **dC, int Ci, int Cj,
                    int lddc, long batch, cudaStream_t s)
{
    if (m <= 0 || n <= 0 || batch <= 0) return;
    const int mt = (m + 31) / 32, nt = (n + 31) / 32;
    const long per = 0x7fffffffL / ((long)mt * nt);
    for (long off = 0; off < batch; off += per) {
        const long cnt = batch - off < per ? batch - off : per;
        gemm_nn_kernel<<<(unsigned)(cnt * mt * nt), 256, 0, s>>>(m, n, k, alpha, dA + off, Ai, Aj, ldda, dB + off, Bi, Bj, lddb,
                                                               beta, dC + off, Ci, Cj, lddc, mt, nt);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("gemm_nn_kernel");
    }
}

}  // namespace mb200
