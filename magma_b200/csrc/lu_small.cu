// Tier S: LU (and fused LU + solve) of matrices with m, n <= 32, register resident.
//
// Replaces magmablas/zgetrf_batched_smallsq_noshfl.cu:34-129 (one matrix per CTA row-group,
// serial shared-memory pivot scan by every thread) and magmablas/zgesv_batched_small.cu:47-162
// (one CTA of n threads per matrix). Design for sm_100a:
//   * a matrix lives in the registers of G lanes, R rows per lane (row = sub + r*G), so a warp
//     factors 32/G matrices at once: n = 16 -> G = 8, R = 2, four matrices per warp (the
//     reference runs one CTA of 16 threads per matrix). R = 2 halves the per-matrix cost of
//     everything that is per-lane rather than per-element: pivot search, bookkeeping, and the
//     shared-memory broadcast of the pivot row (one LDS feeds R rows);
//   * row interchanges are lazy (a lane keeps its rows and only their logical positions `pos`
//     change), the permutation is applied by the final store;
//   * the column loop is straight-line code: every warp-level primitive runs with the full mask
//     (a sub-warp redux.sync mask makes ptxas emit a per-mask emulation loop), the pivot-row
//     stores are predicated, rows that must not be updated use a zero multiplier;
//   * pivot search = 64-bit compare among a lane's own rows, then CREDUX (G = 32) or a log2(G)
//     shuffle butterfly on the high word of |x| + one ballot; the low word and the LAPACK "first
//     maximum" tie-break only run (warp-uniform branch) when two lanes share a high word;
//   * every lane computes the reciprocal of its own candidate while the search is in flight; the
//     winner publishes it with its row through a per-group shared-memory row buffer (128-bit
//     predicated stores / broadcast loads), so the reciprocal is off the critical path;
//   * arithmetic is the canonical order of oracle/lu_oracle.c (reciprocal of the pivot, then
//     a(i,j) = fma(-l, u, a(i,j)) for k increasing; the solve multiplies by the reciprocal of the
//     diagonal like the reference's trsm, magmablas/trsm_template_device.cuh:54-58), so results
//     are bit-identical to it.
#include "common.cuh"

namespace mb200 {

std::atomic<int> g_small_rows{0};  // 0 = tuned default, 1 / 2 = force rows per lane (tests, tuning sweeps)

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr unsigned NOPOS = 0xffffffffu;

__device__ __forceinline__ void sts_pair_if(unsigned addr, double x, double y, unsigned p)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.v2.f64 [%0], {%1, %2};\n\t}"
        :: "r"(addr), "d"(x), "d"(y), "r"(p) : "memory");
}

__device__ __forceinline__ void sts_one_if(unsigned addr, double x, unsigned p)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.f64 [%0], %1;\n\t}"
                 :: "r"(addr), "d"(x), "r"(p) : "memory");
}

__device__ __forceinline__ double2 lds_pair(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ double lds_one(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// pointers fetched from the pointer arrays are device-global: say so (LDG/STG, not generic LD/ST)
__device__ __forceinline__ double ldg_f64(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void stg_f64(double *p, double v)
{
    asm volatile("st.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// max / min over the G lanes of each group; every lane of the warp takes part
template <int G>
__device__ __forceinline__ unsigned group_max(unsigned v)
{
    if (G == 32) return __reduce_max_sync(FULL_MASK, v);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
        const unsigned w = __shfl_xor_sync(FULL_MASK, v, o);
        v = v > w ? v : w;
    }
    return v;
}

template <int G>
__device__ __forceinline__ unsigned group_min(unsigned v)
{
    if (G == 32) return __reduce_min_sync(FULL_MASK, v);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
        const unsigned w = __shfl_xor_sync(FULL_MASK, v, o);
        v = v < w ? v : w;
    }
    return v;
}

constexpr int min_ctas(int N, int R)
{
    // register budget: ~2*N*R data registers + ~40
    return (N * R <= 8) ? 10 : (N * R <= 16) ? 7 : (N * R <= 24) ? 6 : (N * R <= 32) ? 5 : 3;
}

// EXACT: every matrix is N x N (fixed-size batched call with m == n == N): no per-step guards.
template <int N, int G, int R, int NRHS, bool EXACT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, min_ctas(N, R))
lu_small_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo,
                double **__restrict__ dB, int lddb, long batch, const int *__restrict__ index_list)
{
    static_assert(N % 2 == 0 && N <= 32 && G >= 4 && G <= 32 && (R == 1 || R == 2) && G * R <= 32, "shape");
    constexpr int GPW = 32 / G;                          // matrices per warp
    constexpr int NR2 = (NRHS + 1) & ~1;
    constexpr int R0 = N + NR2 + 2;                      // row | rhs | 1/pivot (+pad)
    constexpr int ROWLEN = (R0 % 4 == 2) ? R0 : R0 + 2;  // group stride = 8 banks mod 16
    __shared__ __align__(16) double srow[WARPS_PER_CTA][GPW][2][ROWLEN];

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int grp = lane / G;
    const int sub = lane % G;
    const unsigned gmask = (G == 32) ? FULL_MASK : (((1u << G) - 1u) << (grp * G));

    // No lane leaves early: all warp primitives below use the full mask.
    const long slot = ((long)blockIdx.x * WARPS_PER_CTA + wid) * GPW + grp;
    long b = -1;
    if (slot < batch) b = index_list ? index_list[slot] : slot;
    const bool valid = b >= 0;

    int m = 0, n = 0, ld = 1;
    double *__restrict__ A = nullptr;
    double *B = nullptr;
    if (valid) {
        dims_of(d, b, m, n, ld);
        A = dA[b];
        if (NRHS > 0) B = dB[b];
    }
    const int mn = m < n ? m : n;

    // rows beyond m hold 1.0 so that their (unused) reciprocals stay on the fast path
    double a[R][N];
    double rb[R][NRHS > 0 ? NRHS : 1];
    bool row_ok[R];
    unsigned pos[R];  // logical position of each row this lane holds
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = sub + r * G;
        row_ok[r] = valid && row < m;
        pos[r] = (unsigned)row;
#pragma unroll
        for (int j = 0; j < N; ++j)
            a[r][j] = (row_ok[r] && (EXACT || j < n)) ? ldg_f64(A + row + (size_t)j * ld) : 1.0;
        if (NRHS > 0) {
#pragma unroll
            for (int k = 0; k < NRHS; ++k) rb[r][k] = row_ok[r] ? ldg_f64(B + row + (size_t)k * lddb) : 0.0;
        }
    }

    int myipiv[R];
    double mydinv[R];  // reciprocal of the diagonal of the row that ends at position sub + r*G
#pragma unroll
    for (int r = 0; r < R; ++r) {
        myipiv[r] = 0;
        mydinv[r] = 0.0;
    }
    int info = 0;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(&srow[wid][grp][0][0]);
    const int steps = EXACT ? N : (int)__reduce_max_sync(FULL_MASK, (unsigned)mn);  // warp-uniform

#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (EXACT || i < steps) {
            const bool on = EXACT || (i < mn);  // this group still has a column i
            // ---- this lane's best candidate among its own rows (full 64-bit compare) ----------
            unsigned long long lbits;
            unsigned lpos;
            int w = 0;
            {
                const bool act0 = on && row_ok[0] && (pos[0] >= (unsigned)i);
                lbits = act0 ? ((unsigned long long)__double_as_longlong(a[0][i]) & 0x7fffffffffffffffull) : 0ull;
                lpos = act0 ? pos[0] : NOPOS;
                if (R == 2) {
                    const bool act1 = on && row_ok[R - 1] && (pos[R - 1] >= (unsigned)i);
                    const unsigned long long b1 =
                        act1 ? ((unsigned long long)__double_as_longlong(a[R - 1][i]) & 0x7fffffffffffffffull) : 0ull;
                    const unsigned p1 = act1 ? pos[R - 1] : NOPOS;
                    const bool take1 = (b1 > lbits) || (b1 == lbits && p1 < lpos);
                    w = take1 ? 1 : 0;
                    lbits = take1 ? b1 : lbits;
                    lpos = take1 ? p1 : lpos;
                }
            }
            // reciprocal of this lane's candidate, computed while the search is in flight
            const double lval = (R == 2 && w) ? a[R - 1][i] : a[0][i];
            const double rinv = 1.0 / lval;

            // ---- group-wide search ---------------------------------------------------------------
            const bool lact = (lpos != NOPOS);
            const unsigned hi = (unsigned)(lbits >> 32);
            const unsigned mx = group_max<G>(hi);
            bool cand = lact && (hi == mx);
            unsigned bal = __ballot_sync(FULL_MASK, cand) & gmask;
            const bool unresolved = on && __popc(bal) != 1;  // warp-uniform when G == 32 and EXACT
            if ((G == 32 && EXACT) ? unresolved : __any_sync(FULL_MASK, unresolved)) {
                // some group has colliding high words (or an all-zero column): compare the low
                // words, then break exact ties on the lowest row position like LAPACK's idamax.
                const unsigned lo = cand ? (unsigned)lbits : 0u;
                const unsigned mx2 = group_max<G>(lo);
                cand = cand && (lo == mx2);
                const unsigned kp = cand ? lpos : NOPOS;
                const unsigned mp = group_min<G>(kp);
                cand = cand && (lpos == mp);
                bal = __ballot_sync(FULL_MASK, cand) & gmask;
            }
            const int P = bal ? (__ffs(bal) - 1) : lane;         // lane holding the pivot row
            const unsigned p = __shfl_sync(FULL_MASK, lpos, P);  // its logical position
            const bool is_piv_lane = on && (lane == P);
            if (on) {
                if (sub == (i % G)) myipiv[(i / G) < R ? (i / G) : 0] = (int)p + 1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (is_piv_lane && w == r) pos[r] = (unsigned)i;
                    else if (pos[r] == (unsigned)i) pos[r] = p;
                }
            }

            // ---- publish the pivot row and its reciprocal -------------------------------------
            const unsigned buf = sbase + (unsigned)((i & 1) * ROWLEN * 8);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned pr = (is_piv_lane && w == r) ? 1u : 0u;
#pragma unroll
                for (int j = (i & ~1); j < N; j += 2) sts_pair_if(buf + j * 8, a[r][j], a[r][j + 1], pr);
#pragma unroll
                for (int k = 0; k < NRHS; ++k) sts_one_if(buf + (N + k) * 8, rb[r][k], pr);
            }
            sts_one_if(buf + (N + NR2) * 8, rinv, is_piv_lane ? 1u : 0u);
            __syncwarp();
            const double piv = lds_one(buf + i * 8);
            const double rr = lds_one(buf + (N + NR2) * 8);
            const bool nz = (piv != 0.0);
            if (on && !nz && info == 0) info = i + 1;
            if (NRHS > 0) {
                // the row that just became row i keeps 1/u(i,i) for the back substitution
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (is_piv_lane && w == r) mydinv[r] = rr;
            }
            // Rows that are not updated (already pivoted, padding, singular column) use l = 0:
            // fma(-0, u, a) returns a (a stored -0.0 may become +0.0), so the update needs no
            // branch and no per-element select.
            double l[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool upd = on && nz && row_ok[r] && (pos[r] > (unsigned)i);
                l[r] = upd ? a[r][i] * rr : 0.0;
                if (upd) a[r][i] = l[r];
            }
            // column i+1 first: it feeds the next pivot search and the next reciprocal
            if (i + 1 < N) {
                if (((i + 1) & 1) == 0) {
                    const double2 u = lds_pair(buf + (i + 1) * 8);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        a[r][i + 1] = fma(-l[r], u.x, a[r][i + 1]);
                        a[r][i + 2] = fma(-l[r], u.y, a[r][i + 2]);
                    }
                } else {
                    const double u1 = lds_one(buf + (i + 1) * 8);
#pragma unroll
                    for (int r = 0; r < R; ++r) a[r][i + 1] = fma(-l[r], u1, a[r][i + 1]);
                }
            }
#pragma unroll
            for (int j = ((i + 2) & ~1) + (((i + 1) & 1) == 0 ? 2 : 0); j < N; j += 2) {
                const double2 u = lds_pair(buf + j * 8);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    a[r][j] = fma(-l[r], u.x, a[r][j]);
                    a[r][j + 1] = fma(-l[r], u.y, a[r][j + 1]);
                }
            }
#pragma unroll
            for (int k = 0; k < NRHS; ++k) {
                const double ub = lds_one(buf + (N + k) * 8);
#pragma unroll
                for (int r = 0; r < R; ++r) rb[r][k] = fma(-l[r], ub, rb[r][k]);
            }
        }
    }

    // ---- store the factors in final row order, pivots, info ---------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row_ok[r]) {
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (EXACT || j < n) stg_f64(A + pos[r] + (size_t)j * ld, a[r][j]);
        }
        if (valid && sub + r * G < mn) dipiv[b][sub + r * G] = myipiv[r];
    }
    if (valid && sub == 0) dinfo[b] = info;

    // ---- fused solve: rb holds L^-1 P b; back-substitute with U ------------------------------
    // x(i) = y(i) * (1/u(i,i)); y(q) -= u(q,i) x(i) for q < i  (k decreasing, canonical order)
    if (NRHS > 0) {
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            if (EXACT || i < steps) {
                const bool on = EXACT || (i < n);
                bool own[R];
                bool mine = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    own[r] = on && row_ok[r] && pos[r] == (unsigned)i;
                    mine = mine || own[r];
                }
                const unsigned bq = __ballot_sync(FULL_MASK, mine) & gmask;
                const int Q = bq ? (__ffs(bq) - 1) : lane;
#pragma unroll
                for (int k = 0; k < NRHS; ++k) {
                    double x = (R == 2 && own[R - 1]) ? rb[R - 1][k] * mydinv[R - 1] : rb[0][k] * mydinv[0];
                    x = __shfl_sync(FULL_MASK, x, Q);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (own[r]) rb[r][k] = x;
                        else if (on && row_ok[r] && pos[r] < (unsigned)i) rb[r][k] = fma(-a[r][i], x, rb[r][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row_ok[r] && sub + r * G < n) {
#pragma unroll
                for (int k = 0; k < NRHS; ++k) stg_f64(B + pos[r] + (size_t)k * lddb, rb[r][k]);
            }
        }
    }
}

template <int N, int G, int R, int NRHS>
void launch_one(const Dims &d, bool exact, double **dA, int **dipiv, int *dinfo, double **dB, int lddb,
                long batch, const int *index_list, cudaStream_t s)
{
    constexpr int GPW = 32 / G;
    const long per_cta = WARPS_PER_CTA * GPW;
    const long grid = (batch + per_cta - 1) / per_cta;
    if (exact)
        lu_small_kernel<N, G, R, NRHS, true><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(
            d, dA, dipiv, dinfo, dB, lddb, batch, index_list);
    else
        lu_small_kernel<N, G, R, NRHS, false><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(
            d, dA, dipiv, dinfo, dB, lddb, batch, index_list);
    count_launch();
}

// Shape table: columns rounded up to N, rows covered by G lanes x R rows.
template <int NRHS>
magma_int_t dispatch_n(int max_m, int max_n, const Dims &d, double **dA, int **dipiv, int *dinfo, double **dB,
                       int lddb, long batch, const int *il, cudaStream_t s)
{
    const int rows_pref = g_small_rows;  // 0: tuned default
#define MB200_GO(NN, GG, RR)                                                                       \
    launch_one<NN, GG, RR, NRHS>(d, (!d.vm && d.m == NN && d.n == NN), dA, dipiv, dinfo, dB, lddb, \
                                 batch, il, s)
    const int ncls = (max_n + 3) / 4;  // N = 4 * ncls
    if (max_m <= 8 && max_n <= 8) {
        const bool two = rows_pref != 1;
        if (ncls <= 1) { if (max_m <= 4) MB200_GO(4, 4, 1); else if (two) MB200_GO(4, 4, 2); else MB200_GO(4, 8, 1); }
        else { if (two) MB200_GO(8, 4, 2); else MB200_GO(8, 8, 1); }
    } else if (max_m <= 16 && max_n <= 16) {
        const bool two = rows_pref != 1;
        switch (ncls) {
            case 1: if (two) MB200_GO(4, 8, 2); else MB200_GO(4, 16, 1); break;
            case 2: if (two) MB200_GO(8, 8, 2); else MB200_GO(8, 16, 1); break;
            case 3: if (two) MB200_GO(12, 8, 2); else MB200_GO(12, 16, 1); break;
            default: if (two) MB200_GO(16, 8, 2); else MB200_GO(16, 16, 1); break;
        }
    } else if (max_m <= 32 && max_n <= 32) {
        const bool two = rows_pref == 2;  // default: one row per lane for the 17..32 class
        switch (ncls) {
            case 1: if (two) MB200_GO(4, 16, 2); else MB200_GO(4, 32, 1); break;
            case 2: if (two) MB200_GO(8, 16, 2); else MB200_GO(8, 32, 1); break;
            case 3: if (two) MB200_GO(12, 16, 2); else MB200_GO(12, 32, 1); break;
            case 4: if (two) MB200_GO(16, 16, 2); else MB200_GO(16, 32, 1); break;
            case 5: if (two) MB200_GO(20, 16, 2); else MB200_GO(20, 32, 1); break;
            case 6: if (two) MB200_GO(24, 16, 2); else MB200_GO(24, 32, 1); break;
            case 7: if (two) MB200_GO(28, 16, 2); else MB200_GO(28, 32, 1); break;
            default: if (two) MB200_GO(32, 16, 2); else MB200_GO(32, 32, 1); break;
        }
    } else {
        return -100;
    }
#undef MB200_GO
    MB200_CHECK_LAUNCH("lu_small_kernel");
    return 0;
}

}  // namespace

magma_int_t lu_small_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                            int nrhs, double **dB, int lddb, long batch, const int *index_list,
                            cudaStream_t s)
{
    if (max_m > 32 || max_n > 32 || batch <= 0) return -100;
    if (!d.vm && d.m == d.n && index_list == nullptr && g_small_rows != 9) {  // 9: force the generic kernel (tests)
        const magma_int_t rc = lu_sq_launch(d.n, dA, d.ldda, dipiv, dinfo, nrhs, dB, lddb, batch, s);
        if (rc != -100) return rc;
    }
    if (nrhs == 0) return dispatch_n<0>(max_m, max_n, d, dA, dipiv, dinfo, dB, lddb, batch, index_list, s);
    if (nrhs == 1) return dispatch_n<1>(max_m, max_n, d, dA, dipiv, dinfo, dB, lddb, batch, index_list, s);
    return -100;
}

}  // namespace mb200
