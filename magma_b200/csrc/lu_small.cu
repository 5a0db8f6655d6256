// Tier S: LU (and fused LU + solve) of matrices with m, n <= 32, register resident.
//
// Replaces magmablas/zgetrf_batched_smallsq_noshfl.cu:34-129 (one matrix per CTA row-group,
// serial shared-memory pivot scan by every thread) and magmablas/zgesv_batched_small.cu:47-162
// (one CTA of n threads per matrix). Design for sm_100a:
//   * one matrix per G = 32, 16 or 8 lanes, so a warp factors 1, 2 or 4 matrices at once and
//     no lane idles at n = 16 / n = 8 (the reference leaves half a warp per CTA unused at n=16);
//   * lane = row, the row lives in N registers; row interchanges are lazy (a lane keeps its row
//     and only its logical position `pos` changes), the permutation is applied by the final
//     store, which is still a full 8*G-byte coalesced segment per column;
//   * the column loop is straight-line code: every warp-level primitive runs with the full mask
//     (a sub-warp redux.sync mask makes ptxas emit a per-mask emulation loop), stores of the pivot
//     row and the rank-1 update are predicated instructions, not branches;
//   * pivot search = CREDUX (G = 32) or a log2(G) shuffle butterfly (G < 32) on the high word of
//     |x| + one ballot; the low word and the LAPACK "first maximum" tie-break only run (warp-
//     uniform branch) when two candidates share a high word;
//   * every lane computes the reciprocal of its own candidate while the search is in flight, the
//     winner publishes it with its row through a per-group shared-memory row buffer (128-bit
//     stores / broadcast loads), so the reciprocal is off the critical path;
//   * arithmetic is the canonical order of oracle/lu_oracle.c (reciprocal of the pivot, then
//     a(i,j) = fma(-l, u, a(i,j)) for k increasing), so results are bit-identical to it.
#include "common.cuh"

namespace mb200 {

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr unsigned FULL_MASK = 0xffffffffu;

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

// EXACT: every matrix is N x N (fixed-size batched call with m == n == N): no per-step guards.
template <int N, int G, int NRHS, bool EXACT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (N <= 8 ? 10 : N <= 16 ? 7 : N <= 24 ? 6 : 5))
lu_small_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo,
                double **__restrict__ dB, int lddb, long batch, const int *__restrict__ index_list)
{
    static_assert(N % 2 == 0 && N <= 32 && G >= 8 && G <= 32, "shape");
    constexpr int GPW = 32 / G;                        // matrices per warp
    constexpr int NR2 = (NRHS + 1) & ~1;
    constexpr int R0 = N + NR2 + 2;                    // row | rhs | 1/pivot (+pad)
    constexpr int ROWLEN = (R0 % 4 == 2) ? R0 : R0 + 2;  // group stride = 8 banks mod 16: no conflicts
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
    const bool row_ok = valid && sub < m;

    // rows beyond m hold 1.0 so that their (unused) reciprocals stay on the fast path
    double a[N];
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = (row_ok && (EXACT || j < n)) ? ldg_f64(A + sub + (size_t)j * ld) : 1.0;
    double rb[NRHS > 0 ? NRHS : 1];
    if (NRHS > 0) {
#pragma unroll
        for (int k = 0; k < NRHS; ++k) rb[k] = row_ok ? ldg_f64(B + sub + (size_t)k * lddb) : 0.0;
    }

    int pos = sub;  // logical row position of the row this lane holds
    int myipiv = 0;
    int info = 0;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(&srow[wid][grp][0][0]);
    const int steps = EXACT ? N : (int)__reduce_max_sync(FULL_MASK, (unsigned)mn);  // warp-uniform

    double rinv = 1.0 / a[0];  // reciprocal of this lane's candidate for column 0

#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (EXACT || i < steps) {
            const bool on = EXACT || (i < mn);  // this group still has a column i
            // ---- pivot search over rows at positions >= i -------------------------------------
            const bool act = on && row_ok && (pos >= i);
            const unsigned long long bits =
                (unsigned long long)__double_as_longlong(a[i]) & 0x7fffffffffffffffull;
            const unsigned hi = act ? (unsigned)(bits >> 32) : 0u;
            const unsigned mx = group_max<G>(hi);
            bool cand = act && (hi == mx);
            unsigned bal = __ballot_sync(FULL_MASK, cand) & gmask;
            const bool unresolved = on && __popc(bal) != 1;  // warp-uniform when G == 32 and EXACT
            if ((G == 32 && EXACT) ? unresolved : __any_sync(FULL_MASK, unresolved)) {
                // some group has colliding high words (or an all-zero column): compare the low
                // words, then break exact ties on the lowest row position like LAPACK's idamax.
                const unsigned lo = cand ? (unsigned)bits : 0u;
                const unsigned mx2 = group_max<G>(lo);
                cand = cand && (lo == mx2);
                const unsigned kp = cand ? (unsigned)pos : 0xffffffffu;
                const unsigned mp = group_min<G>(kp);
                cand = cand && ((unsigned)pos == mp);
                bal = __ballot_sync(FULL_MASK, cand) & gmask;
            }
            const int P = bal ? (__ffs(bal) - 1) : lane;   // lane holding the pivot row
            const int p = __shfl_sync(FULL_MASK, pos, P);  // its logical position
            const unsigned is_piv = (on && lane == P) ? 1u : 0u;
            if (on) {
                if (sub == i) myipiv = p + 1;
                if (is_piv) pos = i;
                else if (pos == i) pos = p;
            }

            // ---- publish the pivot row and its reciprocal -------------------------------------
            const unsigned buf = sbase + (unsigned)((i & 1) * ROWLEN * 8);
#pragma unroll
            for (int j = (i & ~1); j < N; j += 2) sts_pair_if(buf + j * 8, a[j], a[j + 1], is_piv);
#pragma unroll
            for (int k = 0; k < NRHS; ++k) sts_one_if(buf + (N + k) * 8, rb[k], is_piv);
            sts_one_if(buf + (N + NR2) * 8, rinv, is_piv);
            __syncwarp();
            const double piv = lds_one(buf + i * 8);
            const double r = lds_one(buf + (N + NR2) * 8);
            const bool nz = (piv != 0.0);
            if (on && !nz && info == 0) info = i + 1;
            // Rows that are not updated (already pivoted, padding, singular column) use l = 0:
            // fma(-0, u, a) returns a (a stored -0.0 may become +0.0), so the update needs no
            // branch and no per-element select.
            const bool upd = on && nz && row_ok && (pos > i);
            const double l = upd ? a[i] * r : 0.0;
            if (upd) a[i] = l;
            // column i+1 first: it feeds the next pivot search and the next reciprocal
            if (i + 1 < N) {
                if (((i + 1) & 1) == 0) {
                    const double2 u = lds_pair(buf + (i + 1) * 8);
                    a[i + 1] = fma(-l, u.x, a[i + 1]);
                    a[i + 2] = fma(-l, u.y, a[i + 2]);
                } else {
                    const double u1 = lds_one(buf + (i + 1) * 8);
                    a[i + 1] = fma(-l, u1, a[i + 1]);
                }
                rinv = 1.0 / a[i + 1];
            }
#pragma unroll
            for (int j = ((i + 2) & ~1) + (((i + 1) & 1) == 0 ? 2 : 0); j < N; j += 2) {
                const double2 u = lds_pair(buf + j * 8);
                a[j] = fma(-l, u.x, a[j]);
                a[j + 1] = fma(-l, u.y, a[j + 1]);
            }
#pragma unroll
            for (int k = 0; k < NRHS; ++k) rb[k] = fma(-l, lds_one(buf + (N + k) * 8), rb[k]);
        }
    }

    // ---- store the factors in final row order, pivots, info ---------------------------------
    if (row_ok) {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (EXACT || j < n) stg_f64(A + pos + (size_t)j * ld, a[j]);
    }
    if (valid && sub < mn) dipiv[b][sub] = myipiv;
    if (valid && sub == 0) dinfo[b] = info;

    // ---- fused solve: rb holds L^-1 P b; back-substitute with U (divide by the diagonal) -----
    if (NRHS > 0) {
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            if (EXACT || i < steps) {
                const bool on = EXACT || (i < n);
                const unsigned bq = __ballot_sync(FULL_MASK, on && row_ok && pos == i) & gmask;
                const int Q = bq ? (__ffs(bq) - 1) : lane;
#pragma unroll
                for (int k = 0; k < NRHS; ++k) {
                    double x = rb[k] / a[i];
                    x = __shfl_sync(FULL_MASK, x, Q);
                    if (on && pos == i) rb[k] = x;
                    else if (on && pos < i) rb[k] = fma(-a[i], x, rb[k]);
                }
            }
        }
        if (valid && sub < n) {
#pragma unroll
            for (int k = 0; k < NRHS; ++k) stg_f64(B + pos + (size_t)k * lddb, rb[k]);
        }
    }
}

template <int N, int G, int NRHS>
void launch_one(const Dims &d, bool exact, double **dA, int **dipiv, int *dinfo, double **dB, int lddb,
                long batch, const int *index_list, cudaStream_t s)
{
    constexpr int GPW = 32 / G;
    const long per_cta = WARPS_PER_CTA * GPW;
    const long grid = (batch + per_cta - 1) / per_cta;
    if (exact)
        lu_small_kernel<N, G, NRHS, true><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(
            d, dA, dipiv, dinfo, dB, lddb, batch, index_list);
    else
        lu_small_kernel<N, G, NRHS, false><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(
            d, dA, dipiv, dinfo, dB, lddb, batch, index_list);
    count_launch();
}

template <int NRHS>
magma_int_t dispatch_n(int K, const Dims &d, double **dA, int **dipiv, int *dinfo, double **dB,
                       int lddb, long batch, const int *il, cudaStream_t s)
{
#define MB200_CASE(NN, GG)                                                                    \
    launch_one<NN, GG, NRHS>(d, (!d.vm && d.m == NN && d.n == NN), dA, dipiv, dinfo, dB, lddb, \
                             batch, il, s);                                                   \
    break;
    switch ((K + 3) / 4) {
        case 0:
        case 1: MB200_CASE(4, 8)
        case 2: MB200_CASE(8, 8)
        case 3: MB200_CASE(12, 16)
        case 4: MB200_CASE(16, 16)
        case 5: MB200_CASE(20, 32)
        case 6: MB200_CASE(24, 32)
        case 7: MB200_CASE(28, 32)
        case 8: MB200_CASE(32, 32)
        default: return -100;
    }
#undef MB200_CASE
    MB200_CHECK_LAUNCH("lu_small_kernel");
    return 0;
}

}  // namespace

magma_int_t lu_small_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo,
                            int nrhs, double **dB, int lddb, long batch, const int *index_list,
                            cudaStream_t s)
{
    const int K = max_m > max_n ? max_m : max_n;
    if (K > 32 || batch <= 0) return -100;
    if (nrhs == 0) return dispatch_n<0>(K, d, dA, dipiv, dinfo, dB, lddb, batch, index_list, s);
    if (nrhs == 1) return dispatch_n<1>(K, d, dA, dipiv, dinfo, dB, lddb, batch, index_list, s);
    return -100;
}

}  // namespace mb200
