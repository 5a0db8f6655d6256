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
//   * pivot search = CREDUX (redux.sync.max.u32) on the high word of |x| + one ballot; the low
//     word and the LAPACK "first maximum" tie-break only run when the high words collide;
//   * the pivot row is broadcast through a per-group shared-memory row buffer with 128-bit
//     stores/loads (a shuffle broadcast would need 2 SHFL per element and is issue-bound);
//   * arithmetic is the canonical order of oracle/lu_oracle.c (reciprocal of the pivot, then
//     a(i,j) = fma(-l, u, a(i,j)) for k increasing), so results are bit-identical to it.
#include "common.cuh"

namespace mb200 {

namespace {

constexpr int WARPS_PER_CTA = 4;

template <int N, int G, int NRHS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
lu_small_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo,
                double **__restrict__ dB, int lddb, long batch, const int *__restrict__ index_list)
{
    static_assert(N % 2 == 0 && N <= 32 && G >= 8 && G <= 32, "shape");
    constexpr int GPW = 32 / G;                      // matrices per warp
    constexpr int ROWLEN = N + ((NRHS + 1) & ~1) + 2;  // +2 doubles staggers the groups' banks
    __shared__ __align__(16) double srow[WARPS_PER_CTA][GPW][2][ROWLEN];

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int grp = lane / G;
    const int sub = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));

    const long slot = ((long)blockIdx.x * WARPS_PER_CTA + wid) * GPW + grp;
    if (slot >= batch) return;  // masks below only name lanes of this group
    const long b = index_list ? index_list[slot] : slot;
    if (b < 0) return;  // unused tail of a vbatched index list

    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    double *__restrict__ A = dA[b];

    double a[N];
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = (sub < m && j < n) ? A[sub + (size_t)j * ld] : 0.0;

    double rb[NRHS > 0 ? NRHS : 1];
    double *B = nullptr;
    if (NRHS > 0) {
        B = dB[b];
#pragma unroll
        for (int k = 0; k < NRHS; ++k) rb[k] = (sub < m) ? B[sub + (size_t)k * lddb] : 0.0;
    }

    int pos = sub;  // logical row position of the row this lane holds
    int myipiv = 0;
    int info = 0;
    double *const buf0 = &srow[wid][grp][0][0];

#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < mn) {
            // ---- pivot search over rows at positions >= i -------------------------------------
            const bool act = (pos >= i) && (sub < m);
            const unsigned long long bits =
                (unsigned long long)__double_as_longlong(a[i]) & 0x7fffffffffffffffull;
            const unsigned hi = act ? (unsigned)(bits >> 32) : 0u;
            const unsigned mx = __reduce_max_sync(gmask, hi);
            bool cand = act && (hi == mx);
            unsigned bal = __ballot_sync(gmask, cand);
            if (__popc(bal) != 1) {  // high words collide (or all are zero): compare low words
                const unsigned lo = cand ? (unsigned)bits : 0u;
                const unsigned mx2 = __reduce_max_sync(gmask, lo);
                cand = cand && (lo == mx2);
                bal = __ballot_sync(gmask, cand);
                if (__popc(bal) != 1) {  // exact tie in |x|: LAPACK takes the first row
                    const unsigned kp = cand ? (unsigned)pos : 0xffffffffu;
                    const unsigned mp = __reduce_min_sync(gmask, kp);
                    cand = cand && ((unsigned)pos == mp);
                    bal = __ballot_sync(gmask, cand);
                }
            }
            const int P = __ffs(bal) - 1;                 // lane holding the pivot row
            const int p = __shfl_sync(gmask, pos, P);     // its logical position
            if (sub == i) myipiv = p + 1;
            if (lane == P) pos = i;
            else if (pos == i) pos = p;

            // ---- broadcast the pivot row ------------------------------------------------------
            double *const buf = buf0 + (i & 1) * ROWLEN;
            if (lane == P) {
#pragma unroll
                for (int j = (i & ~1); j < N; j += 2)
                    *reinterpret_cast<double2 *>(buf + j) = make_double2(a[j], a[j + 1]);
#pragma unroll
                for (int k = 0; k < NRHS; ++k) buf[N + k] = rb[k];
            }
            __syncwarp(gmask);
            const double piv = buf[i];
            if (piv != 0.0) {
                if (pos > i) {
                    const double r = 1.0 / piv;
                    const double l = a[i] * r;
                    a[i] = l;
#pragma unroll
                    for (int j = ((i + 1) & ~1); j < N; j += 2) {
                        const double2 u = *reinterpret_cast<const double2 *>(buf + j);
                        if (j > i) a[j] = fma(-l, u.x, a[j]);
                        a[j + 1] = fma(-l, u.y, a[j + 1]);
                    }
#pragma unroll
                    for (int k = 0; k < NRHS; ++k) rb[k] = fma(-l, buf[N + k], rb[k]);
                }
            } else if (info == 0) {
                info = i + 1;
            }
        }
    }

    // ---- store the factors in final row order, pivots, info ---------------------------------
    if (sub < m) {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j < n) A[pos + (size_t)j * ld] = a[j];
    }
    if (sub < mn) dipiv[b][sub] = myipiv;
    if (sub == 0) dinfo[b] = info;

    // ---- fused solve: rb holds L^-1 P b; back-substitute with U (divide by the diagonal) -----
    if (NRHS > 0) {
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            if (i < n) {
                const unsigned bq = __ballot_sync(gmask, pos == i);
                const int Q = __ffs(bq) - 1;
#pragma unroll
                for (int k = 0; k < NRHS; ++k) {
                    double x = rb[k] / a[i];
                    x = __shfl_sync(gmask, x, Q);
                    if (pos == i) rb[k] = x;
                    else if (pos < i) rb[k] = fma(-a[i], x, rb[k]);
                }
            }
        }
        if (sub < n) {
#pragma unroll
            for (int k = 0; k < NRHS; ++k) B[pos + (size_t)k * lddb] = rb[k];
        }
    }
}

template <int N, int G, int NRHS>
void launch_one(const Dims &d, double **dA, int **dipiv, int *dinfo, double **dB, int lddb,
                long batch, const int *index_list, cudaStream_t s)
{
    constexpr int GPW = 32 / G;
    const long per_cta = WARPS_PER_CTA * GPW;
    const long grid = (batch + per_cta - 1) / per_cta;
    lu_small_kernel<N, G, NRHS><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(d, dA, dipiv, dinfo, dB,
                                                                            lddb, batch, index_list);
    count_launch();
}

template <int NRHS>
magma_int_t dispatch_n(int K, const Dims &d, double **dA, int **dipiv, int *dinfo, double **dB,
                       int lddb, long batch, const int *il, cudaStream_t s)
{
#define MB200_CASE(NN, GG)                                                         \
    launch_one<NN, GG, NRHS>(d, dA, dipiv, dinfo, dB, lddb, batch, il, s);          \
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
