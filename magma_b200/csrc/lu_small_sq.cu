// Tier S, square fast path: LU (and fused LU + solve, nrhs = 1) of N x N matrices, N = G*R in
// {8, 16, 32}, register resident, with the instruction count per matrix as the design target.
//
// Replaces magmablas/zgetrf_batched_smallsq_noshfl.cu:34-129 and magmablas/zgesv_batched_small.cu:
// 47-162 for the shapes BASELINE.json quotes (n = 16 gesv, n = 32 getrf). At these sizes the
// kernel is bound by warp-instruction issue, not by FP64 or HBM (ncu, profiles/): 2 KB..8 KB of
// traffic per matrix leave ~900 issue slots per matrix and SM sub-partition at 80% of HBM speed.
// So, relative to the generic register kernel (lu_small.cu):
//   * no per-row validity / shape guards: a partially filled warp recomputes the last matrix and
//     only its stores are suppressed;
//   * pivot search on the high words of |x| only (one compare per row, shuffle butterfly or
//     CREDUX, one ballot); a warp-uniform slow path does the exact 64-bit "first maximum" search
//     when two candidates share a high word or a column is all zero;
//   * the rows that must not be updated are predicated off (no zero multipliers, no selects);
//   * the reciprocal of the pivot is computed by every lane from the broadcast pivot value;
//   * the factors leave through a per-matrix shared-memory image in FINAL row order, written
//     with one 8-byte store per element and copied out with 128-bit coalesced stores (full
//     sectors; the generic kernel scatters 8-byte global stores);
//   * the fused solve back-substitutes from that image: lane = final row position, U(q,i) read
//     conflict-free from shared memory, no ownership search per step.
// Arithmetic is the canonical order of oracle/lu_oracle.c, so results are bit-identical to it.
#include "common.cuh"

namespace mb200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned NOPOS = 0xffffffffu;
constexpr int WPC = 4;  // warps per CTA

template <int G>
__device__ __forceinline__ unsigned gmax(unsigned v)
{
    if (G == 32) return __reduce_max_sync(FULL, v);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
        const unsigned w = __shfl_xor_sync(FULL, v, o);
        v = v > w ? v : w;
    }
    return v;
}

template <int G>
__device__ __forceinline__ unsigned gmin(unsigned v)
{
    if (G == 32) return __reduce_min_sync(FULL, v);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
        const unsigned w = __shfl_xor_sync(FULL, v, o);
        v = v < w ? v : w;
    }
    return v;
}

__device__ __forceinline__ unsigned hi_abs(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }

__device__ __forceinline__ double ldg64(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int N, int NRHS>
struct SqSmem {                 // one per matrix in flight
    double stage[N * N];        // factors in final row order, dense column-major
    double row[2][N + 2];       // pivot row ping-pong: [0..N) row, [N] right-hand side, [N+1] 1/pivot
    double y[N];                // right-hand side in final row order
    double dinv[N];             // 1 / u(i,i)
};

// exact search (rare): full 64-bit |x| compare, ties to the smallest current row position.
// Returns bit 0 = this lane holds the pivot row, bit 1 = it is the lane's second row.
template <int G>
__device__ __noinline__ unsigned exact_search(unsigned long long v0, unsigned p0, unsigned long long v1, unsigned p1)
{
    unsigned long long lb = v0;
    unsigned lp = p0;
    const bool t1 = (p1 != NOPOS) && (lp == NOPOS || v1 > lb || (v1 == lb && p1 < lp));
    lb = t1 ? v1 : lb;
    lp = t1 ? p1 : lp;
    const unsigned hi = (lp != NOPOS) ? (unsigned)(lb >> 32) : 0u;
    const unsigned mh = gmax<G>(hi);
    bool c = (lp != NOPOS) && (hi == mh);
    const unsigned lo = c ? (unsigned)lb : 0u;
    const unsigned ml = gmax<G>(lo);
    c = c && (lo == ml);
    const unsigned mp = gmin<G>(c ? lp : NOPOS);
    c = c && (lp == mp);
    return (c ? 1u : 0u) | (t1 ? 2u : 0u);
}

template <int N, int G, int R, int NRHS, bool LDN, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
lu_sq_kernel(double *const *__restrict__ dA, int *const *__restrict__ dipiv, int *__restrict__ dinfo,
             double *const *__restrict__ dB, int ldda, long batch)
{
    static_assert(N == G * R && (R == 1 || R == 2) && NRHS <= 1, "shape");
    constexpr int GPW = 32 / G;
    using S = SqSmem<N, NRHS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int grp = lane / G;
    const int sub = lane % G;
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (grp * G));

    const long slot = ((long)blockIdx.x * WPC + wid) * GPW + grp;
    const bool valid = slot < batch;
    const long b = valid ? slot : batch - 1;  // spare groups redo the last matrix, stores suppressed
    S &sm = reinterpret_cast<S *>(smem_raw)[wid * GPW + grp];

    double *__restrict__ A = dA[b];
    const size_t ld = LDN ? (size_t)N : (size_t)ldda;

    double a[R][N];
    double rb[R];
    unsigned pos[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        pos[r] = (unsigned)(sub + r * G);
#pragma unroll
        for (int j = 0; j < N; ++j) a[r][j] = ldg64(A + (sub + r * G) + (size_t)j * ld);
    }
    double *B = nullptr;
    if (NRHS) {
        B = dB[b];
#pragma unroll
        for (int r = 0; r < R; ++r) rb[r] = ldg64(B + sub + r * G);
    }

    int myipiv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) myipiv[r] = 0;
    unsigned zmask = 0;  // bit i set: column i had an exactly zero pivot
    constexpr int ROWLEN = N + 2;
    const unsigned rowbuf = (unsigned)__cvta_generic_to_shared(&sm.row[0][0]);

#pragma unroll
    for (int i = 0; i < N; ++i) {
        const unsigned buf = rowbuf + (unsigned)((i & 1) * ROWLEN * 8);
        // ---- pivot search: high words first ------------------------------------------------------
        bool act[R];
        unsigned h[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            act[r] = pos[r] >= (unsigned)i;
            h[r] = act[r] ? hi_abs(a[r][i]) : 0u;
        }
        unsigned take1 = 0;  // 0/1 flags kept as integers: bools merged across the cold branch get byte-packed
        unsigned hm = h[0];
        if (R == 2) {
            take1 = h[R - 1] > h[0] ? 1u : 0u;
            hm = h[R - 1] > h[0] ? h[R - 1] : h[0];
        }
        // every lane inverts its own candidate while the search is in flight (off the critical path)
        double rinv = 1.0 / ((R == 2 && take1 != 0) ? a[R - 1][i] : a[0][i]);
        const unsigned mx = gmax<G>(hm);
        unsigned cand = (hm == mx) ? 1u : 0u;

        // ---- optimistic publish: whoever holds the largest high word writes its row -----------------
        auto publish = [&](unsigned c, unsigned t1, double ri) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (c != 0 && (R == 1 || t1 == (unsigned)r)) {
#pragma unroll
                    for (int j = (i & ~1); j < N; j += 2)
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(buf + j * 8), "d"(a[r][j]), "d"(a[r][j + 1]) : "memory");
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(buf + N * 8), "d"(NRHS ? rb[r] : 0.0), "d"(ri) : "memory");
                }
            }
        };
        publish(cand, take1, rinv);
        unsigned bal = __ballot_sync(FULL, cand != 0) & gmask;
        bool unres = __popc(bal) != 1;
        if (R == 2) unres = unres || (cand != 0 && h[0] == h[R - 1]);
        if (__any_sync(FULL, unres)) {
            // two candidates share a high word, or the column is all zero: exact search, publish again
            const unsigned long long m63 = 0x7fffffffffffffffull;
            const unsigned long long v0 = act[0] ? ((unsigned long long)__double_as_longlong(a[0][i]) & m63) : 0ull;
            const unsigned long long v1 = (R == 2 && act[R - 1]) ? ((unsigned long long)__double_as_longlong(a[R - 1][i]) & m63) : 0ull;
            const unsigned res = exact_search<G>(v0, act[0] ? pos[0] : NOPOS, v1, (R == 2 && act[R - 1]) ? pos[R - 1] : NOPOS);
            cand = res & 1u;
            take1 = res >> 1;
            bal = __ballot_sync(FULL, cand != 0) & gmask;
            rinv = 1.0 / ((R == 2 && take1 != 0) ? a[R - 1][i] : a[0][i]);
            __syncwarp();
            publish(cand, take1, rinv);
        }
        __syncwarp();

        // ---- broadcast loads, all in flight together ---------------------------------------------------
        double u[N + 2];
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u[N]), "=d"(u[N + 1]) : "r"(buf + N * 8) : "memory");
#pragma unroll
        for (int j = (i & ~1); j < N; j += 2)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u[j]), "=d"(u[j + 1]) : "r"(buf + j * 8) : "memory");

        // ---- bookkeeping in the shadow of the loads: pivot index, row positions ---------------------------
        const int P = 31 - __clz((int)bal);  // the (single) pivot lane of this group
        const unsigned lpos = (R == 2 && take1 != 0) ? pos[R - 1] : pos[0];
        const unsigned p = __shfl_sync(FULL, lpos, P);  // current position of the pivot row
        if (sub == (i % G)) myipiv[i / G] = (int)p + 1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (pos[r] == (unsigned)i) pos[r] = p;
            if (cand != 0 && (R == 1 || take1 == (unsigned)r)) pos[r] = (unsigned)i;
        }

        const double rr = u[N + 1];
        const bool nz = (u[i] != 0.0);
        if (!nz) zmask |= (1u << i);
        if (NRHS && sub == 0) sm.dinv[i] = rr;

        // ---- eliminate -----------------------------------------------------------------------------
        // Rows that are not updated (already pivoted, or a singular column) use l = 0: fma(-0, u, a)
        // returns a, so the update itself needs neither a branch nor a select per element. (A
        // predicated update -- `if (upd) a = fma(..)` or `@p fma` in PTX -- is turned by ptxas into
        // an unconditional DFMA plus two FSELs per element.)
        double l[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool upd = nz && pos[r] > (unsigned)i;
            l[r] = upd ? a[r][i] * rr : 0.0;
            if (upd) a[r][i] = l[r];
        }
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) a[r][j] = fma(-l[r], u[j], a[r][j]);
        }
        if (NRHS) {
#pragma unroll
            for (int r = 0; r < R; ++r) rb[r] = fma(-l[r], u[N], rb[r]);
        }
    }

    // ---- factors -> shared image in final row order -> coalesced 128-bit stores ------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int j = 0; j < N; ++j) sm.stage[pos[r] + j * N] = a[r][j];
        if (NRHS) sm.y[pos[r]] = rb[r];
    }
    __syncwarp();
    if (valid) {
        const bool al16 = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ld & 1) == 0);
        constexpr int CH = (N * N / 2) / G;  // 16-byte chunks per lane
        if (al16) {
#pragma unroll
            for (int t = 0; t < CH; ++t) {
                const int e = 2 * (t * G + sub);  // even element index in the dense image
                const int col = e / N, row = e % N;
                const double2 v2 = *reinterpret_cast<const double2 *>(&sm.stage[e]);
                *reinterpret_cast<double2 *>(A + row + (size_t)col * ld) = v2;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 2 * CH; ++t) {
                const int e = t * G + sub;
                const int col = e / N, row = e % N;
                A[row + (size_t)col * ld] = sm.stage[e];
            }
        }
        int *ip = dipiv[b];
#pragma unroll
        for (int r = 0; r < R; ++r) ip[sub + r * G] = myipiv[r];
        if (sub == 0) dinfo[b] = zmask ? __ffs((int)zmask) : 0;
    }

    // ---- fused solve: back substitution on the shared image, lane = final row position ----------
    if (NRHS) {
        double y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = sm.y[sub + r * G];
        const int gbase = lane & ~(G - 1);
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            const int ri = i / G, si = i % G;
            if (sub == si) y[ri] = y[ri] * sm.dinv[i];
            const double x = __shfl_sync(FULL, y[ri], gbase + si);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r * G < i) {  // compile time: some row of this slot lies above i
                    if (sub + r * G < i) y[r] = fma(-sm.stage[(sub + r * G) + i * N], x, y[r]);
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < R; ++r) B[sub + r * G] = y[r];
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Shuffle variant for G < 32 (several matrices per warp, two rows per lane): the pivot row goes
// from the pivot lane to its group by SHFL.IDX instead of through shared memory. With one matrix
// per 8 lanes a shared-memory publish moves 16 useful bytes per LSU wavefront (one active lane per
// quarter warp) and the broadcast load another wavefront pair; ncu showed that kernel bound by
// l1tex__data_pipe_lsu_wavefronts (75% of peak), not by issue slots or DRAM. A shuffle moves the
// same value to four matrices' groups in two wavefronts and needs no store at all. Factors are
// stored straight from registers (8-byte stores at the final row positions), the fused solve
// back-substitutes in registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl64(double v, int src)
{
    const int lo = __shfl_sync(FULL, __double2loint(v), src);
    const int hi = __shfl_sync(FULL, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

// reciprocal for the cold path: one out-of-line copy instead of a second inlined Newton sequence per step
__device__ __noinline__ double rcp_cold(double x) { return 1.0 / x; }

// x1 if flag else x0, as a select on registers (written as `flag ? a[1][j] : a[0][j]` nvcc turns the
// row array into a dynamically indexed local-memory array)
__device__ __forceinline__ double sel64(unsigned flag, double x1, double x0)
{
    double d;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.f64 %0, %1, %2, q;\n\t}" : "=d"(d) : "d"(x1), "d"(x0), "r"(flag));
    return d;
}

template <int I, int N, typename F>
__device__ __forceinline__ void sq_static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        sq_static_for<I + 1, N>(f);
    }
}

template <int N>
struct SqsSmem {           // one per matrix in flight (fused-solve variant only)
    double stage[N * N];   // factors in final row order, dense column-major
    double y[N];           // right-hand side in final row order
    double dinv[N];        // 1 / u(i,i)
    double pad[8];         // group stride = 16 banks mod 32: the four groups of a warp read their 64-byte
                           // column pieces in two wavefronts instead of four (ncu: 60 excess LDS wavefronts/pass)
};

// WC warps per CTA; LOCK: one CTA-wide barrier per column step keeps the warps of a CTA on the same
// instructions. The unrolled body is ~50 KB of SASS, more than the SM's instruction cache: with the
// warps drifting apart ncu showed the GPC instruction cache at 87% of its request peak and the
// issue rate capped near 45%; in lockstep every fetched line serves all warps of the CTA.
template <int N, int G, int NRHS, bool LDN, int MINB, int WC, bool LOCK>
__global__ void __launch_bounds__(WC * 32, MINB)
lu_sqs_kernel(double *const *__restrict__ dA, int *const *__restrict__ dipiv, int *__restrict__ dinfo,
              double *const *__restrict__ dB, int ldda, long batch)
{
    constexpr int R = 2;
    static_assert(N == G * R && G < 32 && NRHS <= 1, "shape");
    constexpr int GPW = 32 / G;

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int grp = lane / G;
    const int sub = lane % G;
    const unsigned gmask = ((1u << G) - 1u) << (grp * G);

    // (A persistent-warp variant that fetched the next matrices' pointers and pulled their lines into
    // L2 one pass ahead removed the load prologue from the stall profile but was not faster: the
    // kernel is bound by the LSU/shuffle pipe, not by load latency. profiles/README.md, round 1.)
    const size_t ld = LDN ? (size_t)N : (size_t)ldda;
    const long slot = ((long)blockIdx.x * WC + wid) * GPW + grp;
    const bool valid = slot < batch;
    const long b = valid ? slot : batch - 1;  // spare groups redo the last matrix, stores suppressed
    double *__restrict__ A = dA[b];
    double *Bcur = NRHS ? dB[b] : nullptr;
    int *const ip = dipiv[b];  // fetched up front: the pivot store at the end must not wait on a pointer load

    double a[R][N];
    double rb[R];
    double mydinv[R];
    unsigned pos[R];
    int myipiv[R];
    // A lane starts with the ADJACENT rows 2 sub and 2 sub + 1: one 16-byte load per column brings both, and a
    // group's eight lanes cover a whole 128-byte column of the matrix per instruction (with rows sub and sub + G a
    // warp load touched four half-used lines: 4 tag wavefronts per 8-byte instruction, 136 per pass instead of 68).
    // Which rows a lane starts with is immaterial afterwards: everything below works on positions.
    // (n = 8 keeps rows sub and sub + G: the 64-register budget of that variant does not survive the extra path.)
    constexpr bool ADJ = (N >= 16);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        pos[r] = (unsigned)(ADJ ? 2 * sub + r : sub + r * G);
        myipiv[r] = 0;
        mydinv[r] = 0.0;
        rb[r] = 0.0;
    }
    double *B = nullptr;
    if (NRHS) B = Bcur;
    const bool v128 = ADJ &&
                      __all_sync(FULL, ((reinterpret_cast<uintptr_t>(A) | (NRHS ? reinterpret_cast<uintptr_t>(Bcur) : (uintptr_t)0)) & 15) == 0) &&
                      (LDN || (ldda & 1) == 0);
    if (v128) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < N; ++j)
            asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(a[0][j]), "=d"(a[1][j]) : "l"(A + 2 * sub + (size_t)j * ld));
        if (NRHS) asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(rb[0]), "=d"(rb[1]) : "l"(B + 2 * sub));
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < N; ++j) a[r][j] = ldg64(A + pos[r] + (size_t)j * ld);
            if (NRHS) rb[r] = ldg64(B + pos[r]);
        }
    }
    unsigned zmask = 0;  // bit i set: column i had an exactly zero pivot

    // both loops expanded at compile time: a `#pragma unroll` nest of the n = 32 size is left rolled by nvcc, with a[][]
    // in local memory
    sq_static_for<0, N>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        if (LOCK) __syncthreads();
        // ---- pivot search: high words first ------------------------------------------------------
        bool act[R];
        unsigned h[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            act[r] = pos[r] >= (unsigned)i;
            h[r] = act[r] ? hi_abs(a[r][i]) : 0u;
        }
        unsigned take1 = h[1] > h[0] ? 1u : 0u;
        const unsigned hm = h[1] > h[0] ? h[1] : h[0];
        double cv = sel64(take1, a[1][i], a[0][i]);  // this lane's candidate
        double rtrue = 1.0 / cv;                     // inverted while the search is in flight
        // fused solve: a zero pivot travels as a zero "reciprocal", which saves the broadcast of the pivot value
        // (measured: gesv 1.265 -> 1.236 ms; the factor-only kernel is better off with the extra shuffle)
        double rinv = (!NRHS || cv != 0.0) ? rtrue : 0.0;
        const unsigned mx = gmax<G>(hm);
        unsigned cand = (hm == mx) ? 1u : 0u;
        unsigned bal = __ballot_sync(FULL, cand != 0) & gmask;
        bool unres = __popc(bal) != 1;
        unres = unres || (cand != 0 && h[0] == h[1]);
        if (__any_sync(FULL, unres)) {
            // two candidates share a high word, or the column is all zero: exact search
            const unsigned long long m63 = 0x7fffffffffffffffull;
            const unsigned long long v0 = act[0] ? ((unsigned long long)__double_as_longlong(a[0][i]) & m63) : 0ull;
            const unsigned long long v1 = act[1] ? ((unsigned long long)__double_as_longlong(a[1][i]) & m63) : 0ull;
            const unsigned res = exact_search<G>(v0, act[0] ? pos[0] : NOPOS, v1, act[1] ? pos[1] : NOPOS);
            cand = res & 1u;
            take1 = res >> 1;
            bal = __ballot_sync(FULL, cand != 0) & gmask;
            cv = sel64(take1, a[1][i], a[0][i]);
            rtrue = NRHS ? rcp_cold(cv) : 1.0 / cv;
            rinv = (!NRHS || cv != 0.0) ? rtrue : 0.0;
        }
        const int P = 31 - __clz((int)bal);  // the (single) pivot lane of this group

        // ---- broadcasts from the pivot lane ------------------------------------------------------------
        const double rr = shfl64(rinv, P);
        const double piv = NRHS ? rr : shfl64(cv, P);  // only its zero-ness matters
        const unsigned p = __shfl_sync(FULL, take1 ? pos[1] : pos[0], P);  // current position of the pivot row
        if (sub == (i % G)) myipiv[i / G] = (int)p + 1;
        const bool nz = (piv != 0.0);
        if (!nz) zmask |= (1u << i);
        double l[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool pv = (cand != 0) && (take1 == (unsigned)r);
            if (pos[r] == (unsigned)i) pos[r] = p;
            if (pv) pos[r] = (unsigned)i;
            if (NRHS && pv) mydinv[r] = rtrue;  // the solve divides by the pivot whatever it is, like the oracle
            // rows that are not updated use l = 0 (see lu_sq_kernel)
            const bool upd = nz && pos[r] > (unsigned)i;
            l[r] = upd ? a[r][i] * rr : 0.0;
            if (upd) a[r][i] = l[r];
        }
        sq_static_for<i + 1, N>([&](auto jc_) {
            constexpr int j = decltype(jc_)::value;
            const double u = shfl64(sel64(take1, a[1][j], a[0][j]), P);
#pragma unroll
            for (int r = 0; r < R; ++r) a[r][j] = fma(-l[r], u, a[r][j]);
        });
        if (NRHS) {
            const double ub = shfl64(sel64(take1, rb[1], rb[0]), P);
#pragma unroll
            for (int r = 0; r < R; ++r) rb[r] = fma(-l[r], ub, rb[r]);
        }
    });

    if (!NRHS) {
        // ---- factor only: straight from registers, 8-byte stores at the final row positions --------------
        if (valid) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < N; ++j) A[pos[r] + (size_t)j * ld] = a[r][j];
            }
#pragma unroll
            for (int r = 0; r < R; ++r) ip[sub + r * G] = myipiv[r];
            if (sub == 0) dinfo[b] = zmask ? __ffs((int)zmask) : 0;
        }
    } else {
        // ---- fused solve: factors -> shared image in final row order (coalesced 128-bit copy-out), then
        //      back substitution on the image with lane = final row position (no ownership search) ----------
        extern __shared__ __align__(16) unsigned char smem_raw[];
        SqsSmem<N> &sm = reinterpret_cast<SqsSmem<N> *>(smem_raw)[wid * GPW + grp];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < N; ++j) sm.stage[pos[r] + j * N] = a[r][j];
            sm.y[pos[r]] = rb[r];
            // the row at final position q pivoted at step q: its lane kept 1/u(q,q)
            sm.dinv[pos[r]] = mydinv[r];
        }
        __syncwarp();
        if (valid) {
            const bool al16 = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ld & 1) == 0);
            constexpr int CH = (N * N / 2) / G;  // 16-byte chunks per lane
            if (al16) {
#pragma unroll
                for (int t = 0; t < CH; ++t) {
                    const int e = 2 * (t * G + sub);
                    const int col = e / N, row = e % N;
                    *reinterpret_cast<double2 *>(A + row + (size_t)col * ld) = *reinterpret_cast<const double2 *>(&sm.stage[e]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 2 * CH; ++t) {
                    const int e = t * G + sub;
                    const int col = e / N, row = e % N;
                    A[row + (size_t)col * ld] = sm.stage[e];
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) ip[sub + r * G] = myipiv[r];
            if (sub == 0) dinfo[b] = zmask ? __ffs((int)zmask) : 0;
        }
        double y[R], dv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            y[r] = sm.y[sub + r * G];
            dv[r] = sm.dinv[sub + r * G];
        }
        const int gbase = lane & ~(G - 1);
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            const int ri = i / G, si = i % G;
            if (sub == si) y[ri] = y[ri] * dv[ri];
            const double x = shfl64(y[ri], gbase + si);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r * G < i) {  // compile time: some row of this slot lies above i
                    const double m = (sub + r * G < i) ? sm.stage[(sub + r * G) + i * N] : 0.0;
                    y[r] = fma(-m, x, y[r]);
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < R; ++r) B[sub + r * G] = y[r];
        }
    }
}

template <int N, int G, int NRHS, int MINB, int WC, bool LOCK>
magma_int_t launch_sqs(double **dA, int ldda, int **dipiv, int *dinfo, double **dB, long batch, cudaStream_t s)
{
    constexpr int GPW = 32 / G;
    const long per_cta = WC * GPW;
    const long grid = (batch + per_cta - 1) / per_cta;
    const size_t smem = NRHS ? sizeof(SqsSmem<N>) * per_cta : 0;
    static DevOnce once_a, once_b;
    smem_optin(once_a, lu_sqs_kernel<N, G, NRHS, true, MINB, WC, LOCK>, smem);
    smem_optin(once_b, lu_sqs_kernel<N, G, NRHS, false, MINB, WC, LOCK>, smem);
    if (ldda == N)
        lu_sqs_kernel<N, G, NRHS, true, MINB, WC, LOCK><<<(unsigned)grid, WC * 32, smem, s>>>(dA, dipiv, dinfo, dB, ldda, batch);
    else
        lu_sqs_kernel<N, G, NRHS, false, MINB, WC, LOCK><<<(unsigned)grid, WC * 32, smem, s>>>(dA, dipiv, dinfo, dB, ldda, batch);
    count_launch();
    MB200_CHECK_LAUNCH("lu_sqs_kernel");
    return 0;
}

template <int N, int G, int R, int NRHS, int MINB>
magma_int_t launch_sq(double **dA, int ldda, int **dipiv, int *dinfo, double **dB, long batch, cudaStream_t s)
{
    constexpr int GPW = 32 / G;
    const long per_cta = WPC * GPW;
    const long grid = (batch + per_cta - 1) / per_cta;
    const size_t smem = sizeof(SqSmem<N, NRHS>) * per_cta;
    if (ldda == N) {
        auto k = lu_sq_kernel<N, G, R, NRHS, true, MINB>;
        static DevOnce once;
        smem_optin(once, k, smem);
        k<<<(unsigned)grid, WPC * 32, smem, s>>>(dA, dipiv, dinfo, dB, ldda, batch);
    } else {
        auto k = lu_sq_kernel<N, G, R, NRHS, false, MINB>;
        static DevOnce once;
        smem_optin(once, k, smem);
        k<<<(unsigned)grid, WPC * 32, smem, s>>>(dA, dipiv, dinfo, dB, ldda, batch);
    }
    count_launch();
    MB200_CHECK_LAUNCH("lu_sq_kernel");
    return 0;
}

}  // namespace

// Square fast path. Returns -100 when the shape is not covered (caller uses the generic kernel).
magma_int_t lu_sq_launch(int n, double **dA, int ldda, int **dipiv, int *dinfo, int nrhs, double **dB, int lddb,
                         long batch, cudaStream_t s)
{
    (void)lddb;  // nrhs == 1: a single column, lddb is irrelevant
    if (batch <= 0 || nrhs > 1) return -100;
    if (nrhs == 0) {
        switch (n) {
            case 8: return launch_sqs<8, 4, 0, 8, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
            case 16:
                if (g_small_rows == 4) return launch_sqs<16, 8, 0, 4, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
                return launch_sqs<16, 8, 0, 5, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
            case 32:
                if (g_small_rows == 5) return launch_sqs<32, 16, 0, 3, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
                if (g_small_rows == 6) return launch_sqs<32, 16, 0, 2, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
                return g_small_rows == 3 ? launch_sq<32, 32, 1, 0, 4>(dA, ldda, dipiv, dinfo, dB, batch, s) : -100;
            default: return -100;
        }
    }
    switch (n) {
        case 8: return launch_sqs<8, 4, 1, 8, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
        case 16:
            if (g_small_rows == 4) return launch_sqs<16, 8, 1, 4, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
            return launch_sqs<16, 8, 1, 5, 4, false>(dA, ldda, dipiv, dinfo, dB, batch, s);
        case 32: return g_small_rows == 3 ? launch_sq<32, 32, 1, 1, 4>(dA, ldda, dipiv, dinfo, dB, batch, s) : -100;
        default: return -100;
    }
}

}  // namespace mb200
