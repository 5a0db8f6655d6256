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
#include "lu_common.cuh"
#include <type_traits>

namespace mb200 {

namespace {

constexpr int FT = 256;    // threads per CTA
constexpr int FLD = 130;   // leading dimension of the left part: C fragments (col 2q, row g) conflict-free (130 = 2 mod 16)
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

__device__ __forceinline__ void f_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- mbarrier / TMA (1-D bulk copies) ----------------------------------------------------------------------
__device__ __forceinline__ unsigned f_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void f_mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void f_mbar_expect(void *bar, unsigned bytes)  // one arrival + the bytes the copies will deliver
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(f_saddr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void f_mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "FWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@!p bra FWAIT_%=;\n\t}" ::"r"(f_saddr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void f_bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(f_saddr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(f_saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void f_bulk_prefetch_l2(const void *gsrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_bulk_store(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(f_saddr(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void f_bulk_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void f_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- named barriers (ids 1..3; 0 is __syncthreads) ------------------------------------------------------------
constexpr int BAR_PANEL = 1;  // chain warp arrives: sub-panel published; update warps wait
constexpr int BAR_NEXT = 2;   // update warps arrive: the next sub-panel's columns are up to date; chain warp waits
constexpr int BAR_UPD = 3;    // update warps only: block row solved, -U in place
constexpr int BAR_MOVED = 4;  // update warps only: the sub-panel's row permutation has been applied everywhere
__device__ __forceinline__ void f_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void f_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ double f_sel(bool p, double a, double b)  // p ? a : b as two SELs (never a branch)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((unsigned)p));
    return r;
}
// shared-space accesses by 32-bit address (the generic-pointer forms made ptxas rebuild the window base in the chain)
__device__ __forceinline__ double f_lds(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void f_sts_if(bool p, unsigned a, double v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.f64 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void f_sts32_if(bool p, unsigned a, int v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void f_sts8_if(bool p, unsigned a, int v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u8 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "r"(v) : "memory");
}

// Exact pivot choice for the cases the fast path hands over (several rows share the largest high word, or the
// column's high words are all zero): first maximum of |x| over every active row of the warp, ties to the smaller
// logical position. Returns winner lane | slot << 8. Out of line: it runs on structured inputs only.
__device__ __noinline__ int chain_slow_pick(double v0, double v1, double v2, double v3, int p0, int p1, int p2, int p3,
                                            unsigned alive)
{
    unsigned long long lb = 0;
    int lp = NOPOS_I, lk = 0;
    const double v[4] = {v0, v1, v2, v3};
    const int p[4] = {p0, p1, p2, p3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v[k]) & 0x7fffffffffffffffull;
        if (((alive >> k) & 1u) && (b > lb || lp == NOPOS_I || (b == lb && p[k] < lp))) {
            lb = b;
            lp = p[k];
            lk = k;
        }
    }
    const int wl = warp_argmax_lane(lb, lp);
    const int K = __shfl_sync(0xffffffffu, lk, wl);
    return wl | (K << 8);
}

// ---- sub-panel: jb <= 8 columns at (j, j) of view V, rows j..mv-1, one warp, NA register rows per lane ----------
// Lane holds view rows lane + 32*(k0 + k), k < NA, for the whole sub-panel. Nothing moves physically here: the
// shared-memory image of the eight columns stays in the row order the sub-panel started with, and the net permutation
// (S.mdst/msrc, <= 16 moves) is applied afterwards to EVERY column of the view, these eight included, by the update
// warps. Per column: high words of |x| of the live rows -> REDUX.MAX; the rows that match are counted per slot with
// one REDUX.SUM, so winner lane AND slot are warp-uniform and the pivot row is shuffled out of statically named
// registers by one of NA tiny code bodies (per-lane select chains became divergent branch trees in ptxas: 60% of the
// first version's time). The winner's multipliers so far are row i of L11 (S.L11, read by the block-row solve); lane 0
// stores the pivot row's U part; every live row stores its new multiplier. A pivot row's registers are dead from then
// on, so the rank-1 update runs on every slot unconditionally (dead and padding slots compute garbage nobody reads).
template <int NA, int LD, typename SM>
__device__ __forceinline__ void panel_chain(SM &S, const unsigned vbase, const int j, const int jb, const int mv,
                                            const int k0, const int lane, const int row_off)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned l11 = f_saddr(S.L11), sip = f_saddr(S.ipiv), smd = f_saddr(S.mdst), sms = f_saddr(S.msrc);
    const unsigned cbase = vbase + (unsigned)(j * LD) * 8u;  // column j of the view
    double a[NA][8];
    int pos[NA];          // logical position of the row held in slot k
    unsigned alive = 0;   // bit k: slot k holds a row that has not been taken as a pivot yet
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        const int r = lane + 32 * (k0 + k);
        const bool valid = (r >= j) && (r < mv);
        pos[k] = r;
        alive |= valid ? (1u << k) : 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[k][c] = (valid && c < jb) ? f_lds(cbase + (unsigned)(c * LD + r) * 8u) : 0.0;
    }
    const bool lane0 = lane == 0;
    int info = 0;
    int cnt = 0;  // moves recorded so far (warp-uniform)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < jb) {  // warp-uniform
            const int ji = j + i;
            unsigned hv[NA];
            unsigned lmx = 0;
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                hv[k] = ((alive >> k) & 1u) ? ((unsigned)__double2hiint(a[k][i]) & 0x7fffffffu) : 0u;
                lmx = hv[k] > lmx ? hv[k] : lmx;
            }
            const unsigned mx = __reduce_max_sync(FULL, lmx);
            unsigned code = 0;
            double cv = 1.0;
            int lp = 0;
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                const bool hit = hv[k] == mx;
                code += hit ? (1u << (8 * k)) : 0u;
                cv = f_sel(hit, a[k][i], cv);
                lp = hit ? pos[k] : lp;
            }
            const double rinv = rcp_fast_f64(cv);  // this lane's candidate reciprocal, in flight while the vote runs
            asm volatile("" ::"d"(rinv));        // keep it ahead of the vote (ptxas sank it behind the slot switch)
            const unsigned tot = __reduce_add_sync(FULL, code);
            int wl, K;
            const bool unique = (mx != 0u) && ((tot & 0xfefefefeu) == 0u) && (__popc(tot) == 1);
            if (unique) {
                K = (__ffs(tot) - 1) >> 3;
                wl = __ffs(__ballot_sync(FULL, code != 0u)) - 1;
            } else {
                const int pk = chain_slow_pick(a[0][i], NA > 1 ? a[NA > 1 ? 1 : 0][i] : 0.0, NA > 2 ? a[NA > 2 ? 2 : 0][i] : 0.0,
                                               NA > 3 ? a[NA > 3 ? 3 : 0][i] : 0.0, pos[0], NA > 1 ? pos[NA > 1 ? 1 : 0] : 0,
                                               NA > 2 ? pos[NA > 2 ? 2 : 0] : 0, NA > 3 ? pos[NA > 3 ? 3 : 0] : 0, alive);
                wl = pk & 0xff;
                K = pk >> 8;
                lp = pos[0];
#pragma unroll
                for (int k = 1; k < NA; ++k) lp = (K == k) ? pos[k] : lp;
            }
            const int wp = __shfl_sync(FULL, lp, wl);  // logical position of the pivot row (>= ji)
            const bool me = lane == wl;
#pragma unroll
            for (int k = 0; k < NA; ++k) pos[k] = (pos[k] == ji) ? wp : pos[k];
            double u[8];
            auto take = [&](auto Kc) {
                constexpr int KK = decltype(Kc)::value;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c >= i) u[c] = __shfl_sync(FULL, a[KK][c], wl);
                    else f_sts_if(me, l11 + (unsigned)(i * 8 + c) * 8u, a[KK][c]);
                }
                pos[KK] = me ? ji : pos[KK];
                alive = me ? (alive & ~(1u << KK)) : alive;
            };
            if (NA == 1 || K == 0) take(std::integral_constant<int, 0>{});
            else if (NA == 2 || K == 1) take(std::integral_constant<int, (NA > 1 ? 1 : 0)>{});
            else if (NA == 3 || K == 2) take(std::integral_constant<int, (NA > 2 ? 2 : 0)>{});
            else take(std::integral_constant<int, (NA > 3 ? 3 : 0)>{});
            double rv;
            if (unique && mx >= 0x01800000u && mx < 0x7e000000u) {  // warp-uniform: |pivot| in [2^-999, 2^993)
                rv = __shfl_sync(FULL, rinv, wl);
            } else {
                rv = 1.0 / u[i];
            }
            const int prow = wl + 32 * (k0 + K);  // where the pivot row sits in the (unpermuted) image
            // lane 0: the pivot row's U part into the image, pivot index, and the move of this row if it is one
            const unsigned prow_a = cbase + (unsigned)prow * 8u;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c >= i && c < jb) f_sts_if(lane0, prow_a + (unsigned)(c * LD) * 8u, u[c]);
            f_sts32_if(lane0, sip + (unsigned)(row_off + ji) * 4u, row_off + wp);
            f_sts8_if(lane0 && prow != ji, smd + (unsigned)cnt, ji);
            f_sts8_if(lane0 && prow != ji, sms + (unsigned)cnt, prow);
            cnt += (prow != ji) ? 1 : 0;
            if (u[i] != 0.0) {  // warp-uniform; a zero pivot leaves the column unscaled and skips the update (oracle_dgetf2)
#pragma unroll
                for (int k = 0; k < NA; ++k) {
                    const double l = a[k][i] * rv;
                    a[k][i] = l;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c > i) a[k][c] = fma(-l, u[c], a[k][c]);
                }
            } else if (info == 0) {
                info = row_off + ji + 1;
            }
            // column i of every live row is final now
#pragma unroll
            for (int k = 0; k < NA; ++k)
                f_sts_if((alive >> k) & 1u, cbase + (unsigned)(i * LD + lane + 32 * (k0 + k)) * 8u, a[k][i]);
        }
    }
    // rows that were never a pivot but were displaced join the list
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        const int r = lane + 32 * (k0 + k);
        const bool moved = ((alive >> k) & 1u) && pos[k] != r;
        const unsigned mask = __ballot_sync(FULL, moved);
        if (moved) {
            const int idx = cnt + __popc(mask & ((1u << lane) - 1u));
            S.mdst[idx] = (unsigned char)pos[k];
            S.msrc[idx] = (unsigned char)r;
        }
        cnt += __popc(mask);
    }
    if (lane == 0) {
        S.nmoves = cnt;
        if (info != 0 && S.info == 0) S.info = info;
    }
}

// ---- the same sub-panel chain with a ROLLED column loop ----------------------------------------------------------------
// The unrolled version above is ~1000 straight-line instructions per sub-panel, executed once: ncu charged 40% of the chain
// warp's time to instruction fetch (`no_instruction`). Here the current column is always register column 0: the rank-1
// update writes column c into column c-1 (`a[k][c-1] = fma(-l, u[c], a[k][c])` -- the shift is free), so one loop body of
// ~170 instructions serves every column and stays in the instruction cache. Columns past the live window hold garbage that
// only ever shifts towards dead columns. The multipliers of the pivot rows (the 8x8 unit-lower block the block-row solve
// needs) are read from the image after the permutation has been applied, so S.L11 is not used on this path.
template <int NA, int LD, typename SM>
__device__ __forceinline__ void panel_chain_rolled(SM &S, const unsigned vbase, const int j, const int jb, const int mv,
                                                   const int k0, const int lane, const int row_off)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned sip = f_saddr(S.ipiv), smd = f_saddr(S.mdst), sms = f_saddr(S.msrc);
    unsigned cbase = vbase + (unsigned)(j * LD) * 8u;  // column j + i of the view (advances with i)
    double a[NA][8];
    int pos[NA];
    unsigned alive = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        const int r = lane + 32 * (k0 + k);
        const bool valid = (r >= j) && (r < mv);
        pos[k] = r;
        alive |= valid ? (1u << k) : 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[k][c] = (valid && c < jb) ? f_lds(cbase + (unsigned)(c * LD + r) * 8u) : 0.0;
    }
    const bool lane0 = lane == 0;
    const unsigned rowb = (unsigned)(lane + 32 * k0) * 8u;  // byte offset of this lane's slot-0 row inside a column
    int info = 0;
    int cnt = 0;
    double u[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int i = 0; i < jb; ++i) {
        const int ji = j + i;
        const int nlive = jb - i;  // live register columns: 0 .. nlive-1
        unsigned hv[NA];
        unsigned lmx = 0;
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            hv[k] = ((alive >> k) & 1u) ? ((unsigned)__double2hiint(a[k][0]) & 0x7fffffffu) : 0u;
            lmx = hv[k] > lmx ? hv[k] : lmx;
        }
        const unsigned mx = __reduce_max_sync(FULL, lmx);
        unsigned code = 0;
        double cv = 1.0;
        int lp = 0;
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            const bool hit = hv[k] == mx;
            code += hit ? (1u << (8 * k)) : 0u;
            cv = f_sel(hit, a[k][0], cv);
            lp = hit ? pos[k] : lp;
        }
        const double rinv = rcp_fast_f64(cv);
        asm volatile("" ::"d"(rinv));
        const unsigned tot = __reduce_add_sync(FULL, code);
        int wl, K;
        const bool unique = (mx != 0u) && ((tot & 0xfefefefeu) == 0u) && (__popc(tot) == 1);
        if (unique) {
            K = (__ffs(tot) - 1) >> 3;
            wl = __ffs(__ballot_sync(FULL, code != 0u)) - 1;
        } else {
            const int pk = chain_slow_pick(a[0][0], NA > 1 ? a[NA > 1 ? 1 : 0][0] : 0.0, NA > 2 ? a[NA > 2 ? 2 : 0][0] : 0.0,
                                           NA > 3 ? a[NA > 3 ? 3 : 0][0] : 0.0, pos[0], NA > 1 ? pos[NA > 1 ? 1 : 0] : 0,
                                           NA > 2 ? pos[NA > 2 ? 2 : 0] : 0, NA > 3 ? pos[NA > 3 ? 3 : 0] : 0, alive);
            wl = pk & 0xff;
            K = pk >> 8;
            lp = pos[0];
#pragma unroll
            for (int k = 1; k < NA; ++k) lp = (K == k) ? pos[k] : lp;
        }
        const int wp = __shfl_sync(FULL, lp, wl);
        const bool me = lane == wl;
#pragma unroll
        for (int k = 0; k < NA; ++k) pos[k] = (pos[k] == ji) ? wp : pos[k];
        auto take = [&](auto Kc) {
            constexpr int KK = decltype(Kc)::value;
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] = __shfl_sync(FULL, a[KK][c], wl);
            if (nlive > 4) {  // warp-uniform: the second half of the window is dead from column 4 on
#pragma unroll
                for (int c = 4; c < 8; ++c) u[c] = __shfl_sync(FULL, a[KK][c], wl);
            }
            pos[KK] = me ? ji : pos[KK];
            alive = me ? (alive & ~(1u << KK)) : alive;
        };
        if (NA == 1 || K == 0) take(std::integral_constant<int, 0>{});
        else if (NA == 2 || K == 1) take(std::integral_constant<int, (NA > 1 ? 1 : 0)>{});
        else if (NA == 3 || K == 2) take(std::integral_constant<int, (NA > 2 ? 2 : 0)>{});
        else take(std::integral_constant<int, (NA > 3 ? 3 : 0)>{});
        double rv;
        if (unique && mx >= 0x01800000u && mx < 0x7e000000u) {
            rv = __shfl_sync(FULL, rinv, wl);
        } else {
            rv = 1.0 / u[0];
        }
        const int prow = wl + 32 * (k0 + K);
        // lane 0: the pivot row's U part into the image (its physical row), pivot index, its move
        const unsigned prow_a = cbase + (unsigned)prow * 8u;
#pragma unroll
        for (int c = 0; c < 8; ++c) f_sts_if(lane0 && c < nlive, prow_a + (unsigned)(c * LD) * 8u, u[c]);
        f_sts32_if(lane0, sip + (unsigned)(row_off + ji) * 4u, row_off + wp);
        f_sts8_if(lane0 && prow != ji, smd + (unsigned)cnt, ji);
        f_sts8_if(lane0 && prow != ji, sms + (unsigned)cnt, prow);
        cnt += (prow != ji) ? 1 : 0;
        if (u[0] != 0.0) {  // warp-uniform
            double l[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                l[k] = a[k][0] * rv;
                f_sts_if((alive >> k) & 1u, cbase + rowb + (unsigned)(32 * k) * 8u, l[k]);  // column ji of a live row is final
#pragma unroll
                for (int c = 1; c < 5; ++c) a[k][c - 1] = fma(-l[k], u[c], a[k][c]);
            }
            if (nlive > 5) {  // warp-uniform: columns 5..7 are live only in the first three steps
#pragma unroll
                for (int k = 0; k < NA; ++k)
#pragma unroll
                    for (int c = 5; c < 8; ++c) a[k][c - 1] = fma(-l[k], u[c], a[k][c]);
            }
        } else {
            if (info == 0) info = row_off + ji + 1;
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                f_sts_if((alive >> k) & 1u, cbase + rowb + (unsigned)(32 * k) * 8u, a[k][0]);  // unscaled, as the oracle leaves it
#pragma unroll
                for (int c = 1; c < 8; ++c) a[k][c - 1] = a[k][c];
            }
        }
        cbase += (unsigned)LD * 8u;
    }
    // rows that were never a pivot but were displaced join the list
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        const int r = lane + 32 * (k0 + k);
        const bool moved = ((alive >> k) & 1u) && pos[k] != r;
        const unsigned mask = __ballot_sync(FULL, moved);
        if (moved) {
            const int idx = cnt + __popc(mask & ((1u << lane) - 1u));
            S.mdst[idx] = (unsigned char)pos[k];
            S.msrc[idx] = (unsigned char)r;
        }
        cnt += __popc(mask);
    }
    if (lane == 0) {
        S.nmoves = cnt;
        if (info != 0 && S.info == 0) S.info = info;
    }
}

// ---- LU of a resident view: mv x nv at V (leading dimension LD), view row 0 = global row/step row_off -----------
// X (xcols columns, leading dimension FLD, rows aligned with the view's) receives the interchanges only: the L21
// block of the left part while the trailing block is factored.
// Warp 0 runs the pivot chains; warps 1..7 apply each finished sub-panel to the view: net row permutation on every
// column, block row solve, rank-8 DMMA update -- the next sub-panel's eight columns FIRST, so that the chain warp
// starts on them while the bulk of the update is still running (look-ahead of one sub-panel).
#ifndef MB200_CHAIN
#define MB200_CHAIN panel_chain_rolled  // panel_chain: the fully unrolled variant (A/B builds: -DMB200_CHAIN=panel_chain)
#endif
template <int LD, bool TRACK_PERM, int NU, typename SM>
__device__ __forceinline__ void factor_view(SM &S, double *__restrict__ V, const int mv, const int nv, const int row_off,
                                            double *__restrict__ X, const int xcols, const int tid, const int lane, const int w)
{
    // NU update warps (warps 1..NU); warp 0 runs the chains. NT threads take part in the hand-over barriers.
    constexpr int NT = 32 * (NU + 1), NUT = 32 * NU, NH = 2 * NU;
    const int g = lane >> 2, q = lane & 3;
    const int kv = mv < nv ? mv : nv;
    const int tm = (mv + 7) >> 3, tn = (nv + 7) >> 3;
    if (w == 0) {
        const unsigned vbase = f_saddr(V);
        for (int j = 0; j < kv; j += 8) {
            const int jb = (kv - j) < 8 ? (kv - j) : 8;
            if (j > 0) f_bar_sync(BAR_NEXT, NT);
            const int k0 = j >> 5;
            const int na = ((mv + 31) >> 5) - k0;
            if (LD > 98 && na >= 4) MB200_CHAIN<(LD > 98 ? 4 : 1), LD>(S, vbase, j, jb, mv, k0, lane, row_off);
            else if (LD > 66 && na == 3) MB200_CHAIN<(LD > 66 ? 3 : 1), LD>(S, vbase, j, jb, mv, k0, lane, row_off);
            else if (LD > 34 && na == 2) MB200_CHAIN<(LD > 34 ? 2 : 1), LD>(S, vbase, j, jb, mv, k0, lane, row_off);
            else MB200_CHAIN<1, LD>(S, vbase, j, jb, mv, k0, lane, row_off);
            __threadfence_block();
            f_bar_arrive(BAR_PANEL, NT);
        }
        return;
    }
    const int wu = w - 1;            // 0..NU-1
    const int tu = tid - 32;         // 0..NUT-1
    const int hw = tu >> 4;          // half-warp 0..NH-1: owns the columns o = hw, hw + NH, ...
    const int e = lane & 15;
    for (int j = 0; j < kv; j += 8) {
        const int jb = (kv - j) < 8 ? (kv - j) : 8;
        f_bar_sync(BAR_PANEL, NT);
        // ---- the sub-panel's net row permutation on every column of the view (and of X) --------------------------
        const int nmv = S.nmoves;
        if (nmv > 0) {
            const bool eok = e < nmv;
            const int src = eok ? (int)S.msrc[e] : 0, dst = eok ? (int)S.mdst[e] : 0;
            auto apply = [&](double *__restrict__ base, const int ld, const int ncols) {
                for (int o0 = 0; o0 < ncols; o0 += NH * 5) {  // warp-uniform trip count: __syncwarp inside
                    double v[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const int o = o0 + hw + NH * u;
                        if (eok && o < ncols) v[u] = base[o * ld + src];
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const int o = o0 + hw + NH * u;
                        if (eok && o < ncols) base[o * ld + dst] = v[u];
                    }
                }
            };
            apply(V, LD, nv);
            if (xcols > 0) apply(X, FLD, xcols);
            if (TRACK_PERM && w == NU) {
                const unsigned char pp = S.perm[src];
                __syncwarp();
                if (lane < nmv) S.perm[dst] = pp;
            }
        }
        const int nright = nv - j - jb;
        if (nright > 0) {
            f_bar_sync(BAR_MOVED, NUT);
            // ---- block row of the columns to the right: U = L11^-1 * (rows j..j+jb-1), one thread per column
            for (int rc = tu; rc < nright; rc += NUT) {
                const int c = j + jb + rc;
                double *col = V + c * LD + j;
                double x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = (i < jb) ? col[i] : 0.0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
#pragma unroll
                    for (int i = k + 1; i < 8; ++i)
                        if (i < jb) x[i] = fma(-V[(j + k) * LD + j + i], x[k], x[i]);  // L11 from the permuted image
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < jb) {
                        col[i] = x[i];
                        S.Un[i * SM::LDU + c] = -x[i];
                    }
                }
            }
        }
        f_bar_sync(BAR_UPD, NUT);
        // ---- rows below, columns to the right: C -= L21 * U12 (k = 8: two DMMA per 8x8 tile) ---------------------
        const bool upd = (jb == 8) && (j + 8 < mv) && (j + 8 < nv);
        const int t0 = (j + 8) >> 3;
        if (upd) {
            // the next sub-panel's columns first
            const double bf0 = S.Un[q * SM::LDU + 8 * t0 + g];
            const double bf1 = S.Un[(4 + q) * SM::LDU + 8 * t0 + g];
            for (int t = t0 + wu; t < tm; t += NU) {
                const double af0 = V[(j + q) * LD + 8 * t + g];
                const double af1 = V[(j + 4 + q) * LD + 8 * t + g];
                double *cp = V + (8 * t0 + 2 * q) * LD + 8 * t + g;
                double c0 = cp[0], c1 = cp[LD];
                f_dmma(c0, c1, af0, bf0);
                f_dmma(c0, c1, af1, bf1);
                cp[0] = c0;
                cp[LD] = c1;
            }
        }
        if (j + 8 < kv) {
            __threadfence_block();
            f_bar_arrive(BAR_NEXT, NT);
        }
        if (upd && t0 + 1 < tn) {
            for (int t = t0 + wu; t < tm; t += NU) {
                const double af0 = V[(j + q) * LD + 8 * t + g];
                const double af1 = V[(j + 4 + q) * LD + 8 * t + g];
#pragma unroll 2
                for (int ct = t0 + 1; ct < tn; ++ct) {
                    const double bf0 = S.Un[q * SM::LDU + 8 * ct + g];
                    const double bf1 = S.Un[(4 + q) * SM::LDU + 8 * ct + g];
                    double *cp = V + (8 * ct + 2 * q) * LD + 8 * t + g;
                    double c0 = cp[0], c1 = cp[LD];
                    f_dmma(c0, c1, af0, bf0);
                    f_dmma(c0, c1, af1, bf1);
                    cp[0] = c0;
                    cp[LD] = c1;
                }
            }
        }
    }
}

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
        for (int c = w; c < jb; c += 2)
            for (int r = lane; r < mp; r += 32) S.V[c * LD + r] = A[r + (size_t)c * ld];
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
