// Single-warp pivot chains and the factorisation of a shared-memory resident block ("view"): the pieces shared by the
// single-launch tier (lu_fused.cu), the chain panel kernel (lu_fused.cu) and the fused tail of the left-looking slab kernel
// (lu_blocked.cu). See lu_fused.cu for the design notes. Everything here is internal linkage (one copy per translation unit).
#pragma once
#include "lu_common.cuh"
#include <type_traits>

namespace mb200 {

namespace {

constexpr int FLD = 130;   // leading dimension of the single-launch tier's left part (the X operand of factor_view)

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
    // every lane has read its rows before any lane stores a pivot row's U part into another lane's row (same warp, in
    // program order anyway; the barrier states it for compute-sanitizer's racecheck)
    __syncwarp();
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
    // every lane has read its rows before any lane stores a pivot row's U part into another lane's row (same warp, in
    // program order anyway; the barrier states it for compute-sanitizer's racecheck)
    __syncwarp();
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

}  // namespace

}  // namespace mb200
