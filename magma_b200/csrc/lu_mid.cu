// Tier M: LU of one matrix per CTA with the WHOLE matrix resident in the register file
// (32 < max(m, n) <= 128): one HBM read and one HBM write per matrix, one launch per call.
//
// Replaces, for these sizes, the reference's blocked host loop (src/zgetrf_batched.cpp:149-203 ->
// recursive panel src/zgetrf_panel_batched.cpp:101-196 -> fused panel magmablas/zgetf2_kernels.cu:
// 831-1058, setup_pivinfo/adjust_ipiv magmablas/getrf_setup_pivinfo.cu, row-parallel laswp
// magmablas/zlaswp_batched.cu:31-87, recursive trsm magmablas/ztrsm_batched_core.cpp:33-95 and
// cublasDgemmBatched): 42 launches and ~3.7x the algorithmic HBM traffic at n = 128.
//
// Layout: warp w owns the 8 columns [8w, 8w+8); lane l holds rows l, l+32, ... (R of them) of those
// columns in registers (R x 8 doubles). A 128 x 128 matrix is 16 warps x 64 registers of data per
// thread. Right-looking, COLUMN-pipelined:
//   * column j is eliminated by the warp that owns it, entirely inside that warp: pivot search on
//     the high words of |x| (max over the lane's R rows, CREDUX across lanes, one ballot; an exact
//     64-bit "first maximum" search runs only when two candidates share a high word), pivot row
//     broadcast through a private shared-memory row buffer, reciprocal of each lane's own
//     candidate computed while the search is in flight. Row interchanges are lazy: a row never
//     leaves its (lane, r) slot; the slot <-> position maps live in shared memory and are kept by
//     the pivot lane alone (nobody else needs positions until the final store);
//   * as soon as the multipliers of column j exist they are published -- L(:, j) by row slot (zero
//     for rows already pivoted, so consumers need no predicate), the pivot row's slot -- and the
//     column's mbarrier is completed; the owner then finishes its own rank-1 update;
//   * every warp to the right applies column j to its 8 columns as soon as it is published:
//     U(j, :) of its columns is broadcast from the pivot row's lane by shuffle (the pivot slot is
//     warp-uniform, so the choice among the lane's R rows is a uniform branch), then
//     a(r, c) = fma(-l(r, j), u(c), a(r, c)). The next panel's owner trails the current one by one
//     column, so panels follow each other without a bubble;
//   * at the end each warp stores its columns at the final row positions.
// The complete L factor stays in shared memory (128 columns x 1 KB at n = 128): no buffer
// recycling, no consumer counting. Every element sees a(i,j) = fma(-l(i,k), u(k,j), a(i,j)) for k
// increasing: the canonical order of oracle/lu_oracle.c, results are bit-identical to it.
#include "lu_common.cuh"

namespace mb200 {

namespace {

constexpr int CW = 8;  // columns per warp
constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    const int lo = __shfl_sync(FULLM, __double2loint(v), src);
    const int hi = __shfl_sync(FULLM, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ unsigned hi_abs_m(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

// phase 0 of a single-use barrier; the hint lets the hardware park the warp instead of polling
__device__ __forceinline__ void mbar_wait0(unsigned bar)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], 0, 0x989680;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar)
        : "memory");
}

__device__ __forceinline__ int ld_acquire_s32(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_s32(int *p, int v)
{
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// Shared memory of one CTA (dynamic).
template <int R, int NW, int PH>
struct MidSmem {
    double Lcol[NW * CW][R * 32];   // multipliers of column j (of the current phase) by row slot (0: row not updated)
    double rowbuf[NW][2][CW + 2];   // per warp: pivot row ping-pong, [CW] = 1/pivot
    int pivslot[PH * NW * CW];      // slot (r*32 + lane) of the pivot row of column j
    int pos_of[R * 32];             // slot -> current row position
    int slot_at[R * 32];            // row position -> slot
    int ipiv[PH * NW * CW];
    int colready;                   // columns published so far
    int info;
    unsigned long long bar[PH * NW * CW];
};

// PH = 2: the CTA holds only HALF of the columns in registers at a time. Phase 0 factors column blocks
// 0..NW-1 and writes them back in their original row order; phase 1 loads blocks NW..2NW-1, catches up on the
// (already published) columns of phase 0, then factors. At the end the phase-0 columns are re-read and stored
// at their final row positions. Half the registers per matrix => two matrices per SM at n = 128, i.e. two
// independent pivot chains to interleave: the chain (~800 cycles per column), not FP64 or HBM, bounds this tier.
template <int R, int NW, int PH, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
lu_mid_kernel(Dims d, double **__restrict__ dA, int **__restrict__ dipiv, int *__restrict__ dinfo, long batch,
              const int *__restrict__ index_list)
{
    static_assert(PH == 1 || PH == 2, "phases");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MidSmem<R, NW, PH> &sm = *reinterpret_cast<MidSmem<R, NW, PH> *>(smem_raw);

    const long slotb = blockIdx.x;
    const long b = index_list ? index_list[slotb] : slotb;
    if (b < 0) return;  // unused tail of a vbatched index list
    int m, n, ld;
    dims_of(d, b, m, n, ld);
    const int mn = m < n ? m : n;
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(FULLM, tid >> 5, 0);  // warp-uniform for the compiler
    const int nblocks = (n + CW - 1) / CW;          // column blocks that hold data

    if (tid == 0) {
        sm.info = 0;
        sm.colready = 0;
    }
    for (int i = tid; i < PH * NW * CW; i += NW * 32) mbar_init((unsigned)__cvta_generic_to_shared(&sm.bar[i]), 1);
    for (int i = tid; i < R * 32; i += NW * 32) {
        sm.pos_of[i] = i;  // slot r*32 + lane holds row lane + 32 r
        sm.slot_at[i] = i;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    double *__restrict__ A = dA[b];
    double a[R][CW];

    // one published column applied to this warp's block: U(j, block) from the pivot row's lane, then the update
    auto apply_column = [&](int j, int lidx) {
        const int s = sm.pivslot[j];
        const int owner = s & 31, rk = s >> 5;
        double l[R];
#pragma unroll
        for (int r = 0; r < R; ++r) l[r] = sm.Lcol[lidx][r * 32 + lane];
        double u[CW];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (rk == r) {  // warp-uniform
#pragma unroll
                for (int c = 0; c < CW; ++c) u[c] = shfl_f64(a[r][c], owner);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = 0; c < CW; ++c) a[r][c] = fma(-l[r], u[c], a[r][c]);
        }
    };

#pragma unroll 1
  for (int ph = 0; ph < PH; ++ph) {
    const int pbase = ph * NW * CW;                       // first column of this phase
    const int c0 = pbase + w * CW;
    const bool has_block = (ph * NW + w) < nblocks;       // this warp holds columns in this phase
    const int nc = (n - c0) < CW ? (n - c0) : CW;         // columns of this block
    if (has_block) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = lane + 32 * r;
#pragma unroll
            for (int c = 0; c < CW; ++c) a[r][c] = (row < m && c < nc) ? A[row + (size_t)(c0 + c) * ld] : 0.0;
        }
    }
    const int jend = has_block ? (c0 < mn ? c0 : mn) : 0;
    if (PH > 1 && ph > 0) {
        // catch up on the previous phase: all of it is published, its multipliers still sit in Lcol
        const int jc = jend < pbase ? jend : pbase;
#pragma unroll 1
        for (int j = 0; j < jc; ++j) apply_column(j, j);
        __syncthreads();  // every warp is done with the previous phase's Lcol: this phase may overwrite it
    }
    // ---- columns of this phase to the left: apply each one as soon as it is published ---------------------
    int ready = pbase;
#pragma unroll 1
    for (int j = pbase; j < jend; ++j) {
        if (j >= ready) {
            ready = ld_acquire_s32(&sm.colready);
            if (j >= ready) {
                mbar_wait0((unsigned)__cvta_generic_to_shared(&sm.bar[j]));
                ready = j + 1;
            }
        }
        apply_column(j, j - pbase);
    }

    // ---- this warp's own columns ---------------------------------------------------------------------------
    if (has_block && c0 < mn) {
        const int jb = (mn - c0) < CW ? (mn - c0) : CW;
        // done: bit r set when row slot (lane, r) has been a pivot row (or does not exist)
        unsigned done = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = lane + 32 * r;
            const bool dn = (row >= m) || (sm.pos_of[r * 32 + lane] < c0);
            done |= dn ? (1u << r) : 0u;
        }
        int info = 0;
        const unsigned rowbuf = (unsigned)__cvta_generic_to_shared(&sm.rowbuf[w][0][0]);
#pragma unroll
        for (int jj = 0; jj < CW; ++jj) {
            if (jj < jb) {
                const int j = c0 + jj;
                // ---- search: high words of |x| over the lane's live rows, then across lanes ----------
                unsigned h[R];
#pragma unroll
                for (int r = 0; r < R; ++r) h[r] = (done >> r) & 1u ? 0u : hi_abs_m(a[r][jj]);
                unsigned hm = h[0];
                int lr = 0;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    if (h[r] > hm) {
                        hm = h[r];
                        lr = r;
                    }
                }
                int nmatch = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) nmatch += (h[r] == hm && !((done >> r) & 1u)) ? 1 : 0;
                double cv = a[0][jj];
#pragma unroll
                for (int r = 1; r < R; ++r)
                    if (lr == r) cv = a[r][jj];
                double rinv = 1.0 / cv;  // every lane inverts its own candidate during the search
                const unsigned mx = __reduce_max_sync(FULLM, hm);
                bool cand = (hm == mx) && (done != (1u << R) - 1u);
                unsigned bal = __ballot_sync(FULLM, cand);
                if (mx == 0u || __popc(bal) != 1 || __any_sync(FULLM, cand && nmatch > 1)) {
                    // exact search: 64-bit compare, ties to the smallest current row position (also taken for an
                    // all-zero high word: there `lr` may point at a row that is already pivoted)
                    unsigned long long lb = 0;
                    int lp = NOPOS_I;
                    lr = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (!((done >> r) & 1u)) {
                            const unsigned long long v = (unsigned long long)__double_as_longlong(a[r][jj]) & 0x7fffffffffffffffull;
                            const int pr = sm.pos_of[r * 32 + lane];
                            if (lp == NOPOS_I || v > lb || (v == lb && pr < lp)) {
                                lb = v;
                                lp = pr;
                                lr = r;
                            }
                        }
                    }
                    unsigned long long wb;
                    int wp;
                    warp_argmax(lb, lp, wb, wp);
                    cand = (lp == wp) && (lp != NOPOS_I);
                    bal = __ballot_sync(FULLM, cand);
                    cv = a[0][jj];
#pragma unroll
                    for (int r = 1; r < R; ++r)
                        if (lr == r) cv = a[r][jj];
                    rinv = 1.0 / cv;
                }
                // ---- pivot lane: publish its row, keep the permutation ----------------------------------
                const unsigned buf = rowbuf + (unsigned)((jj & 1) * (CW + 2) * 8);
                if (cand) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (lr == r) {
#pragma unroll
                            for (int c = (jj & ~1); c < CW; c += 2)
                                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(buf + c * 8), "d"(a[r][c]), "d"(a[r][c + 1]) : "memory");
                        }
                    }
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(buf + CW * 8), "d"(rinv) : "memory");
                    const int sp = lr * 32 + lane;    // pivot slot
                    const int pp = sm.pos_of[sp];     // its current position
                    const int sj = sm.slot_at[j];     // the slot sitting at position j
                    sm.pos_of[sp] = j;
                    sm.pos_of[sj] = pp;               // (sp == sj when the pivot is already in place)
                    sm.slot_at[pp] = sj;
                    sm.slot_at[j] = sp;
                    if (sp == sj) sm.pos_of[sp] = j;
                    sm.pivslot[j] = sp;
                    sm.ipiv[j] = pp + 1;
                    done |= 1u << lr;
                }
                __syncwarp();
                double u[CW], rr;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rr) : "r"(buf + CW * 8) : "memory");
#pragma unroll
                for (int c = (jj & ~1); c < CW; c += 2)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u[c]), "=d"(u[c + 1]) : "r"(buf + c * 8) : "memory");
                const bool nz = (u[jj] != 0.0);
                if (!nz && info == 0) info = j + 1;
                // ---- multipliers: computed, published, then used ---------------------------------------
                double l[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const bool upd = nz && !((done >> r) & 1u);
                    l[r] = upd ? a[r][jj] * rr : 0.0;
                    if (upd) a[r][jj] = l[r];
                    sm.Lcol[j - pbase][r * 32 + lane] = l[r];
                }
                if (info != 0 && lane == 0 && sm.info == 0) sm.info = info;  // columns finish in order
                __syncwarp();
                if (lane == 0) {
                    st_release_s32(&sm.colready, j + 1);
                    mbar_arrive((unsigned)__cvta_generic_to_shared(&sm.bar[j]));
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int c = jj + 1; c < CW; ++c) a[r][c] = fma(-l[r], u[c], a[r][c]);
                }
            }
        }
    }

    if (PH > 1 && ph + 1 < PH) {
        // end of a non-final phase: wait until its last column is out (the catch-up of the next phase does not
        // wait), then park this block in global memory in its ORIGINAL row order (slot order, coalesced)
        const int pend = mn < pbase + NW * CW ? mn : pbase + NW * CW;
        if (pend > pbase) mbar_wait0((unsigned)__cvta_generic_to_shared(&sm.bar[pend - 1]));
        if (has_block) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = lane + 32 * r;
#pragma unroll
                for (int c = 0; c < CW; ++c)
                    if (row < m && c < nc) A[row + (size_t)(c0 + c) * ld] = a[r][c];
            }
        }
    }
  }  // phases

    // ---- final row positions: known once the last column is done -------------------------------------------
    if (mn > 0) mbar_wait0((unsigned)__cvta_generic_to_shared(&sm.bar[mn - 1]));
    int q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] = sm.pos_of[r * 32 + lane];
#pragma unroll 1
    for (int ph = PH - 1; ph >= 0; --ph) {
        const int c0 = (ph * NW + w) * CW;
        if ((ph * NW + w) >= nblocks) continue;
        const int nc = (n - c0) < CW ? (n - c0) : CW;
        if (ph < PH - 1) {
            // parked block: read it back (same thread wrote these addresses), all rows before any store
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = lane + 32 * r;
#pragma unroll
                for (int c = 0; c < CW; ++c) a[r][c] = (row < m && c < nc) ? A[row + (size_t)(c0 + c) * ld] : 0.0;
            }
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = lane + 32 * r;
            if (row < m) {
#pragma unroll
                for (int c = 0; c < CW; ++c)
                    if (c < nc) A[q[r] + (size_t)(c0 + c) * ld] = a[r][c];
            }
        }
    }
    if (w == 0) {
        int *ip = dipiv[b];
        for (int i = lane; i < mn; i += 32) ip[i] = sm.ipiv[i];
        if (lane == 0) dinfo[b] = sm.info;
    }
}

template <int R, int NW, int PH, int MINB>
magma_int_t launch_mid(const Dims &d, double **dA, int **dipiv, int *dinfo, long batch, const int *il, cudaStream_t s)
{
    auto k = lu_mid_kernel<R, NW, PH, MINB>;
    const size_t smem = sizeof(MidSmem<R, NW, PH>);
    static DevOnce once;
    smem_optin(once, k, smem);
    k<<<(unsigned)batch, NW * 32, smem, s>>>(d, dA, dipiv, dinfo, batch, il);
    count_launch();
    MB200_CHECK_LAUNCH("lu_mid_kernel");
    return 0;
}

}  // namespace

// Register-file-resident LU for max_m <= 128, max_n <= 128. Returns -100 when not covered.
// R (rows per lane) from the row count, phases from the column count; 8 warps per CTA throughout.
magma_int_t lu_mid_launch(const Dims &d, int max_m, int max_n, double **dA, int **dipiv, int *dinfo, long batch,
                          const int *index_list, cudaStream_t s)
{
    if (batch <= 0) return 0;
    if (max_m > 128 || max_n > 128) return -100;
    {
        // max(m, n) > 44: the left-looking slab driver of the blocked tier is ahead (per 50000*128^2-equivalent batch:
        // n = 128 12.2 vs 15.2 ms, n = 96 10.7 vs 18.7, n = 64 9.6 vs 11.5, n = 48 13.7 vs 14.4; n = 40 17.9 vs 16.3) -- short panels with 6..12 CTAs
        // per SM and 4-warp slab updates beat one pivot chain per matrix. magma_b200_set_small_rows(7) keeps this
        // tier up to 128 (A/B runs, tests).
        const int mx = max_m > max_n ? max_m : max_n;
        if (mx > XOVER_MID && g_small_rows != 7 && g_small_rows != 8) return -100;
    }
    if (g_small_rows == 8) {  // the single-phase 16-warp kernel (A/B runs)
        if (max_m <= 64 && max_n <= 64) return launch_mid<2, 8, 1, 3>(d, dA, dipiv, dinfo, batch, index_list, s);
        return launch_mid<4, 16, 1, 1>(d, dA, dipiv, dinfo, batch, index_list, s);
    }
    if (max_m <= 64) {
        // four warps (32 columns per phase): six CTAs, i.e. six pivot chains, per SM instead of three
        // (n = 40: 16.3 -> 14.2 ms per 512000 matrices, n = 64: 11.5 -> 9.7 ms per 200000)
        if (max_n <= 32) return launch_mid<2, 4, 1, 6>(d, dA, dipiv, dinfo, batch, index_list, s);
        if (max_n <= 64) return launch_mid<2, 4, 2, 6>(d, dA, dipiv, dinfo, batch, index_list, s);
        return launch_mid<2, 8, 2, 3>(d, dA, dipiv, dinfo, batch, index_list, s);
    }
    if (max_n <= 64) return launch_mid<4, 8, 1, 2>(d, dA, dipiv, dinfo, batch, index_list, s);
    return launch_mid<4, 8, 2, 2>(d, dA, dipiv, dinfo, batch, index_list, s);
}

}  // namespace mb200
