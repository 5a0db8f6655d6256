// Out-of-place inverse from the LU factors, n <= 64, one launch: inv(A) = U^-1 L^-1 P.
//
// Replaces src/zgetri_outofplace_batched.cpp:81-141 for small matrices -- identity fill (zlaset), row interchanges
// (zlaswp_rowserial), and the two recursive trsm drivers (magmablas/ztrsm_batched_core.cpp) -- with ONE kernel.
// The generic path (identity + getrs, api.cu) gives every CTA a 16-column tile of right-hand sides and spreads the ROWS
// over its warps: at n = 64 one warp of eight works and the factors are read four times (27.5 ms per 100000 matrices).
// Here a CTA of four warps owns a matrix:
//   * the factors come in once, by 1-D TMA copies (one per column) into a padded shared-memory image;
//   * the right-hand side is the permuted identity, built in registers from the pivot trace: no identity in memory;
//   * warp w holds columns [8 CT w, 8 CT (w+1)) of X for ALL rows as DMMA accumulator fragments; per 8-row block the
//     diagonal block is solved inside the warp (shuffles between the lanes of a fragment), every other row tile gets
//     C -= T_tile * X_blk as mma.sync.m8n8k4.f64 with the A fragments read from the image. The backward sweep feeds
//     the fragments in reversed k.
// Every unknown therefore sees exactly the update order of oracle_dgetrs (oracle/lu_oracle.c: k increasing through L,
// k decreasing through U, multiply by the inverted diagonal): results are bit-identical to the generic path.
#include "common.cuh"

namespace mb200 {

std::atomic<int> g_getri_fused{1};  // 0: identity + getrs for every n (A/B runs, tests)

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void gi_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT>
struct GetriSmem {
    static constexpr int NMAX = 8 * NT;
    static constexpr int LD = NMAX + 8;  // A fragments (rows g, columns q): g + q LD distinct mod 16 -> two wavefronts, the minimum
    alignas(16) double T[NMAX * LD];     // the factors, zero outside n x n
    double dinv[NMAX];                   // 1 / u(k,k)
    int ipiv[NMAX];
    int perm[NMAX];                      // original row that the forward interchanges leave at position i
    unsigned long long bar;
};

// NT: 8-row tiles covered (n <= 8 NT); CT: 8-column tiles of X per warp (4 warps: n <= 32 CT)
template <int NT, int CT, int MINB>
__global__ void __launch_bounds__(128, MINB)
getri_fused_kernel(int n, double *const *__restrict__ dA, int ldda, int *const *__restrict__ dipiv,
                   double *const *__restrict__ dinvA, int lddia, long batch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = GetriSmem<NT>;
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    constexpr int LD = SM::LD;
    const long b = blockIdx.x;
    if (b >= batch) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nt = (n + 7) >> 3;
    const int np = 8 * nt;
    const double *__restrict__ A = dA[b];
    const bool bulk = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ldda & 1) == 0) && ((n & 1) == 0);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&S.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < n) S.ipiv[tid] = dipiv[b][tid];
    // zero frame around the n x n factors (the last tile's unused rows and columns)
    if (np > n) {
        const int padr = np - n;
        for (int idx = tid; idx < padr * np; idx += 128) {
            const int j = idx / padr, i = n + (idx - j * padr);
            S.T[i + j * LD] = 0.0;
        }
        for (int idx = tid; idx < padr * n; idx += 128) {
            const int j = n + idx / n, i = idx % n;
            S.T[i + j * LD] = 0.0;
        }
    }
    __syncthreads();
    if (bulk) {
        if (tid < 32) {
            if (tid == 0)
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(&S.bar)),
                             "r"((unsigned)(n * n) * 8u)
                             : "memory");
            __syncwarp();
            for (int j = tid; j < n; j += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(&S.T[j * LD])),
                             "l"(A + (size_t)j * ldda), "r"((unsigned)n * 8u), "r"((unsigned)__cvta_generic_to_shared(&S.bar))
                             : "memory");
        }
    } else {
        for (int j = w; j < n; j += 4)
            for (int i = lane; i < n; i += 32) S.T[i + j * LD] = A[i + (size_t)j * ldda];
    }
    // the pivot trace runs while the copies are in flight
    if (tid < np) {
        int r = tid;
        if (tid < n) {
            for (int k = n - 1; k >= 0; --k) {
                const int p = S.ipiv[k] - 1;
                r = (r == k) ? p : ((r == p) ? k : r);
            }
        }
        S.perm[tid] = (tid < n) ? r : -1;
    }
    if (bulk) {
        asm volatile(
            "{\n\t.reg .pred p;\n"
            "GIWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 0x989680;\n\t"
            "@!p bra GIWAIT_%=;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&S.bar))
            : "memory");
    } else {
        __syncthreads();
    }
    if (tid < np) S.dinv[tid] = (tid < n) ? 1.0 / S.T[tid + tid * LD] : 1.0;
    __syncthreads();

    const int col0 = 8 * CT * w;
    if (col0 >= n) return;  // no CTA-wide barrier below

    // ---- X = P I, as accumulator fragments: lane (g, q) holds X(8 rt + g, col0 + 8 ct + 2q + {0, 1}) -----------------
    double acc[NT][CT][2];
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
        const int p = (rt < nt) ? S.perm[8 * rt + g] : -1;
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            const int j = col0 + 8 * ct + 2 * q;
            acc[rt][ct][0] = (p == j) ? 1.0 : 0.0;
            acc[rt][ct][1] = (p == j + 1) ? 1.0 : 0.0;
        }
    }
    const double *Tg = S.T + g;  // this lane's row inside a tile

    // ---- L y = P I : k increasing ------------------------------------------------------------------------------------
#pragma unroll
    for (int kb = 0; kb < NT; ++kb) {
        if (kb < nt) {
            const double *Td = Tg + 8 * kb + (8 * kb) * LD;  // T(8kb + g, 8kb + .)
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const double l = Td[k * LD];
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double xk = __shfl_sync(FULL, acc[kb][ct][e], 4 * k + q);
                        if (g > k) acc[kb][ct][e] = fma(-l, xk, acc[kb][ct][e]);
                    }
                }
            }
            if (kb + 1 < nt) {
                // B fragments of the solved block: lane (g, q) needs X(4h + q, g) = element g & 1 of lane (4h + q, g >> 1)
                double bf[CT][2];
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src = (4 * h + q) * 4 + (g >> 1);
                        const double v0 = __shfl_sync(FULL, acc[kb][ct][0], src);
                        const double v1 = __shfl_sync(FULL, acc[kb][ct][1], src);
                        bf[ct][h] = (g & 1) ? v1 : v0;
                    }
                }
#pragma unroll
                for (int rt = kb + 1; rt < NT; ++rt) {
                    if (rt < nt) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const double af = -Tg[8 * rt + (8 * kb + 4 * h + q) * LD];
#pragma unroll
                            for (int ct = 0; ct < CT; ++ct) gi_dmma(acc[rt][ct][0], acc[rt][ct][1], af, bf[ct][h]);
                        }
                    }
                }
            }
        }
    }

    // ---- U x = y : k decreasing, multiply by the inverted diagonal ----------------------------------------------------
#pragma unroll
    for (int kb = NT - 1; kb >= 0; --kb) {
        if (kb < nt) {
            const double *Td = Tg + 8 * kb + (8 * kb) * LD;
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                const double dk = S.dinv[8 * kb + k];
                const double u = Td[k * LD];
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (g == k) acc[kb][ct][e] = acc[kb][ct][e] * dk;
                        if (k > 0) {
                            const double xk = __shfl_sync(FULL, acc[kb][ct][e], 4 * k + q);
                            if (g < k) acc[kb][ct][e] = fma(-u, xk, acc[kb][ct][e]);
                        }
                    }
                }
            }
            if (kb > 0) {
                // reversed k: slot 4h + q of the instruction carries k = 7 - (4h + q)
                double bf[CT][2];
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src = (7 - (4 * h + q)) * 4 + (g >> 1);
                        const double v0 = __shfl_sync(FULL, acc[kb][ct][0], src);
                        const double v1 = __shfl_sync(FULL, acc[kb][ct][1], src);
                        bf[ct][h] = (g & 1) ? v1 : v0;
                    }
                }
#pragma unroll
                for (int rt = 0; rt < kb; ++rt) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double af = -Tg[8 * rt + (8 * kb + 7 - (4 * h + q)) * LD];
#pragma unroll
                        for (int ct = 0; ct < CT; ++ct) gi_dmma(acc[rt][ct][0], acc[rt][ct][1], af, bf[ct][h]);
                    }
                }
            }
        }
    }

    // ---- store ----------------------------------------------------------------------------------------------------------
    double *__restrict__ X = dinvA[b];
#pragma unroll
    for (int rt = 0; rt < NT; ++rt) {
        const int i = 8 * rt + g;
        if (i < n) {
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) {
                const int j = col0 + 8 * ct + 2 * q;
                if (j < n) X[i + (size_t)j * lddia] = acc[rt][ct][0];
                if (j + 1 < n) X[i + (size_t)(j + 1) * lddia] = acc[rt][ct][1];
            }
        }
    }
}

template <int NT, int CT, int MINB>
magma_int_t launch_getri(int n, double **dA, int ldda, int **dipiv, double **dinvA, int lddia, long batch, cudaStream_t s)
{
    const size_t smem = sizeof(GetriSmem<NT>);
    static DevOnce once;
    smem_optin(once, getri_fused_kernel<NT, CT, MINB>, smem);
    constexpr long CHUNK = 1L << 24;  // grid.x stays far below 2^31
    for (long off = 0; off < batch; off += CHUNK) {
        const long cnt = (batch - off) < CHUNK ? (batch - off) : CHUNK;
        getri_fused_kernel<NT, CT, MINB><<<(unsigned)cnt, 128, smem, s>>>(n, dA + off, ldda, dipiv + off, dinvA + off, lddia, cnt);
        count_launch();
        MB200_CHECK_LAUNCH("getri_fused_kernel");
    }
    return 0;
}

}  // namespace

// -100: not covered (the caller runs identity + getrs)
magma_int_t getri_fused_launch(int n, double **dA, int ldda, int **dipiv, double **dinvA, int lddia, long batch,
                               cudaStream_t s)
{
    if (!g_getri_fused || n > 64 || batch <= 0) return -100;
    if (n <= 32) return launch_getri<4, 1, 8>(n, dA, ldda, dipiv, dinvA, lddia, batch, s);
    return launch_getri<8, 2, 4>(n, dA, ldda, dipiv, dinvA, lddia, batch, s);
}

}  // namespace mb200

extern "C" void magma_b200_set_getri_fused(int on) { mb200::g_getri_fused = on; }
