// Auxiliary kernels: pointer-array builders, vbatched statistics / binning, the device dlarnv
// stream, and the FP64 / HBM microbenchmarks used as roofline denominators.
//
// Replaces magmablas/zset_pointer.cu:17-26,86-98,118-124,224-231 and magmablas/set_pointer.cu
// (<<<batch,1>>>: one thread per CTA) with flat 256-thread grids, and
// magmablas/vbatched_aux.cu:26-87 + magmablas/vbatched_check.cu:17-62 (two kernels, two blocking
// D2H reads) with one statistics kernel + one read.
#include "common.cuh"

namespace mb200 {

namespace {

__global__ void set_pointer_kernel(void **out, char *base, long elem, long lda, long row, long col,
                                   long batch_offset, long batch)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) out[b] = base + elem * (b * batch_offset + row + col * lda);
}

__global__ void displace_pointers_kernel(void **out, void **in, long elem, long lda, long row, long col, long batch)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) out[b] = (char *)in[b] + elem * (row + col * lda);
}

// dB_b = I (n x n): the right-hand side of the out-of-place inverse (replaces magmablas_zlaset_batched in
// src/zgetri_outofplace_batched.cpp:114)
__global__ void identity_kernel(int n, double **__restrict__ dB, int lddb, long batch)
{
    const long b = blockIdx.y;
    double *__restrict__ B = dB[b];
    const long total = (long)n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        B[i + (size_t)j * lddb] = (i == j) ? 1.0 : 0.0;
    }
}

__global__ void memset_int_kernel(int *p, int v, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// out8 must be zero-initialised except out8[4] = 0 (first bad argument, as a positive index).
__global__ void vbatched_stats_kernel(const int *__restrict__ m, const int *__restrict__ n,
                                      const int *__restrict__ ldda, long batch, int *out8)
{
    int mm = 0, mn_ = 0, mmin = 0, mxn = 0, bad = 0x7fffffff, small = 0, nonempty = 0, c64 = 0, c96 = 0, c128 = 0,
        c256 = 0, c384 = 0;
    for (long b = (long)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (long)gridDim.x * blockDim.x) {
        const int M = m[b], N = n[b], L = ldda[b];
        if (M < 0) bad = min(bad, 1);
        else if (N < 0) bad = min(bad, 2);
        else if (L < max(1, M)) bad = min(bad, 4);
        mm = max(mm, M);
        mn_ = max(mn_, N);
        mmin = max(mmin, min(M, N));
        const long long prod = (long long)max(M, 0) * max(N, 0);
        mxn = max(mxn, (int)min(prod, (long long)0x7fffffff));
        if (M > 0 && N > 0) {
            ++nonempty;
            const int K = max(M, N);
            if (K <= 32) ++small;
            else if (K <= 64) ++c64;
            else if (K <= 96) ++c96;
            else if (K <= 128) ++c128;
            else if (K <= 256) ++c256;
            else if (K <= 384) ++c384;
        }
    }
    const unsigned full = 0xffffffffu;
    mm = __reduce_max_sync(full, mm);
    mn_ = __reduce_max_sync(full, mn_);
    mmin = __reduce_max_sync(full, mmin);
    mxn = __reduce_max_sync(full, mxn);
    bad = __reduce_min_sync(full, bad);
    small = __reduce_add_sync(full, small);
    nonempty = __reduce_add_sync(full, nonempty);
    c64 = __reduce_add_sync(full, c64);
    c96 = __reduce_add_sync(full, c96);
    c128 = __reduce_add_sync(full, c128);
    c256 = __reduce_add_sync(full, c256);
    c384 = __reduce_add_sync(full, c384);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out8 + 0, mm);
        atomicMax(out8 + 1, mn_);
        atomicMax(out8 + 2, mmin);
        atomicMax(out8 + 3, mxn);
        if (bad != 0x7fffffff) atomicMax(out8 + 4, 8 - bad);  // smaller argument index wins
        atomicAdd(out8 + 5, small);
        atomicAdd(out8 + 6, nonempty);
        atomicAdd(out8 + 8, c64);
        atomicAdd(out8 + 9, c96);
        atomicAdd(out8 + 10, c128);
        atomicAdd(out8 + 11, c256);
        atomicAdd(out8 + 12, c384);
    }
}

// Index lists by size class of max(m, n): 0: <= 32 (register tier), 1: <= 64, 2: <= 96, 3: <= mid_max
// (register-file tier, one list per kernel shape), 4: <= 256, 5: <= 384, 6: the rest (blocked tier: each
// class runs its own step sequence sized for its own maximum, so a 130 x 130 matrix does not sit in
// launches sized for 512 x 512). List c lives at lists + c * batch. Order inside a list is arbitrary (atomic cursor); empty matrices are dropped.
__global__ void vbatched_partition_kernel(const int *__restrict__ m, const int *__restrict__ n, long batch,
                                          int *lists, int *counts, int mid_max)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    if (b < batch) {
        const int M = m[b], N = n[b];
        if (M > 0 && N > 0) {
            const int K = max(M, N);
            cls = K <= 32 ? 0 : (K > mid_max ? (K <= 256 ? 4 : (K <= 384 ? 5 : 6)) : (K <= 64 ? 1 : (K <= 96 ? 2 : 3)));
        }
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const unsigned bal = __ballot_sync(full, cls == c);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(counts + c, __popc(bal));
        base = __shfl_sync(full, base, 0);
        if (cls == c) lists[(size_t)c * batch + base + __popc(bal & lt)] = (int)b;
    }
}

// dlarnv(idist = 1): x_i = a^(i+1) * s mod 2^48 (see oracle/lu_oracle.c). Each thread jumps to its
// chunk with square-and-multiply over the precomputed powers a^(2^k), then steps sequentially.
struct LcgPowers { unsigned long long p[48]; };
constexpr unsigned long long LCG_A = 33952834046453ull;
constexpr unsigned long long MASK48 = (1ull << 48) - 1ull;
constexpr int LARNV_CHUNK = 8;

__global__ void dlarnv_kernel(unsigned long long seed, long long n, double *__restrict__ x, LcgPowers pw)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long start = t * LARNV_CHUNK;
    if (start >= n) return;
    unsigned long long s = seed;
    unsigned long long e = (unsigned long long)start;
#pragma unroll 1
    for (int k = 0; k < 48 && e; ++k, e >>= 1)
        if (e & 1ull) s = (s * pw.p[k]) & MASK48;
    const long long end = (start + LARNV_CHUNK < n) ? start + LARNV_CHUNK : n;
    for (long long i = start; i < end; ++i) {
        s = (s * LCG_A) & MASK48;
        x[i] = (double)s * 0x1.0p-48;
    }
}

// ---- microbenchmarks --------------------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = seed + i + threadIdx.x;
    const double a = 1.0 + 1e-9 * seed, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed)
{
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        c0[i] = seed + i;
        c1[i] = seed - i;
    }
    const double a = 1.0 + 1e-9 * threadIdx.x, bv = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c0[i]), "+d"(c1[i])
                             : "d"(a), "d"(bv));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n2)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

void set_pointer_launch(void **out, char *base, long elem, long lda, long row, long col, long batch_offset,
                        long batch, cudaStream_t s)
{
    if (batch <= 0) return;
    set_pointer_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>(out, base, elem, lda, row, col, batch_offset,
                                                                     batch);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("set_pointer_kernel");
}

void displace_pointers_launch(void **out, void **in, long elem, long lda, long row, long col, long batch,
                              cudaStream_t s)
{
    if (batch <= 0) return;
    displace_pointers_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>(out, in, elem, lda, row, col, batch);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("displace_pointers_kernel");
}

void identity_launch(int n, double **dB, int lddb, long batch, cudaStream_t s)
{
    if (n <= 0 || batch <= 0) return;
    const long total = (long)n * n;
    int gx = (int)((total + 255) / 256);
    if (gx > 64) gx = 64;
    for (long off = 0; off < batch; off += 65535) {
        const long cnt = batch - off < 65535 ? batch - off : 65535;
        identity_kernel<<<dim3((unsigned)gx, (unsigned)cnt), 256, 0, s>>>(n, dB + off, lddb, cnt);
        count_launch();
        MB200_CHECK_LAUNCH_VOID("identity_kernel");
    }
}

void memset_int_launch(int *p, int v, long n, cudaStream_t s)
{
    if (n <= 0) return;
    memset_int_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, v, n);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("memset_int_kernel");
}

void vbatched_stats_launch(const int *m, const int *n, const int *ldda, long batch, int *out8, cudaStream_t s)
{
    cudaMemsetAsync(out8, 0, 16 * sizeof(int), s);
    if (batch <= 0) return;
    long blocks = (batch + 255) / 256;
    if (blocks > 1184) blocks = 1184;  // 8 x 148 SMs
    vbatched_stats_kernel<<<(unsigned)blocks, 256, 0, s>>>(m, n, ldda, batch, out8);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("vbatched_stats_kernel");
}

void vbatched_partition_launch(const int *m, const int *n, long batch, int *lists, int *counts, int mid_max,
                               cudaStream_t s)
{
    cudaMemsetAsync(counts, 0, 8 * sizeof(int), s);
    if (batch <= 0) return;
    vbatched_partition_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>(m, n, batch, lists, counts, mid_max);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("vbatched_partition_kernel");
}

void dlarnv_launch(uint64_t seed48, int64_t n, double *dx, cudaStream_t s)
{
    if (n <= 0) return;
    LcgPowers pw;
    unsigned long long p = LCG_A;
    for (int k = 0; k < 48; ++k) {
        pw.p[k] = p;
        p = (p * p) & MASK48;
    }
    const long long threads = (n + LARNV_CHUNK - 1) / LARNV_CHUNK;
    dlarnv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(seed48, n, dx, pw);
    count_launch();
    MB200_CHECK_LAUNCH_VOID("dlarnv_kernel");
}

double fp64_peak_run(int kind, cudaStream_t s)
{
    double *out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 8, iters = 4096;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, s);
        if (kind == 0) dfma_peak_kernel<8><<<grid, 256, 0, s>>>(out, iters, 1.0);
        else dmma_peak_kernel<8><<<grid, 256, 0, s>>>(out, iters, 1.0);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        count_launch();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (kind == 0) flops = 2.0 * 8 * 8 * (double)iters * 256.0 * grid;                 // ILP*8 fma / iter / thread
        else flops = 512.0 * 8 * 4 * (double)iters * (256.0 / 32.0) * grid;                // 8x8x4 mma = 512 flop / warp
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

double hbm_copy_run(size_t bytes, cudaStream_t s)
{
    double2 *a = nullptr, *b = nullptr;
    bytes &= ~(size_t)15;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        cudaGetLastError();
        if (a) cudaFree(a);
        return -1.0;
    }
    cudaMemsetAsync(a, 1, bytes, s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, s);
        copy_kernel<<<sms * 16, 512, 0, s>>>(a, b, bytes / 16);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        count_launch();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return best;
}

}  // namespace mb200
