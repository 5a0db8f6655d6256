// Standalone runtime boundary of libmagma_b200.so: init, devices, queues, memory, copies, errors,
// timers. Replaces (for this path) interface_cuda/{interface,alloc,copy_v2,error}.cpp and
// control/{xerbla,magma_timer}.cpp of the reference; same names, arguments and error behaviour.
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "common.cuh"

namespace mb200 {
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_tier{0};
std::atomic<int> g_chain_panel{3};
std::atomic<int> g_fused_tail{0};
std::atomic<int> g_tall_panel2{1};
std::atomic<int> g_split{0};
static std::atomic<int> g_init_count{0};

#ifdef MB200_INTERPOSE
}  // namespace mb200
#include <mutex>
#include <unordered_map>
namespace mb200 {
QState *qstate(magma_queue_t q)
{
    static std::mutex mu;
    static std::unordered_map<void *, QState *> table;
    std::lock_guard<std::mutex> lock(mu);
    QState *&st = table[(void *)q];
    if (!st) st = (QState *)calloc(1, sizeof(QState));
    // the handle may have been destroyed and re-created by its owner: always ask the real library
    st->stream = (cudaStream_t)magma_queue_get_cuda_stream(q);
    st->device = magma_queue_get_device(q);
    return st;
}
#endif

void *queue_dscratch(magma_queue_t queue, size_t bytes, int slot)
{
    auto *q = MB200_Q(queue);
    if (q->dscratch_bytes[slot] < bytes) {
        // the scratch belongs to the queue's device, whatever device is current
        int prev = 0;
        cudaGetDevice(&prev);
        if (prev != q->device) cudaSetDevice(q->device);
        if (q->dscratch[slot]) {
            cudaStreamSynchronize(q->stream);
            cudaFree(q->dscratch[slot]);
        }
        q->dscratch[slot] = nullptr;
        q->dscratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4;
        const cudaError_t e = cudaMalloc(&q->dscratch[slot], want);
        if (prev != q->device) cudaSetDevice(prev);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        q->dscratch_bytes[slot] = want;
    }
    return q->dscratch[slot];
}

void *queue_hscratch(magma_queue_t queue, size_t bytes)
{
    auto *q = MB200_Q(queue);
    if (q->hscratch_bytes < bytes) {
        if (q->hscratch) cudaFreeHost(q->hscratch);
        q->hscratch = nullptr;
        q->hscratch_bytes = 0;
        if (cudaMallocHost(&q->hscratch, bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        q->hscratch_bytes = bytes;
    }
    return q->hscratch;
}
}  // namespace mb200

using namespace mb200;

#ifndef MB200_INTERPOSE  // interpose mode: everything below is the real libmagma's
extern "C" {

// ---------------------------------------------------------------------------------------------
magma_int_t magma_init(void)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        // No CPU fallback exists on this path: say so loudly.
        fprintf(stderr, "libmagma_b200: magma_init: no CUDA device (%s)\n", cudaGetErrorString(e));
        cudaGetLastError();
        return MAGMA_ERR_NOT_INITIALIZED;
    }
    g_init_count.fetch_add(1);
    return MAGMA_SUCCESS;
}

magma_int_t magma_finalize(void)
{
    if (g_init_count.load() <= 0) {
        fprintf(stderr, "Error: magma_finalize() called, but MAGMA library is not initialized.\n");
        return MAGMA_ERR_NOT_INITIALIZED;
    }
    g_init_count.fetch_sub(1);
    return MAGMA_SUCCESS;
}

void magma_version(magma_int_t *major, magma_int_t *minor, magma_int_t *micro)
{
    if (major) *major = MAGMA_VERSION_MAJOR;
    if (minor) *minor = MAGMA_VERSION_MINOR;
    if (micro) *micro = MAGMA_VERSION_MICRO;
}

void magma_print_environment(void)
{
    int rt = 0, drv = 0, ndev = 0;
    cudaRuntimeGetVersion(&rt);
    cudaDriverGetVersion(&drv);
    cudaGetDeviceCount(&ndev);
    printf("%% libmagma_b200 (MAGMA %d.%d.%d batched-LU ABI), 32-bit magma_int_t, sm_100a only\n",
           MAGMA_VERSION_MAJOR, MAGMA_VERSION_MINOR, MAGMA_VERSION_MICRO);
    printf("%% CUDA runtime %d, driver %d.\n", rt, drv);
    for (int d = 0; d < ndev; ++d) {
        cudaDeviceProp p;
        cudaGetDeviceProperties(&p, d);
        printf("%% device %d: %s, %.1f MiB memory, capability %d.%d, %d SMs\n", d, p.name,
               p.totalGlobalMem / 1048576.0, p.major, p.minor, p.multiProcessorCount);
    }
}

magma_int_t magma_num_gpus(void)
{
    const char *s = getenv("MAGMA_NUM_GPUS");
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (s) {
        char *end;
        long v = strtol(s, &end, 10);
        if (end == s || *end != '\0' || v < 1) {
            fprintf(stderr, "$MAGMA_NUM_GPUS='%s' is an invalid number; using 1 GPU.\n", s);
            return 1;
        }
        if (v > ndev) {
            fprintf(stderr, "$MAGMA_NUM_GPUS='%s' exceeds number of CUDA devices=%d; using %d GPUs.\n",
                    s, ndev, ndev);
            return ndev > 0 ? ndev : 1;
        }
        return (magma_int_t)v;
    }
    return 1;
}

void magma_getdevices(magma_device_t *devices, magma_int_t size, magma_int_t *num_dev)
{
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    int k = 0;
    for (; k < ndev && k < size; ++k) devices[k] = k;
    *num_dev = k;
}

void magma_getdevice(magma_device_t *dev)
{
    int d = 0;
    cudaGetDevice(&d);
    *dev = d;
}

void magma_setdevice(magma_device_t dev) { cudaSetDevice(dev); }

magma_int_t magma_getdevice_arch(void)
{
    int d = 0, major = 0, minor = 0;
    cudaGetDevice(&d);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, d);
    return major * 100 + minor * 10;  // interface.cpp:215: 1000 for sm_100
}

magma_int_t magma_getdevice_multiprocessor_count(void)
{
    int d = 0, v = 0;
    cudaGetDevice(&d);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d);
    return v;
}

size_t magma_mem_size(magma_queue_t queue)
{
    (void)queue;
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return f;
}

// ---------------------------------------------------------------------------------------------
static magma_queue_t new_queue(magma_device_t device)
{
    magma_queue_t q = (magma_queue_t)calloc(1, sizeof(struct magma_queue));
    q->device = device;
    return q;
}

void magma_queue_create_internal(magma_device_t device, magma_queue_t *queue_ptr, const char *func,
                                 const char *file, int line)
{
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    magma_queue_t q = new_queue(device);
    cudaError_t e = cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess)
        fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(e), e, func,
                file, line);
    q->own_stream = true;
    *queue_ptr = q;
    cudaSetDevice(prev);
}

void magma_queue_create_from_cuda_internal(magma_device_t device, void *cuda_stream, void *cublas_handle,
                                           void *cusparse_handle, magma_queue_t *queue_ptr, const char *func,
                                           const char *file, int line)
{
    (void)cublas_handle; (void)cusparse_handle; (void)func; (void)file; (void)line;
    magma_queue_t q = new_queue(device);
    q->stream = (cudaStream_t)cuda_stream;  // NULL = the legacy default stream, as in the reference
    q->own_stream = false;
    *queue_ptr = q;
}

void magma_queue_destroy_internal(magma_queue_t q, const char *func, const char *file, int line)
{
    (void)func; (void)file; (void)line;
    if (!q) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(q->device);
    cudaStreamSynchronize(q->stream);
    if (q->aux_ready) {
        for (int i = 0; i < 2; ++i) cudaStreamDestroy(q->aux_stream[i]);
        for (int i = 0; i < 12; ++i) cudaEventDestroy(q->aux_event[i]);
    }
    for (int i = 0; i < 2; ++i)
        if (q->dscratch[i]) cudaFree(q->dscratch[i]);
    if (q->hscratch) cudaFreeHost(q->hscratch);
    if (q->own_stream) cudaStreamDestroy(q->stream);
    free(q);
    cudaSetDevice(prev);
}

void magma_queue_sync_internal(magma_queue_t q, const char *func, const char *file, int line)
{
    cudaError_t e = cudaStreamSynchronize(q ? q->stream : 0);
    if (e != cudaSuccess)
        fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(e), e, func,
                file, line);
}

magma_int_t magma_queue_get_device(magma_queue_t q) { return q->device; }
void *magma_queue_get_cuda_stream(magma_queue_t q) { return (void *)q->stream; }

// ---------------------------------------------------------------------------------------------
magma_int_t magma_malloc(magma_ptr *ptr_ptr, size_t bytes)
{
    if (bytes == 0) bytes = sizeof(double);  // alloc.cpp:64-66: malloc(0) still returns a pointer
    if (cudaMalloc(ptr_ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        *ptr_ptr = nullptr;
        return MAGMA_ERR_DEVICE_ALLOC;
    }
    return MAGMA_SUCCESS;
}

magma_int_t magma_free_internal(magma_ptr ptr, const char *func, const char *file, int line)
{
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(e), e, func,
                file, line);
        cudaGetLastError();
        return MAGMA_ERR_INVALID_PTR;
    }
    return MAGMA_SUCCESS;
}

magma_int_t magma_malloc_cpu(void **ptr_ptr, size_t bytes)
{
    if (bytes == 0) bytes = sizeof(double);
    if (posix_memalign(ptr_ptr, 64, bytes) != 0) {
        *ptr_ptr = nullptr;
        return MAGMA_ERR_HOST_ALLOC;
    }
    return MAGMA_SUCCESS;
}

magma_int_t magma_free_cpu(void *ptr)
{
    free(ptr);
    return MAGMA_SUCCESS;
}

magma_int_t magma_malloc_pinned(void **ptr_ptr, size_t bytes)
{
    if (bytes == 0) bytes = sizeof(double);
    if (cudaMallocHost(ptr_ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        *ptr_ptr = nullptr;
        return MAGMA_ERR_HOST_ALLOC;
    }
    return MAGMA_SUCCESS;
}

magma_int_t magma_free_pinned_internal(void *ptr, const char *func, const char *file, int line)
{
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(e), e, func,
                file, line);
        cudaGetLastError();
        return MAGMA_ERR_INVALID_PTR;
    }
    return MAGMA_SUCCESS;
}

magma_int_t magma_memset(void *ptr, int value, size_t count)
{
    return cudaMemset(ptr, value, count) == cudaSuccess ? MAGMA_SUCCESS : MAGMA_ERR_INVALID_PTR;
}

magma_int_t magma_memset_async(void *ptr, int value, size_t count, magma_queue_t queue)
{
    return cudaMemsetAsync(ptr, value, count, queue->stream) == cudaSuccess ? MAGMA_SUCCESS
                                                                             : MAGMA_ERR_INVALID_PTR;
}

// ---------------------------------------------------------------------------------------------
// copies. Vectors with inc != 1 go through a 2D copy (one element per "row"), as cublasSetVector.
static void copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                   cudaMemcpyKind kind, cudaStream_t s, bool sync, const char *func, const char *file, int line)
{
    if (width == 0 || height == 0) return;
    cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s);
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        fprintf(stderr, "CUDA runtime error: %s (%d) in %s at %s:%d\n", cudaGetErrorString(e), e, func,
                file, line);
}

#define QS(q) ((q) ? (q)->stream : (cudaStream_t)0)

void magma_setvector_internal(magma_int_t n, magma_int_t es, const void *hx, magma_int_t incx, magma_ptr dy,
                              magma_int_t incy, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(dy, (size_t)es * incy, hx, (size_t)es * incx, es, n, cudaMemcpyHostToDevice, QS(q), true, func, file, line);
}
void magma_getvector_internal(magma_int_t n, magma_int_t es, magma_const_ptr dx, magma_int_t incx, void *hy,
                              magma_int_t incy, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(hy, (size_t)es * incy, dx, (size_t)es * incx, es, n, cudaMemcpyDeviceToHost, QS(q), true, func, file, line);
}
void magma_setvector_async_internal(magma_int_t n, magma_int_t es, const void *hx, magma_int_t incx, magma_ptr dy,
                                    magma_int_t incy, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(dy, (size_t)es * incy, hx, (size_t)es * incx, es, n, cudaMemcpyHostToDevice, QS(q), false, func, file, line);
}
void magma_getvector_async_internal(magma_int_t n, magma_int_t es, magma_const_ptr dx, magma_int_t incx, void *hy,
                                    magma_int_t incy, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(hy, (size_t)es * incy, dx, (size_t)es * incx, es, n, cudaMemcpyDeviceToHost, QS(q), false, func, file, line);
}
void magma_setmatrix_internal(magma_int_t m, magma_int_t n, magma_int_t es, const void *hA, magma_int_t lda,
                              magma_ptr dB, magma_int_t lddb, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(dB, (size_t)es * lddb, hA, (size_t)es * lda, (size_t)es * m, n, cudaMemcpyHostToDevice, QS(q), true, func, file, line);
}
void magma_getmatrix_internal(magma_int_t m, magma_int_t n, magma_int_t es, magma_const_ptr dA, magma_int_t ldda,
                              void *hB, magma_int_t ldb, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(hB, (size_t)es * ldb, dA, (size_t)es * ldda, (size_t)es * m, n, cudaMemcpyDeviceToHost, QS(q), true, func, file, line);
}
void magma_setmatrix_async_internal(magma_int_t m, magma_int_t n, magma_int_t es, const void *hA, magma_int_t lda,
                                    magma_ptr dB, magma_int_t lddb, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(dB, (size_t)es * lddb, hA, (size_t)es * lda, (size_t)es * m, n, cudaMemcpyHostToDevice, QS(q), false, func, file, line);
}
void magma_getmatrix_async_internal(magma_int_t m, magma_int_t n, magma_int_t es, magma_const_ptr dA, magma_int_t ldda,
                                    void *hB, magma_int_t ldb, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(hB, (size_t)es * ldb, dA, (size_t)es * ldda, (size_t)es * m, n, cudaMemcpyDeviceToHost, QS(q), false, func, file, line);
}
void magma_copymatrix_internal(magma_int_t m, magma_int_t n, magma_int_t es, magma_const_ptr dA, magma_int_t ldda,
                               magma_ptr dB, magma_int_t lddb, magma_queue_t q, const char *func, const char *file, int line)
{
    copy2d(dB, (size_t)es * lddb, dA, (size_t)es * ldda, (size_t)es * m, n, cudaMemcpyDeviceToDevice, QS(q), true, func, file, line);
}

// ---------------------------------------------------------------------------------------------
const char *magma_strerror(magma_int_t error)
{
    switch (error) {
        case MAGMA_SUCCESS: return "success";
        case MAGMA_ERR: return "unknown error";
        case MAGMA_ERR_NOT_INITIALIZED: return "not initialized";
        case MAGMA_ERR_NOT_SUPPORTED: return "not supported";
        case MAGMA_ERR_HOST_ALLOC: return "cannot allocate memory on CPU host";
        case MAGMA_ERR_DEVICE_ALLOC: return "cannot allocate memory on GPU device";
        case MAGMA_ERR_INVALID_PTR: return "invalid pointer";
        case MAGMA_ERR_UNKNOWN: return "unknown error";
        case MAGMA_ERR_NOT_IMPLEMENTED: return "not implemented";
        case MAGMA_ERR_NAN: return "NaN detected";
        default: return error < 0 && error > MAGMA_ERR ? "invalid argument" : "unknown MAGMA error code";
    }
}

void magma_xerbla(const char *srname, magma_int_t neg_info)
{
    // same four cases and wording as control/xerbla.cpp:51-72
    if (neg_info < 0)
        fprintf(stderr, "Error in %s, function-specific error (info = %lld)\n", srname, (long long)-neg_info);
    else if (neg_info == 0)
        fprintf(stderr, "No error, why is %s calling xerbla? (info = %lld)\n", srname, (long long)-neg_info);
    else if (neg_info >= -MAGMA_ERR)
        fprintf(stderr, "Error in %s, %s (info = %lld)\n", srname, magma_strerror(-neg_info), (long long)-neg_info);
    else
        fprintf(stderr, "On entry to %s, parameter %lld had an illegal value (info = %lld)\n", srname,
                (long long)neg_info, (long long)-neg_info);
}

real_Double_t magma_wtime(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + t.tv_usec * 1e-6;
}

real_Double_t magma_sync_wtime(magma_queue_t queue)
{
    cudaStreamSynchronize(QS(queue));
    return magma_wtime();
}

// ---------------------------------------------------------------------------------------------
double *magma_doffset_1d(double *x, magma_int_t inc, magma_int_t i) { return x + (ptrdiff_t)(i - 1) * inc; }
magma_int_t *magma_ioffset_1d(magma_int_t *x, magma_int_t inc, magma_int_t i) { return x + (ptrdiff_t)(i - 1) * inc; }
double *magma_doffset_2d(double *A, magma_int_t lda, magma_int_t i, magma_int_t j)
{
    return A + (i - 1) + (ptrdiff_t)(j - 1) * lda;
}
magma_int_t *magma_ioffset_2d(magma_int_t *A, magma_int_t lda, magma_int_t i, magma_int_t j)
{
    return A + (i - 1) + (ptrdiff_t)(j - 1) * lda;
}

}  // extern "C"
#endif  // !MB200_INTERPOSE

extern "C" {
int64_t magma_b200_launch_count(void) { return g_launches.load(); }
void magma_b200_set_tier(int tier) { g_tier = tier; }
void magma_b200_set_small_rows(int rows) { g_small_rows = rows; }
void magma_b200_set_chain_panel(int on) { g_chain_panel = on; }
void magma_b200_set_fused_tail(int level) { g_fused_tail = level; }
void magma_b200_set_tall_panel(int on) { g_tall_panel2 = on; }
void magma_b200_set_split(int parts) { g_split = parts; }
}  // extern "C"
