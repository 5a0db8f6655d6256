/*
 * magma_b200.h -- C ABI of libmagma_b200.so: a B200-native (sm_100a) replacement for MAGMA's
 * batched FP64 LU factor-and-solve path. Every symbol below keeps the name, argument order,
 * argument meaning and error behaviour of the MAGMA 2.10.0 entry point it replaces (cited per
 * declaration as <file>:<line> under the reference tree; the reference ships z-masters that its
 * codegen turns into the d-names used here). Plain C: pointers and sizes only.
 *
 * Conventions (src/zgetrf_batched.cpp:44-65): all arrays -- matrices, pivot vectors, info, and
 * the pointer arrays themselves -- live in DEVICE memory; matrices are column-major; ipiv is
 * 1-based; info_array[b] = 0 or the 1-based column of the first exactly-zero pivot; a routine
 * returns 0, or -i after magma_xerbla() when its i-th argument is illegal. Work is enqueued on
 * the queue's CUDA stream. Unlike the reference (which blocks the host in getrf for n>32 and in
 * the vbatched driver) every fixed-size entry point here is fully asynchronous and allocates
 * nothing.
 */
#ifndef MAGMA_B200_H
#define MAGMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types (include/magma_types.h:67-72,99-100) ------------------------------------------- */
#ifdef MAGMA_ILP64
#error "libmagma_b200 is built LP64 only (magma_int_t = int), like the default MAGMA build"
#endif
typedef int magma_int_t;
typedef int magma_index_t;
typedef int magma_device_t;
typedef double real_Double_t;
typedef void *magma_ptr;
typedef const void *magma_const_ptr;
typedef magma_int_t *magmaInt_ptr;
typedef double *magmaDouble_ptr;
struct magma_queue;                       /* opaque */
typedef struct magma_queue *magma_queue_t;

/* ---- version: we stand in for MAGMA 2.10.0 (include/magma_types.h:527-529) ------------------ */
#define MAGMA_VERSION_MAJOR 2
#define MAGMA_VERSION_MINOR 10
#define MAGMA_VERSION_MICRO 0

/* ---- return codes (include/magma_types.h:548-567) ------------------------------------------ */
#define MAGMA_SUCCESS               0
#define MAGMA_ERR                  -100
#define MAGMA_ERR_NOT_INITIALIZED  -101
#define MAGMA_ERR_NOT_SUPPORTED    -103
#define MAGMA_ERR_HOST_ALLOC       -112
#define MAGMA_ERR_DEVICE_ALLOC     -113
#define MAGMA_ERR_INVALID_PTR      -115
#define MAGMA_ERR_UNKNOWN          -116
#define MAGMA_ERR_NOT_IMPLEMENTED  -117
#define MAGMA_ERR_NAN              -118

/* ---- LAPACK-style enums (include/magma_types.h:612-637) ------------------------------------ */
typedef enum { MagmaFalse = 0, MagmaTrue = 1 } magma_bool_t;
typedef enum { MagmaNoTrans = 111, MagmaTrans = 112, MagmaConjTrans = 113 } magma_trans_t;
typedef enum { MagmaUpper = 121, MagmaLower = 122, MagmaFull = 123 } magma_uplo_t;
typedef enum { MagmaNonUnit = 131, MagmaUnit = 132 } magma_diag_t;
typedef enum { MagmaLeft = 141, MagmaRight = 142 } magma_side_t;

/* ============================================================================================
 * Runtime boundary (standalone mode). Replaces interface_cuda/{interface,alloc,copy_v2,error}.cpp.
 * ============================================================================================ */
magma_int_t magma_init(void);                                   /* interface_cuda/interface.cpp:163 */
magma_int_t magma_finalize(void);                               /* interface.cpp:243 */
void magma_version(magma_int_t *major, magma_int_t *minor, magma_int_t *micro);
void magma_print_environment(void);                             /* interface.cpp:309 */

magma_int_t magma_num_gpus(void);                               /* control/auxiliary.cpp:46 (MAGMA_NUM_GPUS) */
void magma_getdevices(magma_device_t *devices, magma_int_t size, magma_int_t *num_dev);
void magma_getdevice(magma_device_t *dev);
void magma_setdevice(magma_device_t dev);
magma_int_t magma_getdevice_arch(void);                         /* interface.cpp:528 (e.g. 1000) */
magma_int_t magma_getdevice_multiprocessor_count(void);
size_t magma_mem_size(magma_queue_t queue);

/* queues: include/magma_auxiliary.h:259-275 are macros over these _internal symbols */
void magma_queue_create_internal(magma_device_t device, magma_queue_t *queue_ptr,
                                 const char *func, const char *file, int line);
/* stream is a cudaStream_t; cublas / cusparse handles are accepted and ignored (no vendor BLAS
 * on this path). interface.cpp:1008 */
void magma_queue_create_from_cuda_internal(magma_device_t device, void *cuda_stream,
                                           void *cublas_handle, void *cusparse_handle,
                                           magma_queue_t *queue_ptr,
                                           const char *func, const char *file, int line);
void magma_queue_destroy_internal(magma_queue_t queue, const char *func, const char *file, int line);
void magma_queue_sync_internal(magma_queue_t queue, const char *func, const char *file, int line);
magma_int_t magma_queue_get_device(magma_queue_t queue);        /* interface.cpp:798 */
void *magma_queue_get_cuda_stream(magma_queue_t queue);         /* interface.cpp:815 (cudaStream_t) */

#define magma_queue_create(device, queue_ptr) \
    magma_queue_create_internal(device, queue_ptr, __func__, __FILE__, __LINE__)
#define magma_queue_create_from_cuda(device, stream, cublas, cusparse, queue_ptr) \
    magma_queue_create_from_cuda_internal(device, stream, cublas, cusparse, queue_ptr, __func__, __FILE__, __LINE__)
#define magma_queue_destroy(queue) magma_queue_destroy_internal(queue, __func__, __FILE__, __LINE__)
#define magma_queue_sync(queue) magma_queue_sync_internal(queue, __func__, __FILE__, __LINE__)

/* memory: interface_cuda/alloc.cpp:61-110,320-335 */
magma_int_t magma_malloc(magma_ptr *ptr_ptr, size_t bytes);
magma_int_t magma_malloc_cpu(void **ptr_ptr, size_t bytes);
magma_int_t magma_malloc_pinned(void **ptr_ptr, size_t bytes);
magma_int_t magma_free_internal(magma_ptr ptr, const char *func, const char *file, int line);
magma_int_t magma_free_cpu(void *ptr);
magma_int_t magma_free_pinned_internal(void *ptr, const char *func, const char *file, int line);
magma_int_t magma_memset(void *ptr, int value, size_t count);
magma_int_t magma_memset_async(void *ptr, int value, size_t count, magma_queue_t queue);
#define magma_free(ptr) magma_free_internal(ptr, __func__, __FILE__, __LINE__)
#define magma_free_pinned(ptr) magma_free_pinned_internal(ptr, __func__, __FILE__, __LINE__)
static inline magma_int_t magma_imalloc(magmaInt_ptr *p, size_t n) { return magma_malloc((magma_ptr *)p, n * sizeof(magma_int_t)); }
static inline magma_int_t magma_dmalloc(magmaDouble_ptr *p, size_t n) { return magma_malloc((magma_ptr *)p, n * sizeof(double)); }
static inline magma_int_t magma_imalloc_cpu(magma_int_t **p, size_t n) { return magma_malloc_cpu((void **)p, n * sizeof(magma_int_t)); }
static inline magma_int_t magma_dmalloc_cpu(double **p, size_t n) { return magma_malloc_cpu((void **)p, n * sizeof(double)); }
static inline magma_int_t magma_imalloc_pinned(magma_int_t **p, size_t n) { return magma_malloc_pinned((void **)p, n * sizeof(magma_int_t)); }
static inline magma_int_t magma_dmalloc_pinned(double **p, size_t n) { return magma_malloc_pinned((void **)p, n * sizeof(double)); }

/* copies: interface_cuda/copy_v2.cpp; include/magma_copy.h:45-160 */
void magma_setvector_internal(magma_int_t n, magma_int_t elemSize, const void *hx_src, magma_int_t incx,
                              magma_ptr dy_dst, magma_int_t incy, magma_queue_t queue,
                              const char *func, const char *file, int line);
void magma_getvector_internal(magma_int_t n, magma_int_t elemSize, magma_const_ptr dx_src, magma_int_t incx,
                              void *hy_dst, magma_int_t incy, magma_queue_t queue,
                              const char *func, const char *file, int line);
void magma_setvector_async_internal(magma_int_t n, magma_int_t elemSize, const void *hx_src, magma_int_t incx,
                                    magma_ptr dy_dst, magma_int_t incy, magma_queue_t queue,
                                    const char *func, const char *file, int line);
void magma_getvector_async_internal(magma_int_t n, magma_int_t elemSize, magma_const_ptr dx_src, magma_int_t incx,
                                    void *hy_dst, magma_int_t incy, magma_queue_t queue,
                                    const char *func, const char *file, int line);
void magma_setmatrix_internal(magma_int_t m, magma_int_t n, magma_int_t elemSize, const void *hA_src, magma_int_t lda,
                              magma_ptr dB_dst, magma_int_t lddb, magma_queue_t queue,
                              const char *func, const char *file, int line);
void magma_getmatrix_internal(magma_int_t m, magma_int_t n, magma_int_t elemSize, magma_const_ptr dA_src, magma_int_t ldda,
                              void *hB_dst, magma_int_t ldb, magma_queue_t queue,
                              const char *func, const char *file, int line);
void magma_setmatrix_async_internal(magma_int_t m, magma_int_t n, magma_int_t elemSize, const void *hA_src, magma_int_t lda,
                                    magma_ptr dB_dst, magma_int_t lddb, magma_queue_t queue,
                                    const char *func, const char *file, int line);
void magma_getmatrix_async_internal(magma_int_t m, magma_int_t n, magma_int_t elemSize, magma_const_ptr dA_src, magma_int_t ldda,
                                    void *hB_dst, magma_int_t ldb, magma_queue_t queue,
                                    const char *func, const char *file, int line);
void magma_copymatrix_internal(magma_int_t m, magma_int_t n, magma_int_t elemSize, magma_const_ptr dA_src, magma_int_t ldda,
                               magma_ptr dB_dst, magma_int_t lddb, magma_queue_t queue,
                               const char *func, const char *file, int line);
#define magma_setvector(n, es, hx, incx, dy, incy, q) magma_setvector_internal(n, es, hx, incx, dy, incy, q, __func__, __FILE__, __LINE__)
#define magma_getvector(n, es, dx, incx, hy, incy, q) magma_getvector_internal(n, es, dx, incx, hy, incy, q, __func__, __FILE__, __LINE__)
#define magma_setmatrix(m, n, es, hA, lda, dB, lddb, q) magma_setmatrix_internal(m, n, es, hA, lda, dB, lddb, q, __func__, __FILE__, __LINE__)
#define magma_getmatrix(m, n, es, dA, ldda, hB, ldb, q) magma_getmatrix_internal(m, n, es, dA, ldda, hB, ldb, q, __func__, __FILE__, __LINE__)
#define magma_dsetmatrix(m, n, hA, lda, dB, lddb, q) magma_setmatrix_internal(m, n, sizeof(double), hA, lda, dB, lddb, q, __func__, __FILE__, __LINE__)
#define magma_dgetmatrix(m, n, dA, ldda, hB, ldb, q) magma_getmatrix_internal(m, n, sizeof(double), dA, ldda, hB, ldb, q, __func__, __FILE__, __LINE__)
#define magma_dsetvector(n, hx, incx, dy, incy, q) magma_setvector_internal(n, sizeof(double), hx, incx, dy, incy, q, __func__, __FILE__, __LINE__)
#define magma_dgetvector(n, dx, incx, hy, incy, q) magma_getvector_internal(n, sizeof(double), dx, incx, hy, incy, q, __func__, __FILE__, __LINE__)
#define magma_isetvector(n, hx, incx, dy, incy, q) magma_setvector_internal(n, sizeof(magma_int_t), hx, incx, dy, incy, q, __func__, __FILE__, __LINE__)
#define magma_igetvector(n, dx, incx, hy, incy, q) magma_getvector_internal(n, sizeof(magma_int_t), dx, incx, hy, incy, q, __func__, __FILE__, __LINE__)

/* errors and timing: control/xerbla.cpp:51-72, interface_cuda/error.cpp:140-260, control/magma_timer.cpp */
void magma_xerbla(const char *srname, magma_int_t neg_info);
const char *magma_strerror(magma_int_t error);
real_Double_t magma_wtime(void);
real_Double_t magma_sync_wtime(magma_queue_t queue);

/* pointer-array helpers: magmablas/zset_pointer.cu:86-98,224-231; magmablas/set_pointer.cu
 * output_array[b] = input + b*batch_offset + row + column*lda */
void magma_dset_pointer(double **output_array, double *input, magma_int_t lda, magma_int_t row,
                        magma_int_t column, magma_int_t batch_offset, magma_int_t batchCount,
                        magma_queue_t queue);
void magma_iset_pointer(magma_int_t **output_array, magma_int_t *input, magma_int_t lda,
                        magma_int_t row, magma_int_t column, magma_int_t batchSize,
                        magma_int_t batchCount, magma_queue_t queue);
/* output_array[b] = input_array[b] + row + column*lda */
void magma_ddisplace_pointers(double **output_array, double **input_array, magma_int_t lda,
                              magma_int_t row, magma_int_t column, magma_int_t batchCount,
                              magma_queue_t queue);
void magma_idisplace_pointers(magma_int_t **output_array, magma_int_t **input_array, magma_int_t lda,
                              magma_int_t row, magma_int_t column, magma_int_t batchCount,
                              magma_queue_t queue);

/* Fortran helpers, 1-based: fortran/offset.c */
double *magma_doffset_1d(double *x, magma_int_t inc, magma_int_t i);
magma_int_t *magma_ioffset_1d(magma_int_t *x, magma_int_t inc, magma_int_t i);
double *magma_doffset_2d(double *A, magma_int_t lda, magma_int_t i, magma_int_t j);
magma_int_t *magma_ioffset_2d(magma_int_t *A, magma_int_t lda, magma_int_t i, magma_int_t j);

/* ============================================================================================
 * Batched LU: the hot path.
 * ============================================================================================ */

/* A_b = P_b L_b U_b for b < batchCount; general m x n.   src/zgetrf_batched.cpp:81-213
 * errors: -1 (m<0), -2 (n<0), -4 (ldda<max(1,m)); quick return if m==0 || n==0. */
magma_int_t magma_dgetrf_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ldda,
                                 magma_int_t **ipiv_array, magma_int_t *info_array,
                                 magma_int_t batchCount, magma_queue_t queue);

/* op(A_b) X_b = B_b from the factors.   src/zgetrs_batched.cpp:85-181
 * errors: -1 trans, -2 n, -3 nrhs, -5 ldda, -8 lddb. MagmaTrans/MagmaConjTrans follow LAPACK
 * dgetrs (the reference's transposed branch is defective). */
magma_int_t magma_dgetrs_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs,
                                 double **dA_array, magma_int_t ldda, magma_int_t **dipiv_array,
                                 double **dB_array, magma_int_t lddb, magma_int_t batchCount,
                                 magma_queue_t queue);

/* A_b X_b = B_b; A overwritten by LU, B by X.   src/zgesv_batched.cpp:92-154
 * errors: -1 n, -2 nrhs, -4 ldda, -6 lddb. */
magma_int_t magma_dgesv_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                magma_int_t **dipiv_array, double **dB_array, magma_int_t lddb,
                                magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue);

/* inv(A_b) from the factors of magma_dgetrf_batched, out of place (dA_array is read only here; the reference
 * documents it as in/out but never writes it).   src/zgetri_outofplace_batched.cpp:81-141,
 * prototype include/magma_zbatched.h:871-878. errors: -1 n, -3 ldda, -6 lddia. info_array is not written
 * (as in the reference). Asynchronous on the queue (the reference ends with magma_queue_sync). */
magma_int_t magma_dgetri_outofplace_batched(magma_int_t n, double **dA_array, magma_int_t ldda,
                                            magma_int_t **dipiv_array, double **dinvA_array,
                                            magma_int_t lddia, magma_int_t *info_array,
                                            magma_int_t batchCount, magma_queue_t queue);

/* LU without pivoting and its solves.   src/zgetrf_nopiv_batched.cpp:75-170,
 * src/zgetrs_nopiv_batched.cpp:85-180, src/zgesv_nopiv_batched.cpp:85-130; prototypes include/magma_zbatched.h:976-998.
 * getrf: errors -1 m, -2 n, -4 ldda; info_array[b] = first i with U(i,i) == 0 (1-based), else 0. Unlike the reference,
 * which stops factoring a matrix at its first zero diagonal, the factorisation is completed LAPACK-style (the zero
 * pivot's column stays unscaled). At most 512 rows (MAGMA_ERR_NOT_SUPPORTED beyond).
 * getrs: errors -1 trans, -2 n, -3 nrhs, -5 ldda, -8 lddb; gesv: -1 n, -2 nrhs, -4 ldda, -6 lddb. */
magma_int_t magma_dgetrf_nopiv_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ldda,
                                       magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgetrs_nopiv_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, double **dA_array,
                                       magma_int_t ldda, double **dB_array, magma_int_t lddb,
                                       magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgesv_nopiv_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                      double **dB_array, magma_int_t lddb, magma_int_t *info_array,
                                      magma_int_t batchCount, magma_queue_t queue);

/* Variable sizes; m, n, ldda are DEVICE arrays of length batchCount.
 * src/zgetrf_vbatched.cpp:340-398 (checker + setup + workspace inside, blocks the host). */
magma_int_t magma_dgetrf_vbatched(magma_int_t *m, magma_int_t *n, double **dA_array, magma_int_t *ldda,
                                  magma_int_t **ipiv_array, magma_int_t *info_array,
                                  magma_int_t batchCount, magma_queue_t queue);

/* ---- s / c / z precisions of the four entry points -------------------------------------------------
 * The reference generates them from the z masters (src/zgetrf_batched.cpp:11 "@precisions normal z -> s d c");
 * signatures: include/magma_zbatched.h:848 (getrf), :464 (getrs), :1009 (gesv), include/magma_zvbatched.h:57 (getrf_vbatched)
 * and their generated s / c counterparts. One templated implementation (csrc/lu_scz.cu), bit-identical to
 * oracle/lu_oracle_scz.c; complex pivoting uses |re| + |im| like the reference's device code. */
typedef struct { float x, y; } magmaFloatComplex;   /* layout of cuFloatComplex (include/magma_types.h:100) */
typedef struct { double x, y; } magmaDoubleComplex; /* layout of cuDoubleComplex */
magma_int_t magma_sgetrf_batched(magma_int_t m, magma_int_t n, float **dA_array, magma_int_t ldda,
                                 magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                 magma_queue_t queue);
magma_int_t magma_sgetrs_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, float **dA_array,
                                 magma_int_t ldda, magma_int_t **dipiv_array, float **dB_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_sgesv_batched(magma_int_t n, magma_int_t nrhs, float **dA_array, magma_int_t ldda,
                                magma_int_t **dipiv_array, float **dB_array, magma_int_t lddb,
                                magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_sgetrf_vbatched(magma_int_t *m, magma_int_t *n, float **dA_array, magma_int_t *ldda,
                                  magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                  magma_queue_t queue);
magma_int_t magma_cgetrf_batched(magma_int_t m, magma_int_t n, magmaFloatComplex **dA_array, magma_int_t ldda,
                                 magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                 magma_queue_t queue);
magma_int_t magma_cgetrs_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, magmaFloatComplex **dA_array,
                                 magma_int_t ldda, magma_int_t **dipiv_array, magmaFloatComplex **dB_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_cgesv_batched(magma_int_t n, magma_int_t nrhs, magmaFloatComplex **dA_array, magma_int_t ldda,
                                magma_int_t **dipiv_array, magmaFloatComplex **dB_array, magma_int_t lddb,
                                magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_cgetrf_vbatched(magma_int_t *m, magma_int_t *n, magmaFloatComplex **dA_array, magma_int_t *ldda,
                                  magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                  magma_queue_t queue);
magma_int_t magma_zgetrf_batched(magma_int_t m, magma_int_t n, magmaDoubleComplex **dA_array, magma_int_t ldda,
                                 magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                 magma_queue_t queue);
magma_int_t magma_zgetrs_batched(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, magmaDoubleComplex **dA_array,
                                 magma_int_t ldda, magma_int_t **dipiv_array, magmaDoubleComplex **dB_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_zgesv_batched(magma_int_t n, magma_int_t nrhs, magmaDoubleComplex **dA_array, magma_int_t ldda,
                                magma_int_t **dipiv_array, magmaDoubleComplex **dB_array, magma_int_t lddb,
                                magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_zgetrf_vbatched(magma_int_t *m, magma_int_t *n, magmaDoubleComplex **dA_array, magma_int_t *ldda,
                                  magma_int_t **ipiv_array, magma_int_t *info_array, magma_int_t batchCount,
                                  magma_queue_t queue);

/* Expert forms.   src/zgetrf_vbatched.cpp:223-336 and :19-130
 * _work: lwork[0] < 0 is a workspace query (required bytes returned in lwork[0]); asynchronous. */
magma_int_t magma_dgetrf_vbatched_max_nocheck_work(
    magma_int_t *m, magma_int_t *n, magma_int_t max_m, magma_int_t max_n, magma_int_t max_minmn,
    magma_int_t max_mxn, double **dA_array, magma_int_t *ldda, magma_int_t **dipiv_array,
    magma_int_t *info_array, void *work, magma_int_t *lwork, magma_int_t batchCount,
    magma_queue_t queue);
/* minmn and pivinfo_array are accepted for signature compatibility; this implementation needs
 * neither (no pivinfo scratch anywhere on the path). nb/recnb are accepted as hints. */
magma_int_t magma_dgetrf_vbatched_max_nocheck(
    magma_int_t *m, magma_int_t *n, magma_int_t *minmn, magma_int_t max_m, magma_int_t max_n,
    magma_int_t max_minmn, magma_int_t max_mxn, magma_int_t nb, magma_int_t recnb,
    double **dA_array, magma_int_t *ldda, magma_int_t **ipiv_array, magma_int_t **pivinfo_array,
    magma_int_t *info_array, magma_int_t batchCount, magma_queue_t queue);

/* Internal-but-public entry points kept for source compatibility with callers/testers. */
/* n <= 32 square, one kernel.   magmablas/zgetrf_batched_smallsq_noshfl.cu:196-285 */
magma_int_t magma_dgetrf_batched_smallsq_noshfl(magma_int_t n, double **dA_array, magma_int_t ldda,
                                                magma_int_t **ipiv_array, magma_int_t *info_array,
                                                magma_int_t batchCount, magma_queue_t queue);
/* fused factor+solve; returns 0 if it ran, -100 if the shape is outside the fused kernel's range
 * (then the caller falls back to getrf+getrs).   magmablas/zgesv_batched_small.cu:375-493 */
magma_int_t magma_dgesv_batched_small(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                      magma_int_t **dipiv_array, double **dB_array, magma_int_t lddb,
                                      magma_int_t *dinfo_array, magma_int_t batchCount, magma_queue_t queue);
/* LAPACK-order row interchanges k1..k2 (1-based, inclusive) on n columns of each matrix.
 * magmablas/zlaswp_batched.cu:163-206 */
void magma_dlaswp_rowserial_batched(magma_int_t n, double **dA_array, magma_int_t lda, magma_int_t k1,
                                    magma_int_t k2, magma_int_t **ipiv_array, magma_int_t batchCount,
                                    magma_queue_t queue);
/* Solve with a random butterfly transformation instead of pivoting: A <- U^T A V, B <- U^T B, LU without pivoting,
 * triangular solves, X <- V Y. U, V: two-level butterflies of 2n scalars each, drawn with rand() as the reference does
 * when gen = MagmaTrue.   src/zgesv_rbt_batched.cpp:81-166, src/zgerbt_batched.cpp:118-181,
 * magmablas/zgerbt_func_batched.cu:57-210 */
magma_int_t magma_dgesv_rbt_batched(magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                    double **dB_array, magma_int_t lddb, magma_int_t *dinfo_array, magma_int_t batchCount,
                                    magma_queue_t queue);
magma_int_t magma_dgerbt_batched(magma_bool_t gen, magma_int_t n, magma_int_t nrhs, double **dA_array, magma_int_t ldda,
                                 double **dB_array, magma_int_t lddb, double *U, double *V, magma_int_t *info,
                                 magma_int_t batchCount, magma_queue_t queue);
void magmablas_dprbt_batched(magma_int_t n, double **dA_array, magma_int_t ldda, double *du, double *dv,
                             magma_int_t batchCount, magma_queue_t queue);
void magmablas_dprbt_mv_batched(magma_int_t n, magma_int_t nrhs, double *dv, double **db_array, magma_int_t lddb,
                                magma_int_t batchCount, magma_queue_t queue);
void magmablas_dprbt_mtv_batched(magma_int_t n, magma_int_t nrhs, double *du, double **db_array, magma_int_t lddb,
                                 magma_int_t batchCount, magma_queue_t queue);
/* Panel-level entry points of the reference, kept for source compatibility (include/magma_zbatched.h:829-855).
 * Each factors the m x n block at (ai, aj) of every matrix: pivots are 1-based RELATIVE to row ai and written at
 * ipiv_array[b] + ai; a zero pivot at panel step i records gbstep + i + 1 in info_array[b] unless an earlier panel
 * already recorded one (gbstep == 0 resets it). dpivinfo_array / min_recpnb steer the reference's implementation
 * and are ignored here; the work runs on the same kernels as magma_dgetrf_batched.
 *   magma_dgetf2_fused_batched     magmablas/zgetf2_kernels.cu:1005-1058 (n <= 32, gbstep = aj)
 *   magma_dgetf2_batched           src/zgetf2_batched.cpp:243-287
 *   magma_dgetrf_recpanel_batched  src/zgetrf_panel_batched.cpp:101-196 */
magma_int_t magma_dgetf2_fused_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ai, magma_int_t aj,
                                       magma_int_t ldda, magma_int_t **dipiv_array, magma_int_t *info_array,
                                       magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgetf2_batched(magma_int_t m, magma_int_t n, double **dA_array, magma_int_t ai, magma_int_t aj,
                                 magma_int_t lda, magma_int_t **ipiv_array, magma_int_t **dpivinfo_array,
                                 magma_int_t *info_array, magma_int_t gbstep, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgetrf_recpanel_batched(magma_int_t m, magma_int_t n, magma_int_t min_recpnb, double **dA_array,
                                          magma_int_t ai, magma_int_t aj, magma_int_t ldda, magma_int_t **dipiv_array,
                                          magma_int_t **dpivinfo_array, magma_int_t *info_array, magma_int_t gbstep,
                                          magma_int_t batchCount, magma_queue_t queue);
/* Row permutation by a pivinfo table: new[r] = old[pivinfo[r]-1] for the k2-k1 top rows (written to output) and for
 * the rows they came from (written in place in input).   magmablas/zlaswp_batched.cu:47-87 */
void magma_dlaswp_rowparallel_batched(magma_int_t n, double **input_array, magma_int_t input_i, magma_int_t input_j,
                                      magma_int_t ldi, double **output_array, magma_int_t output_i, magma_int_t output_j,
                                      magma_int_t ldo, magma_int_t k1, magma_int_t k2, magma_int_t **pivinfo_array,
                                      magma_int_t batchCount, magma_queue_t queue);
/* x_b <- op(A_b)^-1 x_b in place, increment incb > 0.   magmablas/ztrsv_batched.cu:258-298 */
void magmablas_dtrsv_batched(magma_uplo_t uplo, magma_trans_t transA, magma_diag_t diag, magma_int_t n, double **dA_array,
                             magma_int_t ldda, double **dB_array, magma_int_t incb, magma_int_t batchCount,
                             magma_queue_t queue);
/* side = Left: B_b <- alpha op(A_b)^-1 B_b; side = Right: B_b <- alpha B_b op(A_b)^-1.   magmablas/ztrsm_batched_core.cpp:299-350 */
void magmablas_dtrsm_batched(magma_side_t side, magma_uplo_t uplo, magma_trans_t transA, magma_diag_t diag,
                             magma_int_t m, magma_int_t n, double alpha, double **dA_array, magma_int_t ldda,
                             double **dB_array, magma_int_t lddb, magma_int_t batchCount, magma_queue_t queue);
/* C_b <- alpha op(A_b) op(B_b) + beta C_b on the FP64 tensor pipe, any transposes.  magmablas/zgemm_batched.cpp:269-286 */
void magma_dgemm_batched(magma_trans_t transA, magma_trans_t transB, magma_int_t m, magma_int_t n, magma_int_t k,
                         double alpha, double const *const *dA_array, magma_int_t ldda, double const *const *dB_array,
                         magma_int_t lddb, double beta, double **dC_array, magma_int_t lddc, magma_int_t batchCount,
                         magma_queue_t queue);
/* Strided device front ends (new surface): matrix b at dA + b*strideA (elements), pivots at
 * dipiv + b*stride_piv, right-hand sides at dB + b*strideB. Same results and return codes as the pointer-array forms. */
magma_int_t magma_dgetrf_batched_strided(magma_int_t m, magma_int_t n, double *dA, magma_int_t ldda, magma_int_t strideA,
                                         magma_int_t *dipiv, magma_int_t stride_piv, magma_int_t *info_array,
                                         magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgetrs_batched_strided(magma_trans_t trans, magma_int_t n, magma_int_t nrhs, double *dA, magma_int_t ldda,
                                         magma_int_t strideA, magma_int_t *dipiv, magma_int_t stride_piv, double *dB,
                                         magma_int_t lddb, magma_int_t strideB, magma_int_t batchCount, magma_queue_t queue);
magma_int_t magma_dgesv_batched_strided(magma_int_t n, magma_int_t nrhs, double *dA, magma_int_t ldda, magma_int_t strideA,
                                        magma_int_t *dipiv, magma_int_t stride_piv, double *dB, magma_int_t lddb,
                                        magma_int_t strideB, magma_int_t *dinfo_array, magma_int_t batchCount,
                                        magma_queue_t queue);
/* C_b(Ci.., Cj..) <- alpha op(A_b)(Ai.., Aj..) op(B_b)(Bi.., Bj..) + beta C_b.   magmablas/zgemm_batched.cpp:49-100 */
void magma_dgemm_batched_core(magma_trans_t transA, magma_trans_t transB, magma_int_t m, magma_int_t n, magma_int_t k,
                              double alpha, double const *const *dA_array, magma_int_t Ai, magma_int_t Aj, magma_int_t ldda,
                              double const *const *dB_array, magma_int_t Bi, magma_int_t Bj, magma_int_t lddb,
                              double beta, double **dC_array, magma_int_t Ci, magma_int_t Cj, magma_int_t lddc,
                              magma_int_t batchCount, magma_queue_t queue);

/* Tuning tables re-derived for B200.  control/get_batched_crossover.cpp:300-305,336-342,918-928;
 * control/get_ntcol.cpp:197-210 */
void magma_get_dgetrf_batched_nbparam(magma_int_t n, magma_int_t *nb, magma_int_t *recnb);
void magma_get_dgetrf_vbatched_nbparam(magma_int_t max_m, magma_int_t max_n, magma_int_t *nb, magma_int_t *recnb);
magma_int_t magma_get_dgetrf_batched_ntcol(magma_int_t m, magma_int_t n);
magma_int_t magma_get_dtrsm_batched_stop_nb(magma_side_t side, magma_int_t m, magma_int_t n);

/* ============================================================================================
 * Additions (absent from the reference; prefixed magma_b200_ or magmaf_).
 * ============================================================================================ */

/* Multi-GPU sharding by matrix index, no collective: shard g (device of queues[g]) owns
 * batchCount[g] matrices through its own device-resident arrays. SURVEY.md section 8e. */
magma_int_t magma_b200_dgetrf_batched_mgpu(magma_int_t ngpu, magma_int_t m, magma_int_t n,
                                           double ***dA_array, magma_int_t ldda, magma_int_t ***ipiv_array,
                                           magma_int_t **info_array, const magma_int_t *batchCount,
                                           magma_queue_t *queues);
magma_int_t magma_b200_dgesv_batched_mgpu(magma_int_t ngpu, magma_int_t n, magma_int_t nrhs,
                                          double ***dA_array, magma_int_t ldda, magma_int_t ***dipiv_array,
                                          double ***dB_array, magma_int_t lddb, magma_int_t **dinfo_array,
                                          const magma_int_t *batchCount, magma_queue_t *queues);

/* Host-buffer front ends (strided batches in pageable or pinned HOST memory): chunked H2D ->
 * factor/solve -> D2H, double-buffered on internal streams. lda-strided, matrices back to back
 * (stride lda*n), ipiv stride min(m,n). Blocks until the results are in host memory. */
magma_int_t magma_b200_dgetrf_batched_host(magma_int_t m, magma_int_t n, double *hA, magma_int_t lda,
                                           magma_int_t *hipiv, magma_int_t *hinfo, magma_int_t batchCount,
                                           magma_queue_t queue);
magma_int_t magma_b200_dgesv_batched_host(magma_int_t n, magma_int_t nrhs, double *hA, magma_int_t lda,
                                          magma_int_t *hipiv, double *hB, magma_int_t ldb, magma_int_t *hinfo,
                                          magma_int_t batchCount, magma_queue_t queue);

/* Device-side dlarnv(idist=1): fills dx[0..n) with the same stream LAPACK's dlarnv produces from
 * iseed (host array of 4, updated on return), so benches/tests can synthesise the testers'
 * inputs (testing/testing_zgetrf_batched.cpp:136,180) directly in HBM. */
void magma_b200_dlarnv_uniform(magma_int_t *iseed, int64_t n, double *dx, magma_queue_t queue);

/* Microbenchmarks used for the roofline denominators; return TFLOP/s or GB/s measured with CUDA
 * events on the queue's stream. kind: 0 = DFMA (vector FP64 pipe), 1 = DMMA m8n8k4 (FP64 tensor). */
double magma_b200_fp64_peak_tflops(int kind, magma_queue_t queue);
double magma_b200_hbm_copy_gbs(size_t bytes, magma_queue_t queue);

/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t magma_b200_launch_count(void);
/* Force a tier for tests/benches: 0 = auto, 1 = register/warp (small), 2 = blocked everywhere,
 * 4 = blocked tier and getrs on the DFMA kernels only (no tensor pipe), 5 = no 64-wide pairing,
 * 6 = right-looking blocked driver everywhere, 7 = left-looking slab driver up to 512 rows (default: 448). */
void magma_b200_set_tier(int tier);
/* B200 tier boundaries, the same table the drivers dispatch on: which = 0 register tier (max(m,n) <=), 1 register-file
 * tier (<=), 2 left-looking slab driver (rows <=), 3 single-launch shared-memory tier (<=; 0 = off).
 * Counterpart of control/get_batched_crossover.cpp for this path. */
magma_int_t magma_b200_get_dgetrf_batched_crossover(magma_int_t which);
/* Self-test of the inline reciprocal of lu_fused.cu: mismatches against IEEE 1.0/x over n pseudo-random inputs (0 expected). */
int64_t magma_b200_rcp_selftest(int64_t n, magma_queue_t queue);
/* Level L = 1..4: 32-column panels of (128 - 32 L, 128] rows run on single-warp pivot chains (panel_chain_kernel);
 * default 3 (33..128 rows). 0: the one-thread-per-row panel kernel everywhere (A/B runs). */
void magma_b200_set_chain_panel(int level);
/* Matrices of at most 128 rows in the left-looking driver: 1 = panels of 33..96 rows are factored in the tail of the slab
 * kernel that updated them (no panel launch, no slab round trip), 2 = the last <= 32-row panel too, 0 = off (default:
 * four chains per SM in the slab kernel lose to the 8..12 of the panel kernels, n = 128: 8.99 -> 9.69 ms). */
void magma_b200_set_fused_tail(int level);
/* 1 (default): panels of 129..512 rows in the left-looking driver are factored in two 16-column halves per thread
 * (twice the pivot chains per SM); 0: the 32-column register panel kernel (A/B runs). */
void magma_b200_set_tall_panel(int on);
/* Left-looking driver: the batch is cut into `parts` slices (1..4) that run their kernel chains on separate streams, forked
 * from and joined to the queue's stream, so that the tail of every kernel is filled by another slice's work. 0 (default):
 * two slices for matrices of more than 256 rows (n = 512: 25.45 -> 24.76 ms), one below (where it costs 1-2%); 1 = off. */
void magma_b200_set_split(int parts);
/* 1 (default): magma_dgetri_outofplace_batched runs its single-launch kernel for n <= 64; 0: identity + getrs for every n. */
void magma_b200_set_getri_fused(int on);
/* Largest max(m,n) routed to the single-launch shared-memory tier (lu_fused.cu), 0..128; 0 disables it (A/B runs). */
void magma_b200_set_fused_max(int n);
/* Largest max(m,n) routed to the register-file tier (lu_mid.cu), 32..128; for tuning sweeps. */
void magma_b200_set_mid_max(int n);
/* Register tier layout override for tuning sweeps: rows per lane (1 or 2), 0 = tuned default; 3/4 = alternative
 * square kernels (n = 32 staged / n = 16 with 4 CTAs per SM), 5/6 = n = 32 on the shuffle-broadcast square kernel with 16
 * lanes x 2 rows per matrix (3 / 2 CTAs per SM; 7.5 ms per 10^6 against 6.2 ms), 7 = keep the register-file tier on 65..96
 * (default: left-looking blocked driver there), 8 = single-phase 16-warp register-file kernel, 9 = generic kernel only. */
void magma_b200_set_small_rows(int rows);

/* F77-style by-reference wrappers in the control/magma_df77.cpp convention (device pointers
 * passed as integer handles). New surface: the generated reference layer has no batched LU. */
typedef size_t devptr_t;
void magmaf_dgetrf_batched_(magma_int_t *m, magma_int_t *n, devptr_t *dA_array, magma_int_t *ldda,
                            devptr_t *ipiv_array, devptr_t *info_array, magma_int_t *batchCount,
                            devptr_t *queue, magma_int_t *info);
void magmaf_dgetrs_batched_(const char *trans, magma_int_t *n, magma_int_t *nrhs, devptr_t *dA_array,
                            magma_int_t *ldda, devptr_t *dipiv_array, devptr_t *dB_array, magma_int_t *lddb,
                            magma_int_t *batchCount, devptr_t *queue, magma_int_t *info);
void magmaf_dgesv_batched_(magma_int_t *n, magma_int_t *nrhs, devptr_t *dA_array, magma_int_t *ldda,
                           devptr_t *dipiv_array, devptr_t *dB_array, magma_int_t *lddb, devptr_t *dinfo_array,
                           magma_int_t *batchCount, devptr_t *queue, magma_int_t *info);
void magmaf_dgetrf_vbatched_(devptr_t *m, devptr_t *n, devptr_t *dA_array, devptr_t *ldda,
                             devptr_t *ipiv_array, devptr_t *info_array, magma_int_t *batchCount,
                             devptr_t *queue, magma_int_t *info);

#ifdef __cplusplus
}
#endif
#endif /* MAGMA_B200_H */
