/* TEST INFRASTRUCTURE -- CPU restatement of the reference's batched LU path in the other three precisions
 * (single, complex single, complex double). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 * One body (lu_oracle_tmpl.h), instantiated three times -- the way the reference generates s / c from its z masters
 * (src/zgetrf_batched.cpp:11, src/zgetrs_batched.cpp:12, src/zgesv_batched.cpp:11, src/zgetrf_vbatched.cpp:11).
 * A fourth instantiation in double ("q") exists only so that tests can check the template against oracle_dgetf2 /
 * oracle_dgetrs of lu_oracle.c bit for bit.
 * Pinned by: tests/test_oracle.py (LAPACK through scipy, all three precisions; q-instantiation == lu_oracle.c).
 */
#include <math.h>
#include <stddef.h>

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

#define R float
#define FMA fmaf
#define FABS fabsf
#define CPLX 0
#define T s_t
#define PFX(name) CAT(oracle_s, name)
#include "lu_oracle_tmpl.h"
#undef R
#undef FMA
#undef FABS
#undef CPLX
#undef T
#undef PFX

#define R float
#define FMA fmaf
#define FABS fabsf
#define CPLX 1
#define T c_t
#define PFX(name) CAT(oracle_c, name)
#include "lu_oracle_tmpl.h"
#undef R
#undef FMA
#undef FABS
#undef CPLX
#undef T
#undef PFX

#define R double
#define FMA fma
#define FABS fabs
#define CPLX 1
#define T z_t
#define PFX(name) CAT(oracle_z, name)
#include "lu_oracle_tmpl.h"
#undef R
#undef FMA
#undef FABS
#undef CPLX
#undef T
#undef PFX

#define R double
#define FMA fma
#define FABS fabs
#define CPLX 0
#define T q_t
#define PFX(name) CAT(oracle_q, name)
#include "lu_oracle_tmpl.h"
