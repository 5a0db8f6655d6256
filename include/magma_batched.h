/* magma_batched.h -- the reference splits its batched prototypes into include/magma_batched.h ->
 * magma_{z,d}batched.h / magma_{z,d}vbatched.h; here they all live in magma_b200.h. */
#ifndef MAGMA_BATCHED_H
#define MAGMA_BATCHED_H
#include "magma_b200.h"
#endif
