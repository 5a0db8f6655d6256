/* magma_v2.h -- drop-in name for callers of the reference (include/magma_v2.h): the batched FP64 LU subset of the
 * magma_v2 API implemented by libmagma_b200.so. Everything is declared in magma_b200.h; this header only makes
 * `#include "magma_v2.h"` keep working when a caller switches libraries. */
#ifndef MAGMA_V2_H
#define MAGMA_V2_H
#include "magma_b200.h"
#endif
