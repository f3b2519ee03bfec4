// kernels_inst.cu -- instantiates search_kernel for one strip height (compile with -DOPAL_R=<R>).
#include "search_kernel.cuh"

#ifndef OPAL_R
#error "compile with -DOPAL_R=<rows per thread>"
#endif
#define OPAL_CAT2(a, b) a##b
#define OPAL_CAT(a, b) OPAL_CAT2(a, b)

namespace opalb200 {
static const void* const kTable[9] = {
    (const void*)search_kernel<OPAL_R, kFlavorSWScore, Packed16>,
    (const void*)search_kernel<OPAL_R, kFlavorSWEnd, Packed16>,
    (const void*)search_kernel<OPAL_R, kFlavorGlobal, Packed16>,
    (const void*)search_kernel<OPAL_R, kFlavorSWEndFast, Packed16>,
    (const void*)search_kernel<OPAL_R, kFlavorSWScore, Scalar32>,
    (const void*)search_kernel<OPAL_R, kFlavorSWEnd, Scalar32>,
    (const void*)search_kernel<OPAL_R, kFlavorGlobal, Scalar32>,
    (const void*)search_kernel<OPAL_R, kFlavorSWEnd, Scalar32>,  // no fast variant at 32 bits
    (const void*)search_kernel<OPAL_R, kFlavorGlobal, Packed16, 384>,  // three warps per partition (170-register cap)
};
const void* const* OPAL_CAT(kernel_table_R, OPAL_R)() { return kTable; }

// chained passes: one warp per scheduler partition (128 threads, no register cap), [type * 4 + flavor]
#define OPAL_CHAINED(FLAVOR, TYPE) (const void*)search_kernel<OPAL_R, FLAVOR, TYPE, 128, true>
static const void* const kChainTable[8] = {
#if OPAL_R == 6 || OPAL_R == 9 || OPAL_R == 12 || OPAL_R == 17 || OPAL_R == 24 || OPAL_R == 33
    OPAL_CHAINED(kFlavorSWScore, Packed16), OPAL_CHAINED(kFlavorSWEnd, Packed16), OPAL_CHAINED(kFlavorGlobal, Packed16),
    OPAL_CHAINED(kFlavorSWEndFast, Packed16), OPAL_CHAINED(kFlavorSWScore, Scalar32), OPAL_CHAINED(kFlavorSWEnd, Scalar32),
    OPAL_CHAINED(kFlavorGlobal, Scalar32), OPAL_CHAINED(kFlavorSWEnd, Scalar32),
#else
    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
#endif
};
static_assert(chain_strip_height(OPAL_R) == (OPAL_R == 6 || OPAL_R == 9 || OPAL_R == 12 || OPAL_R == 17 || OPAL_R == 24 || OPAL_R == 33), "keep in step with search_kernel.cuh");
const void* const* OPAL_CAT(kernel_chain_table_R, OPAL_R)() { return kChainTable; }
}  // namespace opalb200
