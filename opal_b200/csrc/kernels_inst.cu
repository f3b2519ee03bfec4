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
}  // namespace opalb200
