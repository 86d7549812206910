// bb_nv.cu -- one translation unit per number of variables: compiled 8 times with -DBB_NV=1..8 (see Makefile) so the
// monomial layout is a compile-time constant in every kernel and the instantiations build in parallel.
#ifndef BB_NV
#error "compile with -DBB_NV=<number of variables>"
#endif
#include "bb_kernels.cuh"

#define BB_CAT2(a, b) a##b
#define BB_CAT(a, b) BB_CAT2(a, b)
const BBKernelTable* BB_CAT(bb_kernel_table_nv, BB_NV)() { return BBLaunch<BB_NV>::table(); }

