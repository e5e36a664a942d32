// smem_inst.cuh — instantiation helper: one translation unit per (precision, ingest format) so the
// heavily unrolled kernels compile in parallel.  The including .cu defines KSPEC_INST_T, KSPEC_INST_FMT,
// KSPEC_INST_NAME and KSPEC_INST_MAXLOG2F.
#include "curscan_smem.cuh"
#include <stdlib.h>

namespace kspec {

int KSPEC_INST_NAME(int log2F, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
#ifdef KSPEC_INST_VARIANTS
    // tuning experiments on the headline shape only (fftSize 2048): KSPEC_VARIANT=1..3 in the environment
    if (log2F == 11) {
        static const int var = [] { const char* e = getenv("KSPEC_VARIANT"); return e ? atoi(e) : 0; }();
        if (var == 1) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 1>(p, grid, st, info);
        if (var == 2) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 2>(p, grid, st, info);
        if (var == 3) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 3>(p, grid, st, info);
    }
#endif
    switch (log2F) {
#define KSPEC_CASE(L) case L: if constexpr (L <= KSPEC_INST_MAXLOG2F) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, L>(p, grid, st, info); else break;
        KSPEC_CASE(4) KSPEC_CASE(5) KSPEC_CASE(6) KSPEC_CASE(7) KSPEC_CASE(8) KSPEC_CASE(9) KSPEC_CASE(10)
        KSPEC_CASE(11) KSPEC_CASE(12) KSPEC_CASE(13) KSPEC_CASE(14)
#undef KSPEC_CASE
        default: break;
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace kspec
