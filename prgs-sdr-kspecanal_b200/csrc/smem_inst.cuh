// smem_inst.cuh — instantiation helper: one translation unit per (precision, ingest format) so the
// heavily unrolled kernels compile in parallel.  The including .cu defines KSPEC_INST_T, KSPEC_INST_FMT,
// KSPEC_INST_NAME and KSPEC_INST_MAXLOG2F.
#include "curscan_smem.cuh"
#include <stdlib.h>

namespace kspec {

int KSPEC_INST_NAME(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    constexpr bool F32 = sizeof(KSPEC_INST_T) == 4;
    if constexpr (F32) {
        if (log2F == 11) {
            // the headline shape (fftSize 2048, float32): tuned layouts, see profiles/README.md
            if (variant == SMEM_VARIANT_FRAMES) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 3, true>(p, grid, st, info);
            int var = variant == SMEM_VARIANT_MULTI ? 4 : 3;
#ifdef KSPEC_INST_VARIANTS
            static const int forced = [] { const char* e = getenv("KSPEC_VARIANT"); return e ? atoi(e) : -1; }();
            if (forced >= 0) { if (variant == SMEM_VARIANT_MULTI) return (int)cudaErrorInvalidValue; var = forced; }
            if (var == 0) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 0>(p, grid, st, info);
            if (var == 1) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 1>(p, grid, st, info);
            if (var == 2) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 2>(p, grid, st, info);
            if (var == 5) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 5>(p, grid, st, info);
            if (var == 7) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 7>(p, grid, st, info);
            if (var == 6) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 6>(p, grid, st, info);
#endif
            if (var == 4) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 4>(p, grid, st, info);
            return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 3>(p, grid, st, info);
        }
    }
    if (variant != SMEM_VARIANT_BASE && variant != SMEM_VARIANT_FRAMES) return (int)cudaErrorInvalidValue;
    const bool vb = variant == SMEM_VARIANT_FRAMES;
    if constexpr (!F32) {
        // fftSize 2048 in float64 (the default precision): one exchange buffer and one TMA stage leave room for three CTAs per
        // SM instead of two: 81 -> 96 G samples/s (profiles/README.md)
        if (log2F == 11) {
#ifdef KSPEC_INST_VARIANTS
            static const int forced = [] { const char* e = getenv("KSPEC_VARIANT64"); return e ? atoi(e) : -1; }();
            if (forced == 0) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 0>(p, grid, st, info);
            if (forced == 9) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 9>(p, grid, st, info);
#endif
            if (vb) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 8, true>(p, grid, st, info);
            return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, 11, 8>(p, grid, st, info);
        }
    }
    switch (log2F) {
#define KSPEC_CASE(L) case L: if constexpr (L <= KSPEC_INST_MAXLOG2F && L != 11) { \
        if (vb) return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, L, 0, true>(p, grid, st, info); \
        return launch_smem_one<KSPEC_INST_T, KSPEC_INST_FMT, L>(p, grid, st, info); } else break;
        KSPEC_CASE(4) KSPEC_CASE(5) KSPEC_CASE(6) KSPEC_CASE(7) KSPEC_CASE(8) KSPEC_CASE(9) KSPEC_CASE(10)
        KSPEC_CASE(11) KSPEC_CASE(12) KSPEC_CASE(13) KSPEC_CASE(14)
#undef KSPEC_CASE
        default: break;
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace kspec
