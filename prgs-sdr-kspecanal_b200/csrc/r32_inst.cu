// fused scan kernel for fftSize 2048 in float32, 32 x 2 x 32 layout (see curscan_r32.cuh; curscan_r32p.cuh is its two-role pipeline): uint8 I/Q and complex64 ingest
#include "curscan_r32p.cuh"

namespace kspec {
// p.scanCounter != nullptr selects the dynamically scheduled form
int launch_r32_u8(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    return p.scanCounter ? launch_r32<KSPEC_IN_U8_IQ, true>(p, grid, st, info) : launch_r32<KSPEC_IN_U8_IQ, false>(p, grid, st, info);
}
int launch_r32_c64(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    return p.scanCounter ? launch_r32<KSPEC_IN_C64, true>(p, grid, st, info) : launch_r32<KSPEC_IN_C64, false>(p, grid, st, info);
}
int launch_r32p_u8(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) { return launch_r32p<KSPEC_IN_U8_IQ>(p, grid, st, info); }
int launch_r32p_c64(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) { return launch_r32p<KSPEC_IN_C64>(p, grid, st, info); }
}  // namespace kspec
