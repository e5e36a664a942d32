// fused small-M Bluestein kernel instantiations (see bigfft_kernels.cuh)
#include "bigfft_kernels.cuh"
namespace kspec {
template <int INFMT> static int blue_t(int logM, const BlueSmallParams& p, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(logM, 4, BLUE_SMALL_MAX_LOGM, (launch_bluestein_smem<INFMT, LL>(p, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
int big_blue_small(int inFmt, int logM, const BlueSmallParams& p, int smCount, cudaStream_t st) {
    if (inFmt == KSPEC_IN_U8_IQ) return blue_t<KSPEC_IN_U8_IQ>(logM, p, smCount, st);
    if (inFmt == KSPEC_IN_C64) return blue_t<KSPEC_IN_C64>(logM, p, smCount, st);
    return blue_t<KSPEC_IN_C128>(logM, p, smCount, st);
}
}  // namespace kspec
