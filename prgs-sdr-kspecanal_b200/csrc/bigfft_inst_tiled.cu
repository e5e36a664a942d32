// tiled column-pass instantiations (cols_tiled_kernel, see bigfft_kernels.cuh)
#include "bigfft_kernels.cuh"
namespace kspec {
template <int INFMT> static int cols_tiled_t(int l1, const OpColsIn<INFMT, false>& op, const cd* tw, int64_t nFrameSlabs, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l1, COLS_TILED_MIN_L, COLS_TILED_MAX_L, (launch_cols_tiled<LL, INFMT>(op, tw, nFrameSlabs, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
// op points to an OpColsIn<inFmt, false>
int big_cols_tiled(int inFmt, int l1, const void* op, const cd* tw, int64_t nFrameSlabs, int smCount, cudaStream_t st) {
    if (inFmt == KSPEC_IN_U8_IQ) return cols_tiled_t<KSPEC_IN_U8_IQ>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_U8_IQ, false>*>(op), tw, nFrameSlabs, smCount, st);
    if (inFmt == KSPEC_IN_C64) return cols_tiled_t<KSPEC_IN_C64>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_C64, false>*>(op), tw, nFrameSlabs, smCount, st);
    return cols_tiled_t<KSPEC_IN_C128>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_C128, false>*>(op), tw, nFrameSlabs, smCount, st);
}
}  // namespace kspec
