// column-pass instantiations of team_fft_kernel (see bigfft_kernels.cuh)
#include "bigfft_kernels.cuh"
namespace kspec {
template <int INFMT> static int cols_in_t(int l1, const OpColsIn<INFMT>& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l1, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpColsIn<INFMT>>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
int big_cols_in(int inFmt, int l1, const void* op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    if (inFmt == KSPEC_IN_U8_IQ) return cols_in_t<KSPEC_IN_U8_IQ>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_U8_IQ>*>(op), tw, nBatch, smCount, st);
    if (inFmt == KSPEC_IN_C64) return cols_in_t<KSPEC_IN_C64>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_C64>*>(op), tw, nBatch, smCount, st);
    return cols_in_t<KSPEC_IN_C128>(l1, *reinterpret_cast<const OpColsIn<KSPEC_IN_C128>*>(op), tw, nBatch, smCount, st);
}
int big_cols_plain(int l1, const OpColsPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l1, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpColsPlain>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
int big_cols_mid(int l1, const OpColsMid& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l1, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpColsMid>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
}  // namespace kspec
