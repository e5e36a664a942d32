// column-pass instantiations of team_fft_kernel (see bigfft_kernels.cuh)
#include "bigfft_kernels.cuh"
namespace kspec {
template <int INFMT, bool BLUE> static int cols_in_t(int l1, const OpColsIn<INFMT, BLUE>& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l1, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpColsIn<INFMT, BLUE>>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
// op points to an OpColsIn<inFmt, blue != 0> (the two forms share one layout)
template <int INFMT> static int cols_in_f(int blue, int l1, const void* op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    if (blue) return cols_in_t<INFMT, true>(l1, *reinterpret_cast<const OpColsIn<INFMT, true>*>(op), tw, nBatch, smCount, st);
    return cols_in_t<INFMT, false>(l1, *reinterpret_cast<const OpColsIn<INFMT, false>*>(op), tw, nBatch, smCount, st);
}
int big_cols_in(int inFmt, int blue, int l1, const void* op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    if (inFmt == KSPEC_IN_U8_IQ) return cols_in_f<KSPEC_IN_U8_IQ>(blue, l1, op, tw, nBatch, smCount, st);
    if (inFmt == KSPEC_IN_C64) return cols_in_f<KSPEC_IN_C64>(blue, l1, op, tw, nBatch, smCount, st);
    return cols_in_f<KSPEC_IN_C128>(blue, l1, op, tw, nBatch, smCount, st);
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
