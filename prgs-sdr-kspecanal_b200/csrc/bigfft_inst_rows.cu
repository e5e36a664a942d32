// row-pass and plain instantiations of team_fft_kernel (see bigfft_kernels.cuh)
#include "bigfft_kernels.cuh"
namespace kspec {
int big_rows_plain(int l2, const OpRowsPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l2, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpRowsPlain>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
int big_rows_mul(int l2, const OpRowsMul& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l2, BIG_MIN_L, BIG_MAX_L, (launch_team_fft<LL, OpRowsMul>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
int big_rows_acc(int l2, int occ3, const RowsAccParams& p, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    if (occ3) { KSPEC_SWITCH_L(l2, BIG_MIN_L, BIG_MAX_L, (launch_team_fft_acc<LL, true>(p, tw, nBatch, smCount, st))) }
    else { KSPEC_SWITCH_L(l2, BIG_MIN_L, BIG_MAX_L, (launch_team_fft_acc<LL, false>(p, tw, nBatch, smCount, st))) }
    return (int)cudaErrorInvalidValue;
}
int big_plain(int l, const OpPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    KSPEC_SWITCH_L(l, 4, BLUE_SMALL_MAX_LOGM, (launch_team_fft<LL, OpPlain>(op, tw, nBatch, smCount, st)))
    return (int)cudaErrorInvalidValue;
}
}  // namespace kspec
