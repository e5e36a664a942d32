// bigfft.cu — placeholder until the four-step / Bluestein engines land (next milestone).
#include "kspec_internal.h"
#include <stdio.h>
namespace kspec {
struct BigFft { int dummy; };
BigFft* bigfft_create(int, int, int64_t F, int, int64_t*, const double*, double, double, cudaStream_t, char* err, size_t errLen) {
    snprintf(err, errLen, "fftSize %lld needs the multi-pass engine, which is not built yet", (long long)F);
    return nullptr;
}
void bigfft_destroy(BigFft*) {}
int bigfft_run(BigFft*, const void*, int64_t, int64_t, const int64_t*, int, int, void*, int64_t*) { return KSPEC_ERR_UNSUPPORTED; }
}  // namespace kspec
