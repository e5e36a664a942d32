// db_math.cuh — 10*log10 in the working precision, shared by the scan kernels and the small per-bin kernels so that a value
// converted in either place is the same bit pattern (data_proc 'LogNoGain', K:106-112).
#pragma once
#include <cuda_runtime.h>

namespace kspec {

// float32 fast mode: MUFU.LG2 (abs error 2^-22 in log2 near 1, 2 ulp elsewhere) -> < 2e-5 dB, far inside the 1e-3 dB budget,
// for ~20 instructions less per bin than log10f.  0 -> -inf as numpy (K:109).
__device__ __forceinline__ float to_db(float v) { return 3.0102999566398120f * __log2f(v); }
__device__ __forceinline__ double to_db(double v) { return 10.0 * log10(v); }

}  // namespace kspec
