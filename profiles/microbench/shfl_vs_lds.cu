// shfl_vs_lds.cu — do warp shuffles share the shared-memory data pipe?  Three kernels with the same loop count: only LDS.64
// (conflict-free), only SHFL.BFLY, both interleaved.  If the mixed kernel takes max(a, b) the pipes are separate; if a + b, shared.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o shfl_vs_lds shfl_vs_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 16;

template <int MODE> __global__ void __launch_bounds__(512, 1) k(float* out) {
    __shared__ float2 buf[2048];
    const int t = threadIdx.x;
    buf[t] = make_float2(t, -t);
    buf[t + 512] = make_float2(t, t);
    buf[t + 1024] = make_float2(-t, t);
    buf[t + 1536] = make_float2(1, t);
    __syncthreads();
    float2 a = make_float2(0, 0);
    float s0 = t, s1 = 2 * t, s2 = 3 * t, s3 = 5 * t;
    int idx = t;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE & 1) {
                const float2 v = buf[(idx + u * 32) & 2047];
                a.x += v.x;
                a.y += v.y;
            }
            if (MODE & 2) {
                s0 = __shfl_xor_sync(0xffffffffu, s0, 1 + (u & 15));
                s1 = __shfl_xor_sync(0xffffffffu, s1, 1 + (u & 15));
            }
        }
        idx = (idx + (int)a.x) & 2047;
        s2 += s0;
        s3 += s1;
    }
    out[blockIdx.x * blockDim.x + t] = a.x + a.y + s2 + s3;
}

template <int MODE> float run(float* d, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<grid, 512>>>(d);
    cudaEventRecord(e0);
    k<MODE><<<grid, 512>>>(d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float* d;
    cudaMalloc(&d, (size_t)sms * 512 * 4);
    const float a = run<1>(d, sms), b = run<2>(d, sms), c = run<3>(d, sms);
    const double warpInstr = (double)ITERS * UNROLL * 16;              // per SM, per kind (LDS.64: 1 per iteration; SHFL: 2)
    printf("SMs %d, max clock %.0f MHz\n", sms, khz / 1e3);
    printf("LDS.64 only   %.3f ms  -> %.2f cycles per warp LDS.64 per SM (ideal 2: 256 B / 128 B per clock)\n", a, a * 1e-3 * khz * 1e3 / warpInstr);
    printf("SHFL only     %.3f ms  -> %.2f cycles per warp SHFL per SM\n", b, b * 1e-3 * khz * 1e3 / (2 * warpInstr));
    printf("both          %.3f ms  (sum %.3f, max %.3f)\n", c, a + b, a > b ? a : b);
    return 0;
}
