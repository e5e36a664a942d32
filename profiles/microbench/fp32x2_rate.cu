// Micro-benchmark: issue/pipe rate of scalar FFMA/FADD vs packed FFMA2/FADD2 (f32x2) on sm_100a, and of a 50/50 mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_rate fp32x2_rate.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) rate(float2* out, float2 seed, int iters) {
    float2 a[8], b = seed, c = make_float2(seed.y, seed.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i, seed.y - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // scalar FFMA x2 (same flops as one FFMA2)
                a[i].x = fmaf(a[i].x, b.x, c.x);
                a[i].y = fmaf(a[i].y, b.y, c.y);
            } else if (MODE == 1) {     // packed FFMA2
                a[i] = __ffma2_rn(a[i], b, c);
            } else if (MODE == 2) {     // scalar FADD x2
                a[i].x = a[i].x + c.x;
                a[i].y = a[i].y + c.y;
            } else if (MODE == 3) {     // packed FADD2
                a[i] = __fadd2_rn(a[i], c);
            } else if (MODE == 4) {     // mix: half the values packed, half scalar (FFMA)
                if (i & 1) a[i] = __ffma2_rn(a[i], b, c);
                else { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            } else {                    // mix FADD
                if (i & 1) a[i] = __fadd2_rn(a[i], c);
                else { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
            }
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int warpsPerSm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000, blocks = sms * warpsPerSm / 4;
    float2* out; cudaMalloc(&out, sizeof(float2) * blocks * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate<MODE><<<blocks, 128>>>(out, make_float2(1.0001f, 0.9999f), 100);
    cudaEventRecord(e0);
    rate<MODE><<<blocks, 128>>>(out, make_float2(1.0001f, 0.9999f), iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // complex-pair ops per SM per cycle: each (i, it) step updates one float2 = 2 lane-flop-ops x 32 lanes per warp
    double pairOps = (double)iters * 8 * 128 * blocks;               // float2 updates
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s warps/SM %2d  %.3f ms  float2-updates/SM/clk %.2f  (lane-ops/SM/clk %.1f)\n", name, warpsPerSm, ms,
           pairOps / sms / cyc, 2 * pairOps / sms / cyc);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 12, 16, 32}) {
        run<0>("scalar FFMA x2", w); run<1>("packed FFMA2", w); run<4>("mix FFMA/FFMA2", w);
        run<2>("scalar FADD x2", w); run<3>("packed FADD2", w); run<5>("mix FADD/FADD2", w);
    }
    return 0;
}
