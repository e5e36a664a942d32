// h2d_rate.cu — host-to-device copy rate of the box for the end-to-end leg of bench.py (2 GiB of complex64 per step).
// nvcc -O2 -o h2d_rate h2d_rate.cu ;  ./h2d_rate
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

static double run(void* d, const void* h, size_t bytes, size_t chunk, int nStreams, cudaStream_t* st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, st[0]);
    if (nStreams == 1 && chunk >= bytes) {
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st[0]);
    } else {
        int i = 0;
        for (size_t o = 0; o < bytes; o += chunk, ++i) {
            const size_t n = bytes - o < chunk ? bytes - o : chunk;
            cudaMemcpyAsync((char*)d + o, (const char*)h + o, n, cudaMemcpyHostToDevice, st[i % nStreams]);
        }
        for (int s = 1; s < nStreams; ++s) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st[s]); cudaStreamWaitEvent(st[0], e, 0); }
    }
    cudaEventRecord(e1, st[0]);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return bytes / (ms * 1e-3) / 1e9;
}

int main() {
    const size_t bytes = (size_t)2 << 30;
    void* d;
    cudaMalloc(&d, bytes);
    cudaStream_t st[4];
    for (int i = 0; i < 4; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    const char* names[3] = {"cudaHostAlloc default", "cudaHostAlloc write-combined", "malloc + cudaHostRegister"};
    for (int kind = 0; kind < 3; ++kind) {
        void* h = nullptr;
        if (kind == 0) cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
        else if (kind == 1) cudaHostAlloc(&h, bytes, cudaHostAllocWriteCombined);
        else { h = aligned_alloc(4096, bytes); memset(h, 1, bytes); cudaHostRegister(h, bytes, cudaHostRegisterDefault); }
        memset(h, 1, bytes);
        run(d, h, bytes, bytes, 1, st);
        printf("%-32s one copy      %.1f GB/s\n", names[kind], run(d, h, bytes, bytes, 1, st));
        printf("%-32s 64 MiB x1 str %.1f GB/s\n", names[kind], run(d, h, bytes, (size_t)64 << 20, 1, st));
        printf("%-32s 64 MiB x2 str %.1f GB/s\n", names[kind], run(d, h, bytes, (size_t)64 << 20, 2, st));
        printf("%-32s 16 MiB x4 str %.1f GB/s\n", names[kind], run(d, h, bytes, (size_t)16 << 20, 4, st));
        if (kind == 2) { cudaHostUnregister(h); free(h); } else cudaFreeHost(h);
    }
    // device -> host for completeness
    void* h;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st[0]);
    cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st[0]);
    cudaEventRecord(e1, st[0]);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("device -> host, one copy                       %.1f GB/s\n", bytes / (ms * 1e-3) / 1e9);
    return 0;
}
