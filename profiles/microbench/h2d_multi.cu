// h2d_multi.cu — what can the HOST deliver to several GPUs at once?  One process per GPU (as bench.py runs), each copies a
// pinned 1 GiB buffer host->device back to back for a fixed wall-clock window that starts at a common absolute time.
//   h2d_multi <device> <start_unix_seconds> <duration_seconds>   ->  "<device> <GB/s>"
// tools/h2d_concurrent.sh runs 1, 2, 4 and 8 of them side by side and sums the rates.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <cuda_runtime.h>

int main(int argc, char** argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const double start = argc > 2 ? atof(argv[2]) : 0.0, dur = argc > 3 ? atof(argv[3]) : 2.0;
    cudaSetDevice(dev);
    const size_t bytes = (size_t)1 << 30;
    void *h, *d;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&d, bytes) != cudaSuccess) { printf("%d alloc failed\n", dev); return 1; }
    memset(h, 1, bytes);
    cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    auto now = [] { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); };
    while (now() < start) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    const double t0 = now();
    size_t moved = 0;
    while (now() - t0 < dur) {
        cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
        moved += bytes;
    }
    printf("%d %.2f\n", dev, moved / (now() - t0) / 1e9);
    return 0;
}
