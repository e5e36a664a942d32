// tc_dft16.cu — microbenchmark: one radix-16 DFT stage of a 2048-point frame (128 butterflies) as a tensor-core GEMM.
//
//   D[128 x 32] = A[128 x 32] * B[32 x 32]      A: butterfly j, K = (re,im) of its 16 inputs;  B: DFT16 written out in re/im
//   tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = 32, K = 8 per instruction (4 per term), accumulator in TMEM (32 columns),
//   tcgen05.ld.32x32b.x32 back to registers: thread t of warp w receives row 32 w + t = the 16 complex outputs of ITS butterfly,
//   exactly the register state after the SIMT stage.  1, 2 or 3 operand-split terms (x_hi F_hi, + x_lo F_hi, + x_hi F_lo).
//
// Reports (a) max relative error of D against a float64 DFT16 for each number of terms (validates the descriptors and
// layouts), (b) SM cycles per 128-butterfly tile for the tensor path (operands resident in shared memory; MMA issue + commit +
// TMEM read-back, double-buffered over two TMEM accumulators) and for the SIMT radix-16 butterfly (packed FP32, registers).
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../prgs-sdr-kspecanal_b200/csrc tc_dft16.cu -o tc_dft16
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "fft_core.cuh"

using namespace kspec;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, 128-byte swizzled operand tile: row r (128 B = 32 tf32) of an 8-row / 1024-byte atom, 16-byte chunks XOR-ed with r % 8
__host__ __device__ inline int sw128_offset(int r, int k) {
    const int chunk = (k >> 2) ^ (r & 7);
    return (r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (k & 3) * 4;
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    // start address >> 4 | LBO (unused for swizzled K-major: 1) | SBO = 1024 B between 8-row groups | version 1 (sm_100) | SWIZZLE_128B
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32, D = F32, A/B = TF32 K-major, N = 32, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(n)); }
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
          "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
          "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
          "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Ops {   // host-prepared operand tiles (already swizzled): A_hi, A_lo [128 x 32], B_hi, B_lo [32 x 32] (K-major, B transposed)
    float aHi[128 * 32], aLo[128 * 32], bHi[32 * 32], bLo[32 * 32];
};

// terms: 1..3.  iters tiles per CTA.  out: D of the LAST tile (for the check).  cyc: cycles of the timed loop (thread 0).
__global__ void __launch_bounds__(128, 4) tc_kernel(const Ops* __restrict__ ops, int terms, int iters, float* __restrict__ out, long long* __restrict__ cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* aHi = reinterpret_cast<float*>(smem);
    float* aLo = aHi + 128 * 32;
    float* bHi = aLo + 128 * 32;
    float* bLo = bHi + 32 * 32;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) { aHi[i] = ops->aHi[i]; aLo[i] = ops->aLo[i]; }
    for (int i = tid; i < 32 * 32; i += 128) { bHi[i] = ops->bHi[i]; bLo[i] = ops->bLo[i]; }
    if (tid == 0) { bar_init(&bars[0], 1); bar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tmemBase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmemBase;
    const uint64_t dAhi = smem_desc(s32(aHi)), dAlo = smem_desc(s32(aLo)), dBhi = smem_desc(s32(bHi)), dBlo = smem_desc(s32(bLo));

    auto issue = [&](int buf) {      // one 128-butterfly tile into accumulator `buf`: `terms` x 4 MMAs of K = 8 (32 bytes of K each)
        const uint32_t d = tm + buf * 32;
        for (int k = 0; k < 4; ++k) mma_tf32(d, dAhi + (uint64_t)(k * 2), dBhi + (uint64_t)(k * 2), k > 0);
        if (terms >= 2) for (int k = 0; k < 4; ++k) mma_tf32(d, dAlo + (uint64_t)(k * 2), dBhi + (uint64_t)(k * 2), 1);
        if (terms >= 3) for (int k = 0; k < 4; ++k) mma_tf32(d, dAhi + (uint64_t)(k * 2), dBlo + (uint64_t)(k * 2), 1);
        mma_commit(&bars[buf]);
    };
    float keep = 0.f;
    float v[32];
    long long t0 = clock64();
    if (tid == 0) issue(0);
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (tid == 0 && it + 1 < iters) issue(buf ^ 1);                // next tile's MMAs run while this tile is read back
        bar_wait(&bars[buf], (it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + buf * 32, v);   // lanes 32 w .. 32 w + 31, columns of this accumulator
#pragma unroll
        for (int i = 0; i < 32; ++i) keep += v[i];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                               // every warp has drained `buf` before it is overwritten
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (blockIdx.x == 0) for (int i = 0; i < 32; ++i) out[tid * 32 + i] = v[i];
    if (keep == 123.456f) out[0] = keep;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}

// the SIMT stage: 128 threads, one radix-16 butterfly each per tile, packed FP32 in registers (what curscan_smem.cuh does)
__global__ void __launch_bounds__(128, 1) simt_kernel(const float2* __restrict__ in, int iters, float2* __restrict__ out, long long* __restrict__ cyc) {
    float2 x[16];
    for (int m = 0; m < 16; ++m) x[m] = in[threadIdx.x * 16 + m];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        dft16<float>(x);
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = make_float2(x[m].x * 0.0625f, x[m].y * 0.0625f);     // keep the values bounded
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    for (int m = 0; m < 16; ++m) out[(blockIdx.x * 128 + threadIdx.x) * 16 + m] = x[m];
}

static float tf32_round(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
    memcpy(&f, &u, 4);
    return f;
}

int main() {
    std::vector<double> xr(128 * 16), xi(128 * 16);
    srand(7);
    for (size_t i = 0; i < xr.size(); ++i) { xr[i] = rand() / (double)RAND_MAX - 0.5; xi[i] = rand() / (double)RAND_MAX - 0.5; }
    Ops* h = new Ops();
    memset(h, 0, sizeof(Ops));
    for (int j = 0; j < 128; ++j)
        for (int k = 0; k < 32; ++k) {
            const float v = (float)((k & 1) ? xi[j * 16 + k / 2] : xr[j * 16 + k / 2]);
            const float hi = tf32_round(v), lo = tf32_round(v - hi);
            h->aHi[sw128_offset(j, k) / 4] = hi;
            h->aLo[sw128_offset(j, k) / 4] = lo;
        }
    for (int n = 0; n < 32; ++n)          // B^T[n][k]: n = 2 k1 (+1 for the imaginary part), k = 2 m (+1)
        for (int k = 0; k < 32; ++k) {
            const double th = 2.0 * M_PI * (double)((k / 2) * (n / 2)) / 16.0;
            double b;
            if (!(n & 1)) b = (k & 1) ? sin(th) : cos(th);
            else b = (k & 1) ? cos(th) : -sin(th);
            const float v = (float)b, hi = tf32_round(v), lo = tf32_round(v - hi);
            h->bHi[sw128_offset(n, k) / 4] = hi;
            h->bLo[sw128_offset(n, k) / 4] = lo;
        }
    Ops* d;
    float* dOut;
    long long* dCyc;
    cudaMalloc(&d, sizeof(Ops));
    cudaMalloc(&dOut, 128 * 32 * 4);
    cudaMalloc(&dCyc, 8);
    cudaMemcpy(d, h, sizeof(Ops), cudaMemcpyHostToDevice);
    const int smemBytes = (2 * 128 * 32 + 2 * 32 * 32) * 4 + 1024;
    cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<float> out(128 * 32);
    const int iters = 4096;
    for (int terms = 1; terms <= 3; ++terms) {
        tc_kernel<<<sms, 128, smemBytes>>>(d, terms, 64, dOut, dCyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tc_kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, ref = 0;
        for (int j = 0; j < 128; ++j)
            for (int k1 = 0; k1 < 16; ++k1) {
                double re = 0, im = 0;
                for (int m = 0; m < 16; ++m) {
                    const double th = -2.0 * M_PI * m * k1 / 16.0;
                    re += xr[j * 16 + m] * cos(th) - xi[j * 16 + m] * sin(th);
                    im += xr[j * 16 + m] * sin(th) + xi[j * 16 + m] * cos(th);
                }
                err = fmax(err, fmax(fabs(out[j * 32 + 2 * k1] - re), fabs(out[j * 32 + 2 * k1 + 1] - im)));
                ref = fmax(ref, fmax(fabs(re), fabs(im)));
            }
        printf("tcgen05 kind::tf32 M128 N32 K32, %d term(s), %d MMAs per 128-butterfly tile: max |err| / max |D| = %.3e\n", terms, 4 * terms, err / ref);
        for (int c = 1; c <= 4; c *= 2) {      // c co-resident CTAs per SM, each with its own operands and TMEM accumulators
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            tc_kernel<<<sms * c, 128, smemBytes>>>(d, terms, iters, dOut, dCyc);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            long long cyc = 0;
            cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
            printf("    %d CTA(s) per SM: %.1f SM cycles per tile and SM (CTA 0: %.1f cycles per own tile), %.3f ms for %d tiles per CTA\n",
                   c, (double)cyc / iters / c, (double)cyc / iters, ms, iters);
        }
    }
    float2* dIn;
    float2* dO2;
    cudaMalloc(&dIn, 128 * 16 * 8);
    cudaMalloc(&dO2, (size_t)sms * 128 * 16 * 8);
    std::vector<float2> hin(128 * 16);
    for (size_t i = 0; i < hin.size(); ++i) hin[i] = make_float2((float)xr[i], (float)xi[i]);
    cudaMemcpy(dIn, hin.data(), hin.size() * 8, cudaMemcpyHostToDevice);
    for (int warps = 0; warps < 1; ++warps) {
        simt_kernel<<<sms, 128>>>(dIn, iters, dO2, dCyc);
        cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
        printf("SIMT radix-16 butterfly, packed FP32, 128 threads (4 warps) alone on the SM: %.1f SM cycles per 128-butterfly tile (latency-bound at 4 warps; "
               "the pipe needs 128 x 160 lane-ops / 128 lanes = 160 cycles)\n", (double)cyc / iters);
    }
    return 0;
}
