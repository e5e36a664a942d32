// curscan_r32p.cuh — the 32 x 2 x 32 layout of curscan_r32.cuh as a two-role pipeline (warp specialisation).
//
// The R32 kernel is bound by latency, not by a pipe: a 64-thread team runs load -> DFT32 -> twiddle -> shuffle -> exchange
// -> DFT32 -> |X| as one serial program per warp, needs 64 (data) + 32 (window) + 32 (accumulators) registers per thread, so
// only 12 warps fit on an SM, and the two warps of a team wait for each other twice per frame.  Here a team is FOUR warps
// in two roles that meet only through shared memory:
//
//   role A (2 warps)  stage 0: staged samples -> window -> DFT32 -> boundary twiddle -> shfl.xor radix-2 -> exchange write
//                     registers: data + window.            One frame ahead of role B.
//   role B (2 warps)  stage 1: exchange read -> DFT32 -> |X| -> cumulate; per-scan epilogue (dB, Max/Min, waterfall row)
//                     registers: data + accumulators.
//
// The exchange buffer is double buffered and handed over with mbarriers (full / empty, 64 arrivals each), so neither role
// waits for the other in steady state; 128 registers per thread are enough for either role, which gives 16 warps per SM
// (four teams, eight frames in flight) instead of 12.  Roles are assigned by warp group (warps 0-7: A, 8-15: B), so every
// SM sub-partition hosts two A warps and two B warps.  Same arithmetic as curscan_r32_kernel, bit for bit.
#pragma once
#include "curscan_r32.cuh"

namespace kspec {

struct R32PCfg {
    static constexpr int TEAMS = 4, NT = 64, CTA = TEAMS * 2 * NT;         // 512 threads
};

template <int INFMT> struct R32PStage {
    using S = R32Stage<INFMT>;
    static constexpr int TEAM_BYTES = 2 * R32Cfg::EX_BYTES + S::STAGE_BYTES;
    static constexpr int TW_OFS = R32PCfg::TEAMS * TEAM_BYTES;
    static constexpr int SMEM_BYTES = TW_OFS + R32Cfg::TW_BYTES;
    static constexpr bool OK = SMEM_BYTES <= 227 * 1024;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int INFMT>
__global__ void __launch_bounds__(R32PCfg::CTA, 1) curscan_r32p_kernel(const ScanParams p) {
    using C = R32Cfg;
    using SC = R32Stage<INFMT>;
    using PC = R32PStage<INFMT>;
    using IN = R32Raw<INFMT>;
    constexpr int P = C::P, F = C::F, NT = C::NT, TEAMS = R32PCfg::TEAMS, PITCH = C::PITCH;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar_all[TEAMS][5];                        // stage full, exchange full[2], exchange empty[2]
    __shared__ int32_t foffs[C::MAX_FRAMES];                       // K:386 frame starts inside a scan
    const int warp = threadIdx.x >> 5;
    const bool roleA = warp < 2 * TEAMS;
    const int team = (roleA ? warp : warp - 2 * TEAMS) >> 1;
    const int tid = ((warp & 1) << 5) | (threadIdx.x & 31);        // 0..63 inside the role team
    unsigned char* tbase = smem_raw + team * PC::TEAM_BYTES;
    float2* ex0 = reinterpret_cast<float2*>(tbase);
    unsigned char* stage = tbase + 2 * C::EX_BYTES;
    float2* stw = reinterpret_cast<float2*>(smem_raw + PC::TW_OFS);
    uint64_t* mbStage = &mbar_all[team][0];
    uint64_t* mbFull = &mbar_all[team][1];
    uint64_t* mbEmpty = &mbar_all[team][3];
    // barrier among the 64 threads of this role team
    const int barId = 1 + team + (roleA ? 0 : TEAMS);
    auto sync = [barId] { asm volatile("bar.sync %0, %1;" ::"r"(barId), "n"(NT) : "memory"); };

    const float2* __restrict__ gtw = reinterpret_cast<const float2*>(p.tw);      // exp(-2 pi i k / 2048)
    r32_build_twiddles(stw, gtw, threadIdx.x, R32PCfg::CTA);
    for (int i = threadIdx.x; i < p.nFrames; i += R32PCfg::CTA) foffs[i] = p.frameOffs[i];
    if (threadIdx.x < TEAMS) {
        mbar_init(&mbar_all[threadIdx.x][0], 1);
        for (int i = 1; i < 5; ++i) mbar_init(&mbar_all[threadIdx.x][i], NT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int slot = blockIdx.x * TEAMS + team;
    const int64_t scansPerIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (p.nScans + scansPerIter - 1) / scansPerIter;
    const int nFrames = p.nFrames;

    if (roleA) {
        // ================================ role A: stage 0 =====================================================================
        const int lane = tid & 31;
        const int upper = lane >> 4;                               // partner = lane ^ 16
        const int jp = (lane & 15) + 16 * (tid >> 5);              // j mod 32
        const int j = jp + 32 * upper;                             // owns x[j + 64 m]
        const bool leader = tid == 0;
        const float* __restrict__ gwin = reinterpret_cast<const float*>(p.win);
        const float u8off = (float)p.u8Offset;
        const float wscale = INFMT == KSPEC_IN_U8_IQ ? (float)p.u8Scale : 1.0f;      // the uint8 scale rides in the window
        float win[P];
#pragma unroll
        for (int m = 0; m < P; ++m) {
            const float w = gwin[j + NT * m];
            win[m] = ((upper && (m & 1)) ? -w : w) * wscale;                  // rotates the upper thread's DFT32 outputs by 16 slots
        }
        float2 omega = gtw[32 * jp];                               // W_64^(j mod 32); the upper thread computes b - a
        if (upper) omega = make_float2(-omega.x, -omega.y);
        const uint64_t polStream = l2_policy_evict_first();
        const float4* tw4 = reinterpret_cast<const float4*>(stw) + tid;
        float2* exw = ex0 + (16 * upper) * PITCH + jp;

        const int64_t totalElems = p.nScans * p.scanStride;
        auto issue = [&](int64_t sc, int f) {                      // leader only: fetch frame f of scan sc (see curscan_r32.cuh)
            constexpr int64_t GM = SC::SLACK > 0 ? SC::SLACK - 1 : 0;
            const int64_t e0 = sc * p.scanStride + foffs[f];
            const int64_t e0a = e0 & ~GM;
            int64_t e1a = (e0 + F + GM) & ~GM;
            const int64_t total = (totalElems + GM) & ~GM;
            if (e1a > total) e1a = total;
            const uint32_t bytes = (uint32_t)((e1a - e0a) * SC::EB);
            mbar_expect_tx(mbStage, bytes);
            tma_load_1d_hint(stage, reinterpret_cast<const unsigned char*>(p.samples) + e0a * SC::EB, bytes, mbStage, polStream);
        };
        if (leader && iters > 0) issue(slot < p.nScans ? slot : p.nScans - 1, 0);

        uint32_t g = 0;                                            // frames done by this team (parities only)
        for (int64_t it = 0; it < iters; ++it) {
            int64_t scanC = it * scansPerIter + slot;
            if (scanC >= p.nScans) scanC = p.nScans - 1;           // idle teams shadow the last scan
            const int64_t sbase = scanC * p.scanStride;
            for (int f = 0; f < nFrames; ++f, ++g) {
                float2 b[P];
                mbar_wait(mbStage, g & 1);
                {
                    const int mis = SC::SLACK > 0 ? (((int)sbase + foffs[f]) & (SC::SLACK - 1)) : 0;
                    const typename IN::raw_t* sp = reinterpret_cast<const typename IN::raw_t*>(stage) + mis + j;
#pragma unroll
                    for (int m = 0; m < P; ++m) b[m] = IN::get(sp[NT * m], u8off);
                }
                sync();                                            // both A warps have consumed the stage buffer
                if (leader) {
                    const bool lastF = f + 1 == nFrames;
                    if (!lastF || it + 1 < iters) {
                        int64_t sc = lastF ? scanC + scansPerIter : scanC;
                        if (sc >= p.nScans) sc = p.nScans - 1;
                        issue(sc, lastF ? 0 : f + 1);
                    }
                }
                const uint32_t buf = g & 1;
                mbar_wait(&mbEmpty[buf], ((g >> 1) & 1) ^ 1);      // role B has released this buffer (first two uses: free)
                r32_stage0(b, win, tw4, omega, exw + buf * (C::EX_BYTES / 8));
                mbar_arrive(&mbFull[buf]);
            }
        }
    } else {
        // ================================ role B: stage 1 + cumulate + per-scan outputs ========================================
        const bool avgScaled = p.cumuMode == KSPEC_CUMU_AVG && nFrames <= 96;     // see curscan_smem.cuh
        const float linScale = avgScaled ? (float)ldexp(p.linScale, -(nFrames - 1)) : (float)p.linScale;
        const uint64_t polKeep = l2_policy_evict_last();
        uint32_t g = 0;
        for (int64_t it = 0; it < iters; ++it) {
            const int64_t scan = it * scansPerIter + slot;
            const bool valid = scan < p.nScans;
            float acc[P];
            float avgW = 1.0f;
            uint32_t buf = 0;
            for (int f = 0; f < nFrames; ++f, ++g) {
                float2 b[P];
                buf = g & 1;
                mbar_wait(&mbFull[buf], (g >> 1) & 1);
                {
                    const float4* q4 = reinterpret_cast<const float4*>(ex0 + buf * (C::EX_BYTES / 8) + tid * PITCH);
#pragma unroll
                    for (int i = 0; i < P; i += 2) {
                    const float4 v = q4[i >> 1];
                    b[i] = make_float2(v.x, v.y);
                    b[i + 1] = make_float2(v.z, v.w);
                }
                }
                if (f + 1 < nFrames) mbar_arrive(&mbEmpty[buf]);   // the last frame's buffer doubles as the epilogue's scratch row
                dft32(b);
                float mag[P];
#pragma unroll
                for (int m = 0; m < P; ++m) mag[m] = kabs_fma(b[m]);
                if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) {
#pragma unroll
                    for (int m = 0; m < P; ++m) acc[m] = mag[m];
                } else if (avgScaled) {
#pragma unroll
                    for (int m = 0; m < P; ++m) acc[m] = fmaf(mag[m], avgW, acc[m]);
                    avgW += avgW;
                } else if (p.cumuMode == KSPEC_CUMU_AVG) {
#pragma unroll
                    for (int m = 0; m < P; ++m) acc[m] = (acc[m] + mag[m]) * 0.5f;
                } else if (p.cumuMode == KSPEC_CUMU_MAX) {
#pragma unroll
                    for (int m = 0; m < P; ++m) acc[m] = fmaxf(acc[m], mag[m]);
                } else {
#pragma unroll
                    for (int m = 0; m < P; ++m) acc[m] = fminf(acc[m], mag[m]);
                }
            }
            // per-scan epilogue (see curscan_r32.cuh); role A meanwhile fills the other buffer with the next scan's first frame
            float* erow = reinterpret_cast<float*>(ex0 + buf * (C::EX_BYTES / 8));
            sync();                                                // both B warps have read the last frame's exchange data
#pragma unroll
            for (int m = 0; m < P; ++m) erow[(tid + NT * m) ^ (F >> 1)] = acc[m] * linScale;
            sync();
            const bool hmDone = scan_epilogue_rows<NT>(p, erow, scan, valid, it, slot, tid, polKeep);
            if (p.hm != nullptr && !hmDone) {
                sync();
                const int W = p.hmW, gsz = F / W;
                float* __restrict__ hm = reinterpret_cast<float*>(p.hm);
                for (int w = tid; w < W; w += NT) {
                    float r = erow[w * gsz];
                    if (p.hmMode == KSPEC_COMPRESS_MAX) {
                        for (int q = 1; q < gsz; ++q) r = fmaxf(r, erow[w * gsz + q]);
                    } else if (p.hmMode == KSPEC_COMPRESS_MIN) {
                        for (int q = 1; q < gsz; ++q) r = fminf(r, erow[w * gsz + q]);
                    } else if (p.hmMode == KSPEC_COMPRESS_AVG) {
                        for (int q = 1; q < gsz; ++q) r += erow[w * gsz + q];
                        r /= (float)gsz;
                    }
                    if (valid) hm[scan * W + w] = r;
                }
            }
            mbar_arrive(&mbEmpty[buf]);                            // each thread releases after its own last read of the scratch row
        }
    }
}

template <int INFMT>
static int launch_r32p(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    using PC = R32PStage<INFMT>;
    if constexpr (!PC::OK) {
        return (int)cudaErrorInvalidValue;
    } else {
        auto k = curscan_r32p_kernel<INFMT>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        if (info) {
            info->ctaThreads = R32PCfg::CTA;
            info->smemBytes = PC::SMEM_BYTES;
            info->teams = R32PCfg::TEAMS;
            info->stages = 1;
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, R32PCfg::CTA, PC::SMEM_BYTES);
            info->ctasPerSm = nb;
            return 0;
        }
        k<<<grid, R32PCfg::CTA, PC::SMEM_BYTES, st>>>(p);
        return (int)cudaGetLastError();
    }
}

}  // namespace kspec
