// curscan_smem.cuh — the fused hot-path kernel for power-of-two fftSize that fits in shared memory.
//
// One launch replaces, for a whole batch of scans:
//   sdr_curscan            K:385-397   overlapped frame gather, window multiply, FFT, |X|, normalise, cumulate, fftshift
//   data_cumu              K:124-147   RAW / AVG (halving) / MAX / MIN across the frames of a scan
//   data_proc              K:100-112   optional low clip, 10*log10(.) - gain, optional inf -> 0
//   zero_span stats        K:471-476   per-bin Max / Min partials, last rows for the Avg recurrence
//   _data_plotcompress     K:168-202   waterfall row = compress(dB - adj)
//
// Mapping: a team of NT = F/P threads owns one scan at a time (P = 16 complex points per thread for F >= 128);
// teams stride over the scans of the batch.  A frame never leaves the SM: IQ samples are converted and windowed
// while they are loaded into registers (each sample comes from HBM once; the overlapped re-reads of later
// frames hit L1/L2), the FFT runs in registers with shared-memory exchanges, and magnitude + cumulate stay in
// the registers that own the bins.  Only the per-scan outputs are written.
#pragma once
#include "fft_core.cuh"
#include "kspec_internal.h"

namespace kspec {

template <typename T, int LOG2F> struct SmemCfg {
    static constexpr int LOG2P = LOG2F >= 7 ? 4 : (LOG2F >= 5 ? 3 : 2);
    static constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
    static constexpr int CTA = NT < 128 ? 128 : NT;
    static constexpr int TEAMS = CTA / NT;
    static constexpr bool F32 = sizeof(T) == 4;
    static constexpr bool REGTAB = F32 && LOG2F <= 11;   // window + twiddles live in registers across frames
    static constexpr int FPAD = padded_len(F);
    static constexpr int BUF_BYTES = FPAD * (int)sizeof(cx<T>) * TEAMS;
    static constexpr bool DBUF = 2 * BUF_BYTES <= 160 * 1024;
    static constexpr int SMEM_BYTES = (DBUF ? 2 : 1) * BUF_BYTES;
    static constexpr int MINB = CTA == 128 ? (F32 ? 3 : 2) : (CTA == 256 ? (F32 ? 2 : 1) : 1);
    static constexpr int NTW = twiddle_count<LOG2F, LOG2P>();
    static constexpr int NX = exchange_count<LOG2F, LOG2P>();
};

// ---- fused IQ ingest: one sample -> cx<T>, already multiplied by the window value -----------------------------
template <typename T, int INFMT> struct Ingest;
template <typename T> struct Ingest<T, KSPEC_IN_U8_IQ> {
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T off, T scale) {
        const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(base) + i);
        return mkcx<T>((((T)v.x - off) * scale) * w, (((T)v.y - off) * scale) * w);
    }
};
template <typename T> struct Ingest<T, KSPEC_IN_C64> {
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T, T) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(base) + i);
        return mkcx<T>((T)v.x * w, (T)v.y * w);
    }
};
template <typename T> struct Ingest<T, KSPEC_IN_C128> {
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T, T) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(base) + i);
        return mkcx<T>((T)v.x * w, (T)v.y * w);
    }
};

__device__ __forceinline__ float kabs(float2 a) {
    const float s = a.x * a.x + a.y * a.y;
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}
__device__ __forceinline__ double kabs(double2 a) { return sqrt(a.x * a.x + a.y * a.y); }

__device__ __forceinline__ float to_db(float v) { return 10.0f * log10f(v); }
__device__ __forceinline__ double to_db(double v) { return 10.0 * log10(v); }

template <typename T> __device__ __forceinline__ T pos_inf();
template <> __device__ __forceinline__ float pos_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double pos_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T, int INFMT, int LOG2F>
__global__ void __launch_bounds__(SmemCfg<T, LOG2F>::CTA, SmemCfg<T, LOG2F>::MINB)
curscan_smem_kernel(const ScanParams p) {
    using C = SmemCfg<T, LOG2F>;
    constexpr int P = C::P, F = C::F, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P;
    constexpr int L0 = stage_l<LOG2F, LOG2P>(0);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    // two exchange buffers per team (bufA/bufB) when they fit, else one.  Exchange x of a frame uses bufA for even
    // x and bufB for odd x; with an odd exchange count the roles swap after every frame, so a buffer is never
    // rewritten before a full barrier separates it from its last readers.
    cx<T>* bufA = reinterpret_cast<cx<T>*>(smem_raw) + team * C::FPAD;
    cx<T>* bufB = C::DBUF ? bufA + TEAMS * C::FPAD : bufA;

    const T* __restrict__ gwin = reinterpret_cast<const T*>(p.win);
    const cx<T>* __restrict__ gtw = reinterpret_cast<const cx<T>*>(p.tw);
    auto sync = [] { __syncthreads(); };

    // tables that stay in registers for the life of the CTA (fast path)
    T win[C::REGTAB ? P : 1];
    cx<T> twl[(C::REGTAB && C::NTW > 0) ? C::NTW : 1];
    if constexpr (C::REGTAB) {
#pragma unroll
        for (int m = 0; m < P; ++m) win[m] = gwin[tid + NT * m];
        load_twiddles<T, LOG2F, LOG2P>(twl, gtw, tid);
    }

    const T u8off = (T)p.u8Offset, u8scale = (T)p.u8Scale;
    const T linScale = (T)p.linScale;
    const int slot = blockIdx.x * TEAMS + team;                 // stats partial owned by this team
    const int64_t scansPerIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (p.nScans + scansPerIter - 1) / scansPerIter;

    for (int64_t it = 0; it < iters; ++it) {
        const int64_t scan = it * scansPerIter + slot;
        const bool valid = scan < p.nScans;
        const int64_t scanC = valid ? scan : p.nScans - 1;      // idle teams shadow the last scan (uniform barriers)
        const int64_t sbase = scanC * p.scanStride;

        T acc[P];
        for (int f = 0; f < p.nFrames; ++f) {
            const int64_t fbase = sbase + p.frameOffs[f];
            cx<T> b[P];
#pragma unroll
            for (int m = 0; m < P; ++m) {
                const T w = C::REGTAB ? win[m] : __ldg(&gwin[tid + NT * m]);
                b[m] = Ingest<T, INFMT>::load(p.samples, fbase + tid + NT * m, w, u8off, u8scale);
            }
            butterflies<T, P, (1 << L0), false>(b, nullptr);
            fft_tail<T, LOG2F, LOG2P, C::REGTAB, C::DBUF, L0, 0, 0>(b, twl, gtw, bufA, bufB, tid, sync);
            if constexpr (C::DBUF && (C::NX & 1)) { cx<T>* t = bufA; bufA = bufB; bufB = t; }
            // |X| and cumulate over the frames of this scan (data_cumu, K:124-147); normalisation is applied once per scan
            if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = kabs(b[m]);
            } else if (p.cumuMode == KSPEC_CUMU_AVG) {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = (acc[m] + kabs(b[m])) * (T)0.5;
            } else if (p.cumuMode == KSPEC_CUMU_MAX) {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = fmax(acc[m], kabs(b[m]));
            } else {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = fmin(acc[m], kabs(b[m]));
            }
        }

        // ---------------- per-scan epilogue: thread owns bins k = tid + NT*m, shifted position j = k ^ F/2 -------------
        T* __restrict__ rows = reinterpret_cast<T*>(p.rows);
        const bool needDb = (p.rowsKind == KSPEC_ROWS_DB) || p.wantStats || (p.hm != nullptr);
        const bool needRow = (p.hm != nullptr);
        // scratch row for the waterfall compress: bufA, the buffer the last exchange did NOT use (see above)
        T* erow = reinterpret_cast<T*>(bufA);
        if (needRow && !C::DBUF) __syncthreads();
#pragma unroll
        for (int m = 0; m < P; ++m) {
            const int j = (tid + NT * m) ^ (F >> 1);
            T lin = acc[m] * linScale;
            if (valid && p.rowsKind == KSPEC_ROWS_LINEAR) rows[scan * F + j] = lin;
            if (needDb) {
                if (p.dbClip) lin = fmax(lin, (T)p.minAmp);
                T db = to_db(lin) - (T)p.gain;
                if (p.infToZero && isinf(db)) db = (T)0;
                if (valid && p.rowsKind == KSPEC_ROWS_DB) rows[scan * F + j] = db;
                if (p.wantStats) {
                    T* __restrict__ wmax = reinterpret_cast<T*>(p.wsMax) + (int64_t)slot * F;
                    T* __restrict__ wmin = reinterpret_cast<T*>(p.wsMin) + (int64_t)slot * F;
                    if (it == 0) {
                        wmax[j] = valid ? db : -pos_inf<T>();
                        wmin[j] = valid ? db : pos_inf<T>();
                    } else if (valid) {
                        wmax[j] = fmax(wmax[j], db);
                        wmin[j] = fmin(wmin[j], db);
                    }
                    const int64_t ar = scan - (p.nScans - p.avgWin);
                    if (valid && ar >= 0) reinterpret_cast<T*>(p.avgRows)[ar * F + j] = db;
                }
                if (needRow) erow[j] = p.adj ? db - reinterpret_cast<const T*>(p.adj)[j] : db;
            }
        }
        if (needRow) {
            __syncthreads();
            // _data_plotcompress (K:184-200): W groups of g adjacent bins
            const int W = p.hmW, g = F / W;
            T* __restrict__ hm = reinterpret_cast<T*>(p.hm);
            const T* row = erow;
            for (int w = tid; w < W; w += NT) {
                T r = row[w * g];
                if (p.hmMode == KSPEC_COMPRESS_MAX) {
                    for (int q = 1; q < g; ++q) r = fmax(r, row[w * g + q]);
                } else if (p.hmMode == KSPEC_COMPRESS_MIN) {
                    for (int q = 1; q < g; ++q) r = fmin(r, row[w * g + q]);
                } else if (p.hmMode == KSPEC_COMPRESS_AVG) {
                    for (int q = 1; q < g; ++q) r += row[w * g + q];
                    r /= (T)g;
                }
                if (valid) hm[scan * W + w] = r;
            }
            __syncthreads();
        }
    }
}

template <typename T, int INFMT, int LOG2F>
static int launch_smem_one(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    using C = SmemCfg<T, LOG2F>;
    auto k = curscan_smem_kernel<T, INFMT, LOG2F>;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev & 63] = true;
    }
    if (info) {
        info->ctaThreads = C::CTA;
        info->smemBytes = C::SMEM_BYTES;
        info->teams = C::TEAMS;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, C::CTA, C::SMEM_BYTES);
        info->ctasPerSm = nb;
        return 0;
    }
    k<<<grid, C::CTA, C::SMEM_BYTES, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace kspec
