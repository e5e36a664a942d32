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
#include "db_math.cuh"
#include "kspec_internal.h"

namespace kspec {

// VAR selects a tuning variant of the same kernel (A/B experiments, see profiles/README.md):
//   0 production   1 twiddles via L1 instead of registers, 4 CTAs/SM   2 as 1 without TMA staging   3 as 0 with one stage
//   5 as 0 with 2 CTAs/SM (255 registers)   6 as 3 with the twiddles in a linearised shared-memory table
//   7 as 4 without TMA staging (frames loaded straight from L1/L2 into registers)
//   8 / 9 float64 occupancy variants: one exchange buffer + one TMA stage, 3 / 4 CTAs per SM (170 / 128 registers)
//   4 one 512-thread CTA per SM = four independent 128-thread teams (named barriers) sharing one shared-memory twiddle table
template <typename T, int LOG2F, int VAR = 0> struct SmemCfg {
    static constexpr int LOG2P = LOG2F >= 7 ? 4 : (LOG2F >= 5 ? 3 : 2);
    static constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
    static constexpr bool MULTI = (VAR == 4 || VAR == 7) && NT >= 32;  // teams of whole warps that never wait for each other
    static constexpr int CTA = MULTI ? 4 * NT : (NT < 128 ? 128 : NT);
    static constexpr int TEAMS = CTA / NT;
    static constexpr bool F32 = sizeof(T) == 4;
    static constexpr bool REGTAB = F32 && LOG2F <= 11;   // window (and, by default, twiddles) live in registers across frames
    static constexpr bool TWREG = REGTAB && !(VAR == 1 || VAR == 2 || VAR == 6 || MULTI);
    static constexpr bool TWSMEM = MULTI || VAR == 6;    // twiddle table copied to shared memory once per CTA
    static constexpr int FPAD = padded_len(F);
    static constexpr int BUF_BYTES = FPAD * (int)sizeof(cx<T>) * TEAMS;
    static constexpr bool OCC = VAR == 8 || VAR == 9;
    static constexpr bool DBUF = !OCC && 2 * BUF_BYTES <= 160 * 1024;
    static constexpr int SMEM_BYTES = (DBUF ? 2 : 1) * BUF_BYTES;
    static constexpr int MINB = VAR == 8 ? 3 : VAR == 9 ? 4 : MULTI ? 1 : (VAR == 5) ? 2 : (VAR == 1 || VAR == 2) ? 4 : (CTA == 128 ? (F32 ? 3 : 2) : (CTA == 256 ? (F32 ? 2 : 1) : 1));
    static constexpr int NTW = twiddle_count<LOG2F, LOG2P>();
    static constexpr int NX = exchange_count<LOG2F, LOG2P>();
};

// ---- fused IQ ingest: one sample -> cx<T>, already multiplied by the window value -----------------------------
template <typename T, int INFMT> struct Ingest;
template <typename T> struct Ingest<T, KSPEC_IN_U8_IQ> {
    static constexpr int ELEM_BYTES = 2;
    typedef uchar2 raw_t;
    static __device__ __forceinline__ cx<T> conv(uchar2 v, T w, T off, T scale) {
        return mkcx<T>((((T)v.x - off) * scale) * w, (((T)v.y - off) * scale) * w);
    }
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T off, T scale) {
        return conv(__ldg(reinterpret_cast<const uchar2*>(base) + i), w, off, scale);
    }
};
template <typename T> struct Ingest<T, KSPEC_IN_C64> {
    static constexpr int ELEM_BYTES = 8;
    typedef float2 raw_t;
    static __device__ __forceinline__ cx<T> conv(float2 v, T w, T, T) { return cscale(mkcx<T>((T)v.x, (T)v.y), w); }
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T off, T scale) {
        return conv(__ldg(reinterpret_cast<const float2*>(base) + i), w, off, scale);
    }
};
template <typename T> struct Ingest<T, KSPEC_IN_C128> {
    static constexpr int ELEM_BYTES = 16;
    typedef double2 raw_t;
    static __device__ __forceinline__ cx<T> conv(double2 v, T w, T, T) { return cscale(mkcx<T>((T)v.x, (T)v.y), w); }
    static __device__ __forceinline__ cx<T> load(const void* base, int64_t i, T w, T off, T scale) {
        return conv(__ldg(reinterpret_cast<const double2*>(base) + i), w, off, scale);
    }
};

// ---- TMA frame staging ----------------------------------------------------------------------------------------------
// The raw samples of the NEXT frame(s) are fetched by one elected thread with cp.async.bulk (global -> shared,
// completion on an mbarrier) while the CTA transforms the current frame, so the first FFT stage reads its operands
// from shared memory and never waits on HBM/L2 latency.  A bulk copy needs 16-byte aligned addresses and sizes; frame
// offsets int(i*F*r) (K:386) can be odd, so a copy covers the 16-byte granules around its frame (at most SLACK extra
// elements) and is clipped at the end of the batch: nothing outside the caller's samples is ever read.
template <typename T, int INFMT, int LOG2F, int VAR = 0> struct StageCfg {
    using C = SmemCfg<T, LOG2F, VAR>;
    static constexpr int EB = Ingest<T, INFMT>::ELEM_BYTES;
    static constexpr int SLACK = EB >= 16 ? 0 : 16 / EB;                       // elements
    static constexpr int STAGE_BYTES = ((C::F + SLACK) * EB + 127) / 128 * 128;
    static constexpr int BUDGET = 225 * 1024;
    static constexpr bool OK = (C::TEAMS == 1 || C::MULTI) && (C::DBUF || C::OCC);     // 8192 f64 staged on one buffer: 38 -> 36 GS/s
    static constexpr int STG_AUTO = !OK ? 0 : (C::MINB * (C::SMEM_BYTES + 2 * STAGE_BYTES + 1024) <= BUDGET ? 2
                                            : (C::MINB * (C::SMEM_BYTES + STAGE_BYTES + 1024) <= BUDGET ? 1 : 0));
    static constexpr int STG = (VAR == 2 || VAR == 7) ? 0 : ((C::MULTI || C::OCC) ? 1 : ((VAR == 3 || VAR == 6) ? (STG_AUTO > 1 ? 1 : STG_AUTO) : STG_AUTO));
    static constexpr int EX_BYTES = (C::SMEM_BYTES + 127) / 128 * 128;
    static constexpr int STAGE_TEAMS = C::MULTI ? C::TEAMS : 1;
    static constexpr int TW_OFS = EX_BYTES + STAGE_TEAMS * STG * STAGE_BYTES;
    static constexpr int SMEM_BYTES = TW_OFS + (C::TWSMEM ? C::F * (int)sizeof(cx<T>) : 0);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KSPEC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KSPEC_DONE;\n\t"
        "bra KSPEC_WAIT;\n\t"
        "KSPEC_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float kabs(float2 a) {
    const float2 q = __fmul2_rn(a, a);
    const float s = q.x + q.y;
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}
// |X| in float64.  The correctly rounded sqrt() costs ~8 FP64 operations plus a guarded slow path per bin and frame, one sixth
// of the FP64 work of a kernel whose bound is the FP64 pipe.  Here: the hardware reciprocal-square-root seed (MUFU.RSQ64H,
// relative error ~2^-22) and one Goldschmidt step: relative error ~2^-43 (5e-13 dB; the float64 engines are specified to
// 1e-8 dB).  Exact zeros (a silent capture) stay zero.
__device__ __forceinline__ double kabs(double2 a) {
    const double p = a.x * a.x + a.y * a.y;
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    const double g = p * r, h = 0.5 * r;
    const double e = fma(-g, h, 0.5);
    return p > 1e-290 ? fma(g, e, g) : 0.0;
}

template <typename T> __device__ __forceinline__ T pos_inf();
template <> __device__ __forceinline__ float pos_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double pos_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

// Max / Min of non-negative values as integer reductions in L2 (no return value, hence no load latency): the normalised amplitudes
// are >= 0, where IEEE order is signed-integer order on the bit patterns.  The slot is initialised by a plain store.
__device__ __forceinline__ void red_max_nonneg(float* p, float v) { asm volatile("red.global.max.s32 [%0], %1;" ::"l"(p), "r"(__float_as_int(v)) : "memory"); }
__device__ __forceinline__ void red_min_nonneg(float* p, float v) { asm volatile("red.global.min.s32 [%0], %1;" ::"l"(p), "r"(__float_as_int(v)) : "memory"); }
__device__ __forceinline__ void red_max_nonneg(double* p, double v) { asm volatile("red.global.max.s64 [%0], %1;" ::"l"(p), "l"(__double_as_longlong(v)) : "memory"); }
__device__ __forceinline__ void red_min_nonneg(double* p, double v) { asm volatile("red.global.min.s64 [%0], %1;" ::"l"(p), "l"(__double_as_longlong(v)) : "memory"); }

// Per-scan outputs from the normalised, fftshift-ed linear row in shared memory (erow[F]): thread tid walks the positions
// tid + NT i in a ROLLED loop (compact code: the unrolled form was a third of the kernel and evicted the frame loop from the
// instruction cache) with coalesced global accesses.  data_proc K:100-112, zero_span K:469-478.  The running Max/Min of this
// team (K:471-474) are kept as LINEAR amplitudes (10 log10 is monotone; stats_finish_kernel converts them exactly as the rows are
// converted here) and updated by fire-and-forget reductions.  Leaves dB - adj in erow for the waterfall compress.
template <typename T, int NT, int F>
__device__ __forceinline__ void smem_epilogue_rows(const ScanParams& p, T* erow, int64_t scan, bool valid, int64_t it, int slot, int tid) {
    T* __restrict__ rows = reinterpret_cast<T*>(p.rows);
    const bool needDb = (p.rowsKind == KSPEC_ROWS_DB) || p.wantStats || (p.hm != nullptr);
    const bool needRow = (p.hm != nullptr);
    T* __restrict__ wmax = reinterpret_cast<T*>(p.wsMax) + (int64_t)slot * F;
    T* __restrict__ wmin = reinterpret_cast<T*>(p.wsMin) + (int64_t)slot * F;
    const T gain = (T)p.gain, minAmp = (T)p.minAmp;
    const T* __restrict__ adj = reinterpret_cast<const T*>(p.adj);
    const int64_t ar = scan - (p.nScans - p.avgWin);
    T* __restrict__ avgRow = (valid && ar >= 0 && p.wantStats) ? reinterpret_cast<T*>(p.avgRows) + ar * F : nullptr;
#pragma unroll 4
    for (int jj = tid; jj < F; jj += NT) {
        T lin = erow[jj];
        if (valid && p.rowsKind == KSPEC_ROWS_LINEAR) rows[scan * F + jj] = lin;
        if (needDb) {
            if (p.dbClip) lin = fmax(lin, minAmp);
            T db = to_db(lin) - gain;
            if (p.infToZero && isinf(db)) db = (T)0;
            if (valid && p.rowsKind == KSPEC_ROWS_DB) rows[scan * F + jj] = db;
            if (p.wantStats) {
                if (it == 0) {
                    wmax[jj] = valid ? lin : (T)0;                 // idle team: identities of max / min over amplitudes
                    wmin[jj] = valid ? lin : pos_inf<T>();
                } else if (valid) {
                    red_max_nonneg(&wmax[jj], lin);
                    red_min_nonneg(&wmin[jj], lin);
                }
                if (avgRow) avgRow[jj] = db;
            }
            if (needRow) erow[jj] = adj ? db - adj[jj] : db;
        }
    }
}

// VB ("virtual bases"): the frame-parallel form for small batches.  Every frame of every scan is launched as a one-frame
// scan of its own whose first sample comes from the table p.scanBase (= scan*fullSize + frame offset), so that a single
// scan of 15..71 frames spreads over as many teams instead of walking its frames on one; frames_combine_kernel
// (epilogue.cu) then applies data_cumu over the per-frame rows.  A separate instantiation: the batch kernels stay as they are.
template <typename T, int INFMT, int LOG2F, int VAR = 0, bool VB = false>
__global__ void __launch_bounds__(SmemCfg<T, LOG2F, VAR>::CTA, SmemCfg<T, LOG2F, VAR>::MINB)
curscan_smem_kernel(const ScanParams p) {
    using C = SmemCfg<T, LOG2F, VAR>;
    using SC = StageCfg<T, INFMT, LOG2F, VAR>;
    using IN = Ingest<T, INFMT>;
    constexpr int P = C::P, F = C::F, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P, STG = SC::STG;
    constexpr int L0 = stage_l<LOG2F, LOG2P>(0);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar_all[SC::STAGE_TEAMS][2];
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    uint64_t* mbar = mbar_all[C::MULTI ? team : 0];
    const bool leader = C::MULTI ? (tid == 0) : (threadIdx.x == 0);      // issues this team's bulk copies
    // two exchange buffers per team (bufA/bufB) when they fit, else one.  Exchange x of a frame uses bufA for even
    // x and bufB for odd x; with an odd exchange count the roles swap after every frame, so a buffer is never
    // rewritten before a full barrier separates it from its last readers.
    cx<T>* bufA = reinterpret_cast<cx<T>*>(smem_raw) + team * C::FPAD;
    cx<T>* bufB = C::DBUF ? bufA + TEAMS * C::FPAD : bufA;
    unsigned char* stage0 = smem_raw + SC::EX_BYTES + (C::MULTI ? team * STG * SC::STAGE_BYTES : 0);

    const T* __restrict__ gwin = reinterpret_cast<const T*>(p.win);
    const cx<T>* __restrict__ gtw = reinterpret_cast<const cx<T>*>(p.tw);
    // team barrier: the whole CTA, or (independent teams) a named barrier over this team's NT threads
    auto sync = [team] {
        if constexpr (C::MULTI) asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(NT) : "memory");
        else __syncthreads();
        (void)team;
    };
    const cx<T>* twtab = reinterpret_cast<const cx<T>*>(p.twLin);      // linearised table (fft_core.cuh), global
    if constexpr (C::TWSMEM) {
        cx<T>* stw = reinterpret_cast<cx<T>*>(smem_raw + SC::TW_OFS);
        for (int i = threadIdx.x; i < F - (1 << L0); i += C::CTA) stw[i] = twtab[i];
        twtab = stw;
    }

    // tables that stay in registers for the life of the CTA (fast path)
    T win[C::REGTAB ? P : 1];
    cx<T> twl[(C::TWREG && C::NTW > 0) ? C::NTW : 1];
    if constexpr (C::REGTAB) {
#pragma unroll
        for (int m = 0; m < P; ++m) win[m] = gwin[tid + NT * m];
    }
    if constexpr (C::TWREG) load_twiddles<T, LOG2F, LOG2P>(twl, gtw, tid);

    const T u8off = (T)p.u8Offset, u8scale = (T)p.u8Scale;
    // AVG over the frames of a scan is the halving recurrence a_k = (a_{k-1} + m_k)/2 (data_cumu, K:137-139).  With
    // B_k = 2^k a_k it becomes B_k = B_{k-1} + 2^(k-1) m_k: one FMA per bin and frame instead of an add and a multiply, and
    // bit-identical (scaling by powers of two commutes with rounding); the 2^-(n-1) is folded into the per-scan scale.
    // Only while 2^(n-2) |X| stays far from overflow.
    const bool avgScaled = p.cumuMode == KSPEC_CUMU_AVG && p.nFrames <= (sizeof(T) == 4 ? 96 : 900);
    const T linScale = avgScaled ? (T)ldexp(p.linScale, -(p.nFrames - 1)) : (T)p.linScale;
    const int slot = blockIdx.x * TEAMS + team;                 // stats partial owned by this team
    const int64_t scansPerIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (p.nScans + scansPerIter - 1) / scansPerIter;

    // ---- staged pipeline bookkeeping (one elected thread issues the bulk copies) ---------------------------------------
    const int64_t totalFrames = iters * p.nFrames;          // frames this CTA walks through, g = it*nFrames + f
    auto issue = [&](int64_t g) {
        // called by thread 0 only: fetch frame g into stage buffer g % STG
        const int64_t it = g / p.nFrames;
        const int f = (int)(g - it * p.nFrames);
        int64_t sc = it * scansPerIter + slot;
        if (sc >= p.nScans) sc = p.nScans - 1;
        const int64_t e0 = (VB ? __ldg(&p.scanBase[sc]) : sc * p.scanStride) + p.frameOffs[f];
        // 16-byte granules covering [e0, e0+F), never past the end of the batch (the batch length is a whole number of
        // granules whenever fullSize is, which holds for every fullSize the reference can produce, K:926-929)
        constexpr int64_t GM = SC::SLACK > 0 ? SC::SLACK - 1 : 0;
        const int64_t e0a = e0 & ~GM;
        int64_t e1a = (e0 + F + GM) & ~GM;
        const int64_t total = ((VB ? p.totalElems : p.nScans * p.scanStride) + GM) & ~GM;
        if (e1a > total) e1a = total;
        const int s = (STG == 2) ? (int)(g & 1) : 0;
        const uint32_t bytes = (uint32_t)((e1a - e0a) * SC::EB);
        mbar_expect_tx(&mbar[s], bytes);
        tma_load_1d(stage0 + s * SC::STAGE_BYTES, reinterpret_cast<const unsigned char*>(p.samples) + e0a * SC::EB, bytes, &mbar[s]);
    };
    if constexpr (STG > 0) {
        if (leader) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if constexpr (STG > 0 || C::TWSMEM) __syncthreads();
    if constexpr (STG > 0) {
        if (leader) {
            for (int g = 0; g < STG && g < totalFrames; ++g) issue(g);
        }
    }

    int64_t g = 0;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t scan = it * scansPerIter + slot;
        const bool valid = scan < p.nScans;
        const int64_t scanC = valid ? scan : p.nScans - 1;      // idle teams shadow the last scan (uniform barriers)
        const int64_t sbase = VB ? __ldg(&p.scanBase[scanC]) : scanC * p.scanStride;

        T acc[P];
        T avgW = (T)1;                                           // 2^(f-1) for frame f >= 1
        for (int f = 0; f < p.nFrames; ++f, ++g) {
            const int64_t fbase = sbase + p.frameOffs[f];
            cx<T> b[P];
            if constexpr (STG > 0) {
                const int s = (STG == 2) ? (int)(g & 1) : 0;
                const uint32_t parity = (STG == 2) ? (uint32_t)((g >> 1) & 1) : (uint32_t)(g & 1);
                mbar_wait(&mbar[s], parity);
                const typename IN::raw_t* sp = reinterpret_cast<const typename IN::raw_t*>(stage0 + s * SC::STAGE_BYTES) +
                                               (SC::SLACK > 0 ? (int)(fbase & (SC::SLACK - 1)) : 0);
#pragma unroll
                for (int m = 0; m < P; ++m) {
                    const T w = C::REGTAB ? win[m] : __ldg(&gwin[tid + NT * m]);
                    b[m] = IN::conv(sp[tid + NT * m], w, u8off, u8scale);
                }
            } else {
#pragma unroll
                for (int m = 0; m < P; ++m) {
                    const T w = C::REGTAB ? win[m] : __ldg(&gwin[tid + NT * m]);
                    b[m] = IN::load(p.samples, fbase + tid + NT * m, w, u8off, u8scale);
                }
            }
            butterflies<T, P, (1 << L0), false>(b, nullptr);
            if constexpr (STG > 0) {
                // the first exchange's barrier also tells thread 0 that every thread has consumed the stage buffer
                auto sync_and_refill = [&] {
                    sync();
                    if (leader && g + STG < totalFrames) issue(g + STG);
                };
                fft_tail_first<T, LOG2F, LOG2P, C::TWREG, C::DBUF, L0, !C::TWSMEM>(b, twl, twtab, bufA, bufB, tid, sync, sync_and_refill);
            } else {
                fft_tail<T, LOG2F, LOG2P, C::TWREG, C::DBUF, L0, 0, 0, !C::TWSMEM>(b, twl, twtab, bufA, bufB, tid, sync);
            }
            if constexpr (C::DBUF && (C::NX & 1)) { cx<T>* t = bufA; bufA = bufB; bufB = t; }
            if constexpr (sizeof(T) == 8) {
                // bUsePSD (K:374-384): sum of |X|^2 over the segments; mean and 1/(Fs*sum(w^2)) are in linScale
                if (p.cumuMode == KSPEC_CUMU_PSD) {
#pragma unroll
                    for (int m = 0; m < P; ++m) {
                        const T pw = b[m].x * b[m].x + b[m].y * b[m].y;
                        acc[m] = (f == 0) ? pw : acc[m] + pw;
                    }
                    continue;
                }
            }
            // |X| and cumulate over the frames of this scan (data_cumu, K:124-147); normalisation is applied once per scan
            if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = kabs(b[m]);
            } else if (avgScaled) {
#pragma unroll
                for (int m = 0; m < P; ++m) acc[m] = fma(kabs(b[m]), avgW, acc[m]);
                avgW += avgW;
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
        // the normalised row goes through shared memory (bufA: the buffer the last exchange did NOT use, see above) in
        // fftshift-ed order; smem_epilogue_rows walks it by position
        const bool needRow = (p.hm != nullptr);
        T* erow = reinterpret_cast<T*>(bufA);
        if (!C::DBUF) sync();
#pragma unroll
        for (int m = 0; m < P; ++m) erow[(tid + NT * m) ^ (F >> 1)] = acc[m] * linScale;
        sync();
        smem_epilogue_rows<T, NT, F>(p, erow, scan, valid, it, slot, tid);
        if (needRow) {
            sync();
            // _data_plotcompress (K:184-200): W groups of g adjacent bins
            const int W = p.hmW, gsz = F / W;
            T* __restrict__ hm = reinterpret_cast<T*>(p.hm);
            const T* row = erow;
            for (int w = tid; w < W; w += NT) {
                T r = row[w * gsz];
                if (p.hmMode == KSPEC_COMPRESS_MAX) {
                    for (int q = 1; q < gsz; ++q) r = fmax(r, row[w * gsz + q]);
                } else if (p.hmMode == KSPEC_COMPRESS_MIN) {
                    for (int q = 1; q < gsz; ++q) r = fmin(r, row[w * gsz + q]);
                } else if (p.hmMode == KSPEC_COMPRESS_AVG) {
                    for (int q = 1; q < gsz; ++q) r += row[w * gsz + q];
                    r /= (T)gsz;
                }
                if (valid) hm[scan * W + w] = r;
            }
        }
        sync();                                                  // the scratch row is an exchange buffer of the next frame
    }
}

template <typename T, int INFMT, int LOG2F, int VAR = 0, bool VB = false>
static int launch_smem_one(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    using C = SmemCfg<T, LOG2F, VAR>;
    using SC = StageCfg<T, INFMT, LOG2F, VAR>;
    auto k = curscan_smem_kernel<T, INFMT, LOG2F, VAR, VB>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SC::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    if (info) {
        info->ctaThreads = C::CTA;
        info->smemBytes = SC::SMEM_BYTES;
        info->teams = C::TEAMS;
        info->stages = SC::STG;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, C::CTA, SC::SMEM_BYTES);
        info->ctasPerSm = nb;
        return 0;
    }
    k<<<grid, C::CTA, SC::SMEM_BYTES, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace kspec
