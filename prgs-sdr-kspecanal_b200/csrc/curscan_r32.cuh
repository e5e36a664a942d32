// curscan_r32.cuh — the headline kernel: fused hot path for fftSize 2048 in float32 on large batches.
//
// Same job as curscan_smem_kernel (curscan_smem.cuh) — sdr_curscan K:385-397, data_cumu K:124-147, data_proc K:100-112,
// zero_span stats K:471-476, _data_plotcompress K:168-202 for a whole batch of scans in one launch — with a transform
// layout chosen for the pipe that bounds that kernel, the shared-memory data path (profiles/README.md):
//
//   2048 = 32 x 2 x 32.  A team of 64 threads (two warps) owns one frame, 32 complex points per thread.
//     stage 0   thread j holds x[j + 64 m], m = 0..31 (ingest + window fused into the load): DFT32 over m in registers,
//               then the boundary twiddle W_2048^(j k1) from a shared-memory table laid out [slot][thread].
//     radix 2   the 64-point transform over j starts with the butterflies (j, j + 32).  The two threads sit 16 lanes
//               apart in one warp and swap HALF of their values with shfl.xor: u = a + b, v = (a - b) W_64^j.
//               The upper thread's stage-0 outputs are rotated by 16 slots (its window carries the sign (-1)^m), so both
//               threads send slots 16..31 and keep 0..15: no lane-dependent register selection anywhere.
//     exchange  ONE pass through shared memory (the 16/16/8 layout needs two): row rho = k1 + 32 c, column j mod 32.
//     stage 1   thread rho: DFT32 over the row -> bins rho + 64 kappa, kappa = 0..31, i.e. the same "thread t owns
//               bins t + NT m" convention as the other kernels, so |X|, cumulate and the per-scan epilogue are shared.
//
//   Per frame and SM this is ~576 shared-memory wavefronts (stage read 128, twiddles 128, shuffles 64, exchange 256)
//   instead of ~890, one team barrier pair over two warps instead of two over four, and the same ~85 k FP32 lane
//   operations.  Six teams (384 threads, <= 168 registers) share one SM; each team walks scans slot, slot + G, ...
//   tools/r32_model.py is the index-level numpy model of this data flow (checked against numpy.fft).
#pragma once
#include "curscan_smem.cuh"

namespace kspec {

struct R32Cfg {
    static constexpr int F = 2048, LOG2F = 11, P = 32, NT = 64, TEAMS = 6, CTA = NT * TEAMS;
    static constexpr int PITCH = 34;                               // exchange row pitch (complex): conflict-free 128-bit row reads
    static constexpr int EX_BYTES = 64 * PITCH * 8;                // one exchange buffer per team (17 408 B)
    static constexpr int TW_BYTES = P * NT * 8;                    // boundary twiddles [slot pair][thread][2]
    static constexpr int MAX_FRAMES = 1024;                        // frame offsets of a scan are kept in shared memory
};

template <int INFMT> struct R32Stage {
    static constexpr int EB = Ingest<float, INFMT>::ELEM_BYTES;
    static constexpr int SLACK = EB >= 16 ? 0 : 16 / EB;
    static constexpr int STAGE_BYTES = ((R32Cfg::F + SLACK) * EB + 127) / 128 * 128;
    static constexpr int TEAM_BYTES = R32Cfg::EX_BYTES + STAGE_BYTES;
    static constexpr int TW_OFS = R32Cfg::TEAMS * TEAM_BYTES;
    static constexpr int SMEM_BYTES = TW_OFS + R32Cfg::TW_BYTES;
    static constexpr bool OK = SMEM_BYTES <= 227 * 1024;
};

// exp(-2 pi i n / 32)
__device__ __forceinline__ constexpr double root32_re(int n) {
    constexpr double c[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                             0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0};
    n &= 31;
    return n <= 8 ? c[n] : (n <= 16 ? -c[16 - n] : (n <= 24 ? -c[n - 16] : c[32 - n]));
}
__device__ __forceinline__ constexpr double root32_im(int n) { return -root32_re((n + 24) & 31); }   // -sin(x) = -cos(x - pi/2)

// DFT8 of x[i] * w[i]: the window multiply is folded into the first butterfly layer (a*wa +- b*wb = one multiply and two
// fused multiply-adds instead of two multiplies and two adds)
__device__ __forceinline__ void dft8_win(float2 (&x)[8], const float (&w)[8]) {
    const float h = 0.70710678118654752440f;
    float2 s[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = cscale(x[i + 4], w[i + 4]);
        const float2 wi = make_float2(w[i], w[i]);
        s[i] = __ffma2_rn(x[i], wi, t);
        d[i] = __ffma2_rn(x[i], wi, make_float2(-t.x, -t.y));
    }
    // evens: dft4(x0, x2, x4, x6);  odds: dft4(x1, x3, x5, x7)   (first layer done: s = a + b, d = a - b)
    float2 e0 = s[0] + s[2], e2 = s[0] - s[2], e1 = d[0] + mul_mi(d[2]), e3 = d[0] - mul_mi(d[2]);
    float2 o0 = s[1] + s[3], q2 = s[1] - s[3], q1 = d[1] + mul_mi(d[3]), q3 = d[1] - mul_mi(d[3]);
    const float2 o1 = cmul(q1, make_float2(h, -h));
    const float2 o2 = mul_mi(q2);
    const float2 o3 = cmul(q3, make_float2(-h, -h));
    x[0] = e0 + o0; x[4] = e0 - o0;
    x[1] = e1 + o1; x[5] = e1 - o1;
    x[2] = e2 + o2; x[6] = e2 - o2;
    x[3] = e3 + o3; x[7] = e3 - o3;
}

// 32-point forward DFT: 4 x DFT8 over n1 (n = 4 n1 + n2), twiddles W_32^(n2 k1), 8 x DFT4 over n2.
// dft32_head does everything but the last layer; WIN: the inputs are multiplied by w[] on the way in.
template <bool WIN>
__device__ __forceinline__ void dft32_head(float2 (&x)[32], const float (&w)[32]) {
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        float2 y[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) y[n1] = x[4 * n1 + n2];
        if constexpr (WIN) {
            float wy[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) wy[n1] = w[4 * n1 + n2];
            dft8_win(y, wy);
        } else {
            dft8<float>(y);
        }
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) x[4 * k1 + n2] = y[k1];
    }
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1)
#pragma unroll
        for (int n2 = 1; n2 < 4; ++n2) {
            const int q = n2 * k1;
            if (q == 8) x[4 * k1 + n2] = mul_mi(x[4 * k1 + n2]);
            else x[4 * k1 + n2] = cmul(x[4 * k1 + n2], make_float2((float)root32_re(q), (float)root32_im(q)));
        }
}
// natural order in and out
__device__ __forceinline__ void dft32(float2 (&x)[32]) {
    const float none[32] = {};
    dft32_head<false>(x, none);
    float2 o[32];
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        dft4<float>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) o[k1 + 8 * k2] = x[4 * k1 + k2];
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = o[k];
}

// Stage 0 of a frame for one thread: windowed DFT32 of its 32 samples, boundary twiddle W_2048^(j k1), the radix-2 butterflies
// with the thread 16 lanes away (it sends its slots 16..31 and keeps 0..15) and the exchange writes, as ONE software pipeline
// over the eight groups of the transform's last layer: group k1 finishes slots {k1, k1+8, k1+16, k1+24}, fetches their four
// twiddles with two 128-bit loads, swaps (k1+16, k1+24) with the partner and stores u (row s + 16 upper) and v (row + 32).
// The shared-memory and shuffle latencies of one group hide behind the arithmetic of the others.
__device__ __forceinline__ void r32_stage0(float2 (&b)[32], const float (&win)[32], const float4* tw4, float2 omega, float2* exw) {
    constexpr int NT = R32Cfg::NT, PITCH = R32Cfg::PITCH;
    dft32_head<true>(b, win);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        dft4<float>(b[4 * k1], b[4 * k1 + 1], b[4 * k1 + 2], b[4 * k1 + 3]);     // -> slots k1 + 8 k2
        const float4 tA = tw4[(2 * k1) * NT], tB = tw4[(2 * k1 + 1) * NT];
        const float2 y0 = cmul(b[4 * k1], make_float2(tA.x, tA.y));               // slot k1
        const float2 y1 = cmul(b[4 * k1 + 1], make_float2(tA.z, tA.w));           // slot k1 + 8
        const float2 y2 = cmul(b[4 * k1 + 2], make_float2(tB.x, tB.y));           // slot k1 + 16
        const float2 y3 = cmul(b[4 * k1 + 3], make_float2(tB.z, tB.w));           // slot k1 + 24
        float2 r0, r1;
        r0.x = __shfl_xor_sync(0xffffffffu, y2.x, 16);
        r0.y = __shfl_xor_sync(0xffffffffu, y2.y, 16);
        r1.x = __shfl_xor_sync(0xffffffffu, y3.x, 16);
        r1.y = __shfl_xor_sync(0xffffffffu, y3.y, 16);
        exw[k1 * PITCH] = y0 + r0;                                                // u, s = k1
        exw[(k1 + 32) * PITCH] = cmul(y0 - r0, omega);                            // v
        exw[(k1 + 8) * PITCH] = y1 + r1;                                          // u, s = k1 + 8
        exw[(k1 + 40) * PITCH] = cmul(y1 - r1, omega);                            // v
    }
}

// raw sample -> float2 without scale or window (both live in the window registers of the R32 kernels)
template <int INFMT> struct R32Raw;
template <> struct R32Raw<KSPEC_IN_U8_IQ> {
    typedef unsigned short raw_t;                 // I in the low byte, Q in the high byte (octave/load_rtlsdr.m:8-12)
    // byte -> float without the conversion unit (I2F runs at a quarter of the FP32 rate): 0x4B000000 | b is the float
    // 2^23 + b exactly, one byte-permute per component, then two packed subtractions; same value as (float)b - off
    static __device__ __forceinline__ float2 get(unsigned short v, float off) {
        const float2 m = make_float2(__uint_as_float(__byte_perm((unsigned)v, 0x4B000000u, 0x7540)),
                                     __uint_as_float(__byte_perm((unsigned)v, 0x4B000000u, 0x7541)));
        const float2 b = __fadd2_rn(m, make_float2(-8388608.0f, -8388608.0f));
        return __fadd2_rn(b, make_float2(-off, -off));
    }
};
template <> struct R32Raw<KSPEC_IN_C64> {
    typedef float2 raw_t;
    static __device__ __forceinline__ float2 get(float2 v, float) { return v; }
};

// |X| with one multiply and one fused multiply-add in front of the square root
__device__ __forceinline__ float kabs_fma(float2 a) {
    const float s = fmaf(a.y, a.y, a.x * a.x);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ float ld_hint(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// Max / Min of non-negative floats as integer reductions in L2 (no return value, hence no load latency): the normalised
// amplitudes are >= 0, where float order is signed-integer order.  The slot is initialised by a plain store.
__device__ __forceinline__ void red_max_pos(float* p, float v, uint64_t pol) {
    asm volatile("red.global.max.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(__float_as_int(v)), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_min_pos(float* p, float v, uint64_t pol) {
    asm volatile("red.global.min.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(__float_as_int(v)), "l"(pol) : "memory");
}

// Per-scan outputs from the normalised, fftshift-ed linear row in shared memory (erow[F]).  data_proc K:100-112, zero_span
// K:469-478, _data_plotcompress K:184-200.  The running Max/Min of this team (K:471-474) are kept as LINEAR amplitudes
// (10 log10 is monotone; stats_finish_kernel converts them exactly as the rows are converted here).
// Returns true when the waterfall row has been written as well (fast path), false when dB - adj is left in erow for the
// caller's compress step.
template <int NT>
__device__ __forceinline__ bool scan_epilogue_rows(const ScanParams& p, float* erow, int64_t scan, bool valid, int64_t it, int slot, int tid,
                                                   uint64_t polKeep) {
    constexpr int F = R32Cfg::F;
    float* __restrict__ rows = reinterpret_cast<float*>(p.rows);
    const bool needDb = (p.rowsKind == KSPEC_ROWS_DB) || p.wantStats || (p.hm != nullptr);
    const bool needRow = (p.hm != nullptr);
    float* __restrict__ wmax = reinterpret_cast<float*>(p.wsMax) + (int64_t)slot * F;
    float* __restrict__ wmin = reinterpret_cast<float*>(p.wsMin) + (int64_t)slot * F;
    const float gain = (float)p.gain, minAmp = (float)p.minAmp;
    const float* __restrict__ adj = reinterpret_cast<const float*>(p.adj);
    const int64_t ar = scan - (p.nScans - p.avgWin);
    float* __restrict__ avgRow = (valid && ar >= 0 && p.wantStats) ? reinterpret_cast<float*>(p.avgRows) + ar * F : nullptr;

    // Fast path — the streaming case: no rows wanted, Max/Min partials, a MAX or MIN waterfall row of 4-bin groups, no
    // baseline, no clip.  Max/Min and the group reduce work on the linear values (monotone map), so only one dB conversion per
    // waterfall bin is left: 4 instructions per bin instead of ~40.
    if (p.rowsKind == KSPEC_ROWS_NONE && p.wantStats && needRow && adj == nullptr && avgRow == nullptr && !p.dbClip && !p.infToZero &&
        p.hmW * 4 == F && (p.hmMode == KSPEC_COMPRESS_MAX || p.hmMode == KSPEC_COMPRESS_MIN)) {
        float* __restrict__ hm = reinterpret_cast<float*>(p.hm) + scan * p.hmW;
        const bool useMax = p.hmMode == KSPEC_COMPRESS_MAX;
#pragma unroll 2
        for (int q = tid; q < F / 4; q += NT) {
            const float4 v = reinterpret_cast<const float4*>(erow)[q];
            float* mx = wmax + 4 * q;
            float* mn = wmin + 4 * q;
            if (it == 0) {
                const float4 hi = valid ? v : make_float4(0.0f, 0.0f, 0.0f, 0.0f);            // idle team: identity of max over amplitudes
                const float4 lo = valid ? v : make_float4(pos_inf<float>(), pos_inf<float>(), pos_inf<float>(), pos_inf<float>());
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(mx), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w), "l"(polKeep) : "memory");
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(mn), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "l"(polKeep) : "memory");
            } else if (valid) {
                red_max_pos(mx, v.x, polKeep); red_max_pos(mx + 1, v.y, polKeep); red_max_pos(mx + 2, v.z, polKeep); red_max_pos(mx + 3, v.w, polKeep);
                red_min_pos(mn, v.x, polKeep); red_min_pos(mn + 1, v.y, polKeep); red_min_pos(mn + 2, v.z, polKeep); red_min_pos(mn + 3, v.w, polKeep);
            }
            const float g4 = useMax ? fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) : fminf(fminf(v.x, v.y), fminf(v.z, v.w));
            if (valid) hm[q] = to_db(g4) - gain;
        }
        return true;
    }
#pragma unroll 4
    for (int jj = tid; jj < F; jj += NT) {
        float lin = erow[jj];
        if (valid && p.rowsKind == KSPEC_ROWS_LINEAR) rows[scan * F + jj] = lin;
        if (needDb) {
            if (p.dbClip) lin = fmaxf(lin, minAmp);
            float db = to_db(lin) - gain;
            if (p.infToZero && isinf(db)) db = 0.0f;
            if (valid && p.rowsKind == KSPEC_ROWS_DB) rows[scan * F + jj] = db;
            if (p.wantStats) {
                if (it == 0) {
                    st_hint(&wmax[jj], valid ? lin : 0.0f, polKeep);
                    st_hint(&wmin[jj], valid ? lin : pos_inf<float>(), polKeep);
                } else if (valid) {
                    red_max_pos(&wmax[jj], lin, polKeep);
                    red_min_pos(&wmin[jj], lin, polKeep);
                }
                if (avgRow) avgRow[jj] = db;
            }
            if (needRow) erow[jj] = adj ? db - adj[jj] : db;
        }
    }
    return false;
}

// boundary twiddles W_2048^(j k1) in shared memory: slot s of thread t holds k1 = (s + 16 upper) mod 32 after stage 0; the table
// is laid out [2 (s & 7) + (s >> 4)][t][(s >> 3) & 1]: the slots {g, g+8} and {g+16, g+24} of a last-layer group are one
// 128-bit load each
__device__ __forceinline__ void r32_build_twiddles(float2* stw, const float2* __restrict__ gtw, int thread, int nthreads) {
    constexpr int NT = R32Cfg::NT, P = R32Cfg::P, F = R32Cfg::F;
    for (int i = thread; i < P * NT; i += nthreads) {
        const int s = i / NT, t = i % NT;
        const int up = (t & 31) >> 4, jj = (t & 15) + 16 * (t >> 5) + 32 * up;
        stw[((2 * (s & 7) + (s >> 4)) * NT + t) * 2 + ((s >> 3) & 1)] = gtw[(jj * ((s + 16 * up) & 31)) & (F - 1)];
    }
}

// DYN: scans beyond a team's first are tickets from a global counter instead of the static stride (see below)
template <int INFMT, bool DYN>
__global__ void __launch_bounds__(R32Cfg::CTA, 1) curscan_r32_kernel(const ScanParams p) {
    using C = R32Cfg;
    using SC = R32Stage<INFMT>;
    using IN = R32Raw<INFMT>;
    constexpr int P = C::P, F = C::F, NT = C::NT, TEAMS = C::TEAMS, PITCH = C::PITCH;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar_all[TEAMS];
    __shared__ int32_t nextScanAll[TEAMS];                         // next scan of each team (dynamic scheduling, see below)
    __shared__ int32_t foffs[C::MAX_FRAMES];                       // K:386 frame starts inside a scan
    const int team = threadIdx.x / NT;
    const int tid = threadIdx.x % NT;                              // = rho in stage 1: owns bins tid + 64 kappa
    const int lane = tid & 31;
    const int upper = lane >> 4;                                   // partner = lane ^ 16
    const int jp = (lane & 15) + 16 * (tid >> 5);                  // j mod 32
    const int j = jp + 32 * upper;                                 // stage 0: owns x[j + 64 m]
    float2* ex = reinterpret_cast<float2*>(smem_raw + team * SC::TEAM_BYTES);
    unsigned char* stage = smem_raw + team * SC::TEAM_BYTES + C::EX_BYTES;
    float2* stw = reinterpret_cast<float2*>(smem_raw + SC::TW_OFS);
    uint64_t* mbar = &mbar_all[team];
    const bool leader = tid == 0;
    auto sync = [team] { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(NT) : "memory"); };

    const float* __restrict__ gwin = reinterpret_cast<const float*>(p.win);
    const float2* __restrict__ gtw = reinterpret_cast<const float2*>(p.tw);      // exp(-2 pi i k / 2048)
    // boundary twiddles: slot s of thread t holds k1 = (s + 16 upper) mod 32 after stage 0
    r32_build_twiddles(stw, gtw, threadIdx.x, C::CTA);
    for (int i = threadIdx.x; i < p.nFrames; i += C::CTA) foffs[i] = p.frameOffs[i];
    // window in registers (the uint8 scale rides in it); the upper threads carry (-1)^m: their DFT32 outputs come out rotated
    // by 16 slots
    const float u8off = (float)p.u8Offset;
    const float wscale = INFMT == KSPEC_IN_U8_IQ ? (float)p.u8Scale : 1.0f;
    float win[P];
#pragma unroll
    for (int m = 0; m < P; ++m) {
        const float w = gwin[j + NT * m];
        win[m] = ((upper && (m & 1)) ? -w : w) * wscale;
    }
    float2 omega = gtw[32 * jp];                                   // W_64^(j mod 32); the upper thread computes b - a
    if (upper) omega = make_float2(-omega.x, -omega.y);

    const bool avgScaled = p.cumuMode == KSPEC_CUMU_AVG && p.nFrames <= 96;       // see curscan_smem.cuh
    const float linScale = avgScaled ? (float)ldexp(p.linScale, -(p.nFrames - 1)) : (float)p.linScale;
    const int slot = blockIdx.x * TEAMS + team;
    // Team `slot` starts with scan `slot`.  Static form: it goes on with slot + nSlots, ...  Dynamic form (DYN): every further
    // scan is a ticket from a global counter (one atomic per scan, fetched a whole scan ahead), because a static partition makes
    // the launch as long as its slowest team -- a team whose SM was still busy with another kernel (the NCCL exchange of the
    // previous batch of a sharded capture) when the grid started, or the last wave of a batch that is far from a multiple of
    // the team count.  The launcher picks the form (the ticket costs ~2 % on a full, undisturbed launch).
    const int nSlots = (int)gridDim.x * TEAMS;
    const int nScans32 = (int)p.nScans;                           // the launcher keeps batches below 2^31 scans
    // L2 policies are created where they are used (one instruction; two registers each if they were kept): the samples stream
    // through L2 once (evict first), the per-team Max/Min partials stay L2 resident (evict last)

    // leader only: fetch frame f of scan sc into the stage buffer.  General mode: one bulk copy of the 16-byte granules around
    // the frame (offsets can be odd), clipped at the end of the batch (see curscan_smem.cuh).  Ring mode (every frame starts
    // F/2 after the previous one, i.e. 50 % overlap): the stage buffer is two half-frame slots, hop h lives in slot h & 1, and
    // only the NEW hop of a frame is copied: each sample crosses L2 -> shared memory once.
    const int64_t totalElems = p.nScans * p.scanStride;
    const bool ring = p.hopRing != 0;
    constexpr int HALF_BYTES = (F / 2) * SC::EB;
    auto issue = [&](int64_t sc, int f) {
        const uint64_t polStream = l2_policy_evict_first();
        if (ring) {
            const int h = f == 0 ? 0 : f + 1;                      // first hop to fetch
            const uint32_t bytes = f == 0 ? 2 * HALF_BYTES : HALF_BYTES;
            const int64_t e0 = sc * p.scanStride + (int64_t)h * (F / 2);
            mbar_expect_tx(mbar, bytes);
            tma_load_1d_hint(stage + (h & 1) * HALF_BYTES, reinterpret_cast<const unsigned char*>(p.samples) + e0 * SC::EB, bytes, mbar, polStream);
            return;
        }
        constexpr int64_t GM = SC::SLACK > 0 ? SC::SLACK - 1 : 0;
        const int64_t e0 = sc * p.scanStride + foffs[f];
        const int64_t e0a = e0 & ~GM;
        int64_t e1a = (e0 + F + GM) & ~GM;
        const int64_t total = (totalElems + GM) & ~GM;
        if (e1a > total) e1a = total;
        const uint32_t bytes = (uint32_t)((e1a - e0a) * SC::EB);
        mbar_expect_tx(mbar, bytes);
        tma_load_1d_hint(stage, reinterpret_cast<const unsigned char*>(p.samples) + e0a * SC::EB, bytes, mbar, polStream);
    };
    if (leader) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (slot >= p.nScans) {
        // a team without work still owns a Max/Min partial: the identities of max / min over amplitudes
        if (p.wantStats) {
            float* wmax = reinterpret_cast<float*>(p.wsMax) + (int64_t)slot * F;
            float* wmin = reinterpret_cast<float*>(p.wsMin) + (int64_t)slot * F;
            for (int i = tid; i < F; i += NT) { wmax[i] = 0.0f; wmin[i] = pos_inf<float>(); }
        }
        return;
    }
    if (leader) issue(slot, 0);

    uint32_t g = 0;                                                // frames this team has processed (mbarrier parity)
    bool firstScan = true;
    int nxtLeader = 0;
    for (int scan = slot; scan < nScans32; scan = DYN ? nextScanAll[team] : scan + nSlots) {
        const bool valid = true;
        const int64_t it = firstScan ? 0 : 1;
        firstScan = false;
        const int64_t sbase = (int64_t)scan * p.scanStride;

        float acc[P];
        float avgW = 1.0f;
        for (int f = 0; f < p.nFrames; ++f, ++g) {
            float2 b[P];
            {
                mbar_wait(mbar, g & 1);
                const int mis = (SC::SLACK > 0 && !ring) ? (((int)sbase + foffs[f]) & (SC::SLACK - 1)) : 0;
                const typename IN::raw_t* sp0 = reinterpret_cast<const typename IN::raw_t*>(stage + ((ring && (f & 1)) ? HALF_BYTES : 0)) + mis + j;
                const typename IN::raw_t* sp1 = reinterpret_cast<const typename IN::raw_t*>(stage + ((ring && (f & 1)) ? 0 : HALF_BYTES)) + mis + j;
#pragma unroll
                for (int m = 0; m < P / 2; ++m) b[m] = IN::get(sp0[NT * m], u8off);
#pragma unroll
                for (int m = 0; m < P / 2; ++m) b[P / 2 + m] = IN::get(sp1[NT * m], u8off);
            }
            // every thread of the team is past its stage reads and (program order) past the previous frame's exchange reads
            sync();
            if (leader) {
                if constexpr (DYN) {
                    if (f == 0) nxtLeader = (int)atomicAdd(p.scanCounter, 1u) + nSlots;      // needed a whole scan from now
                } else {
                    nxtLeader = scan + nSlots;
                }
                if (f + 1 < p.nFrames) {
                    issue(scan, f + 1);
                } else {
                    if constexpr (DYN) nextScanAll[team] = nxtLeader;                     // read by the team after the epilogue's barriers
                    if (nxtLeader < nScans32) issue(nxtLeader, 0);
                }
            }
            r32_stage0(b, win, reinterpret_cast<const float4*>(stw) + tid, omega, ex + (16 * upper) * PITCH + jp);
            sync();
            {
                const float4* q4 = reinterpret_cast<const float4*>(ex + tid * PITCH);
#pragma unroll
                for (int i = 0; i < P; i += 2) {
                    const float4 v = q4[i >> 1];
                    b[i] = make_float2(v.x, v.y);
                    b[i + 1] = make_float2(v.z, v.w);
                }
            }
            dft32(b);
            // |X| in the same basic block as the transform's last layer (the square roots overlap the butterflies), then
            // cumulate over the frames of this scan (data_cumu, K:124-147); normalisation once per scan
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

        // ---------------- per-scan epilogue ---------------------------------------------------------------------------------
        // The normalised row goes through shared memory in fftshift-ed order (bin k -> position k ^ F/2) and a rolled loop
        // walks the positions tid + 64 i: compact code (the frame loop stays resident in the instruction cache) and
        // coalesced global accesses.  The running Max/Min of this team (K:471-474) are fire-and-forget reductions in L2.
        float* erow = reinterpret_cast<float*>(ex);
        sync();                                                    // the last frame's exchange reads are done
#pragma unroll
        for (int m = 0; m < P; ++m) erow[(tid + NT * m) ^ (F >> 1)] = acc[m] * linScale;
        sync();
        const bool hmDone = scan_epilogue_rows<NT>(p, erow, scan, valid, it, slot, tid, l2_policy_evict_last());
        if (p.hm != nullptr && !hmDone) {
            sync();
            // _data_plotcompress (K:184-200): W groups of adjacent bins
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
            // the next frame's exchange writes come after that frame's first team barrier: no barrier needed here
        }
    }
}

template <int INFMT, bool DYN = false>
static int launch_r32(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info) {
    using SC = R32Stage<INFMT>;
    if constexpr (!SC::OK) {
        return (int)cudaErrorInvalidValue;
    } else {
        auto k = curscan_r32_kernel<INFMT, DYN>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SC::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        if (info) {
            info->ctaThreads = R32Cfg::CTA;
            info->smemBytes = SC::SMEM_BYTES;
            info->teams = R32Cfg::TEAMS;
            info->stages = 1;
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, R32Cfg::CTA, SC::SMEM_BYTES);
            info->ctasPerSm = nb;
            return 0;
        }
        k<<<grid, R32Cfg::CTA, SC::SMEM_BYTES, st>>>(p);
        return (int)cudaGetLastError();
    }
}

}  // namespace kspec
