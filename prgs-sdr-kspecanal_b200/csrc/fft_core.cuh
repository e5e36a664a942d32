// fft_core.cuh — register-resident Stockham FFT building blocks for sm_100a.
//
// Replaces numpy's pocketfft call at kspecanal.py:391 (np.fft.fft of one windowed frame).  Written from
// the Stockham autosort formulation; nothing here derives from pocketfft or the reference.
//
// Layout contract used by every kernel in this library.  A team of NT = F/P threads owns one F-point frame;
// thread `tid` holds P complex values b[m], m = 0..P-1, that stand for element  tid + NT*m  of the current
// stage's input vector (and, after the last stage, for output bin  tid + NT*m).  A stage of radix R treats the
// P registers as V = P/R butterflies; butterfly v uses registers b[v + t*V], t = 0..R-1, and is "virtual thread"
// j = tid + v*NT of the textbook formulation:
//      in [j + t*F/R]                                   (== tid + NT*(v + t*V): the registers above)
//      twiddle  exp(-2 pi i * t * (j mod Ns) / (Ns*R))  (Ns = product of earlier radices; stage 0 has none)
//      out[(j div Ns)*Ns*R + (j mod Ns) + t*Ns]
// Between stages the values go through a padded shared-memory buffer (one extra element per 16 keeps both the
// scattered writes and the unit-stride reads conflict free for 8- and 16-byte elements).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kspec {

template <typename T> struct CxOf;
template <> struct CxOf<float>  { using type = float2; };
template <> struct CxOf<double> { using type = double2; };
template <typename T> using cx = typename CxOf<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mkcx(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }

// complex float = one 64-bit register pair (re, im): Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2, PTX
// add/mul/fma.f32x2) operate on the pair in one issue slot, and their operand modifiers make the complex idioms free:
// negation, half swap with per-half sign (".LO_HI.NP": multiplication by +-i) and scalar broadcast (".F32").
//   complex add / sub        1 FADD2            (scalar code: 2 FADD)
//   a + (+-i) b              1 FADD2            (2 FADD)
//   real scale               1 FMUL2            (2 FMUL)
//   complex multiply         FMUL2 + FFMA2      (2 FMUL + 2 FFMA)
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { double2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { double2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
__device__ __forceinline__ double2 cmul(double2 a, double2 w) {
    double2 r;
    r.x = a.x * w.x - a.y * w.y;
    r.y = a.x * w.y + a.y * w.x;
    return r;
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { a.x *= s; a.y *= s; return a; }
template <typename C> __device__ __forceinline__ C mul_mi(C a) { C r; r.x = a.y;  r.y = -a.x; return r; }  // a * (-i)
template <typename C> __device__ __forceinline__ C mul_pi(C a) { C r; r.x = -a.y; r.y = a.x;  return r; }  // a * (+i)

// ---- small forward DFTs, natural order in and out ------------------------------------------------------------
template <typename T> __device__ __forceinline__ void dft2(cx<T>& a, cx<T>& b) {
    cx<T> s = a + b, d = a - b;
    a = s; b = d;
}

template <typename T> __device__ __forceinline__ void dft4(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3) {
    cx<T> a = x0 + x2, b = x0 - x2, c = x1 + x3, d = mul_mi(x1 - x3);
    x0 = a + c; x2 = a - c; x1 = b + d; x3 = b - d;
}

template <typename T> __device__ __forceinline__ void dft8(cx<T> (&x)[8]) {
    const T h = (T)0.70710678118654752440;
    // evens and odds
    dft4<T>(x[0], x[2], x[4], x[6]);
    dft4<T>(x[1], x[3], x[5], x[7]);
    // odd outputs times W8^k, k = 0..3
    cx<T> o1 = cmul(x[3], mkcx<T>(h, -h));                                // * (1-i)/sqrt2
    cx<T> o2 = mul_mi(x[5]);                                              // * -i
    cx<T> o3 = cmul(x[7], mkcx<T>(-h, -h));                               // * (-1-i)/sqrt2
    cx<T> e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1];
    x[0] = e0 + o0; x[4] = e0 - o0;
    x[1] = e1 + o1; x[5] = e1 - o1;
    x[2] = e2 + o2; x[6] = e2 - o2;
    x[3] = e3 + o3; x[7] = e3 - o3;
}

template <typename T> __device__ __forceinline__ void dft16(cx<T> (&x)[16]) {
    // n = 4*n1 + n2, k = k1 + 4*k2
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<T>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);   // -> y[k1][n2] at x[4*k1+n2]
    // twiddles W16^(n2*k1)
    x[5]  = cmul(x[5],  mkcx<T>(c1, -s1));                                    // W^1
    x[6]  = cmul(x[6],  mkcx<T>(h, -h));                                      // W^2
    x[7]  = cmul(x[7],  mkcx<T>(s1, -c1));                                    // W^3
    x[9]  = cmul(x[9],  mkcx<T>(h, -h));                                      // W^2
    x[10] = mul_mi(x[10]);                                                    // W^4
    x[11] = cmul(x[11], mkcx<T>(-h, -h));                                     // W^6
    x[13] = cmul(x[13], mkcx<T>(s1, -c1));                                    // W^3
    x[14] = cmul(x[14], mkcx<T>(-h, -h));                                     // W^6
    x[15] = cmul(x[15], mkcx<T>(-c1, s1));                                    // W^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<T>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);  // -> X[k1+4*k2] at x[4*k1+k2]
    // transpose 4x4 so that x[k] is bin k
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { cx<T> t = x[4 * a + b]; x[4 * a + b] = x[4 * b + a]; x[4 * b + a] = t; }
}

template <typename T, int R> __device__ __forceinline__ void dftR(cx<T> (&x)[R]) {
    if constexpr (R == 2) dft2<T>(x[0], x[1]);
    else if constexpr (R == 4) dft4<T>(x[0], x[1], x[2], x[3]);
    else if constexpr (R == 8) dft8<T>(x);
    else { static_assert(R == 16, "radix"); dft16<T>(x); }
}

// ---- compile-time stage schedule ---------------------------------------------------------------------------------
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }

// log2 radix of the stage that starts with log2(Ns) == lns
template <int LOG2F, int LOG2P> __host__ __device__ constexpr int stage_l(int lns) { return cmin(LOG2P, LOG2F - lns); }

// number of twiddle factors a thread needs for all stages >= 1 (kept in registers on the fast path)
template <int LOG2F, int LOG2P> __host__ __device__ constexpr int twiddle_count() {
    int n = 0, lns = stage_l<LOG2F, LOG2P>(0);
    while (lns < LOG2F) {
        int l = stage_l<LOG2F, LOG2P>(lns);
        n += ((1 << LOG2P) >> l) * ((1 << l) - 1);
        lns += l;
    }
    return n;
}
template <int LOG2F, int LOG2P> __host__ __device__ constexpr int exchange_count() {
    int n = 0, lns = stage_l<LOG2F, LOG2P>(0);
    while (lns < LOG2F) { lns += stage_l<LOG2F, LOG2P>(lns); ++n; }
    return n;
}

__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4) + 1; }

// one radix-R stage on the P registers of a thread; TW (P/R)*(R-1) twiddles or nullptr for stage 0
template <typename T, int P, int R, bool HASTW>
__device__ __forceinline__ void butterflies(cx<T> (&b)[P], const cx<T>* tw) {
    constexpr int V = P / R;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        cx<T> x[R];
#pragma unroll
        for (int t = 0; t < R; ++t) x[t] = b[v + t * V];
        if constexpr (HASTW) {
#pragma unroll
            for (int t = 1; t < R; ++t) x[t] = cmul(x[t], tw[v * (R - 1) + (t - 1)]);
        }
        dftR<T, R>(x);
#pragma unroll
        for (int t = 0; t < R; ++t) b[v + t * V] = x[t];
    }
}

// scatter the outputs of a radix-R stage (Ns = 1<<LNS before it) into the padded buffer
template <typename T, int P, int R, int NT, int LNS>
__device__ __forceinline__ void scatter(const cx<T> (&b)[P], cx<T>* sm, int tid) {
    constexpr int V = P / R;
    constexpr int NS = 1 << LNS;
    // pad_idx(base + t*NS) == pad_idx(base) + t*NS + (t*NS >> 4) whenever the low 4 bits of base cannot carry into
    // the padding term: one base address per butterfly, compile-time offsets per element (no per-element integer math)
    constexpr bool FAST = (NS >= 16) || ((NS * R) % 16 == 0);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int j = tid + v * NT;
        const int base = ((j >> LNS) << LNS) * R + (j & (NS - 1));
        if constexpr (FAST) {
            cx<T>* q = sm + pad_idx(base);
#pragma unroll
            for (int t = 0; t < R; ++t) q[t * NS + ((t * NS) >> 4)] = b[v + t * V];
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) sm[pad_idx(base + t * NS)] = b[v + t * V];
        }
    }
}

template <typename T, int P, int NT>
__device__ __forceinline__ void gather(cx<T> (&b)[P], const cx<T>* sm, int tid) {
    if constexpr (NT % 16 == 0) {
        const cx<T>* q = sm + pad_idx(tid);
#pragma unroll
        for (int m = 0; m < P; ++m) b[m] = q[m * (NT + NT / 16)];
    } else {
#pragma unroll
        for (int m = 0; m < P; ++m) b[m] = sm[pad_idx(tid + NT * m)];
    }
}

// fill the per-thread twiddle list for all stages >= 1 from the global table tw[k] = exp(-2 pi i k / F)
template <typename T, int LOG2F, int LOG2P, int LNS = stage_l<LOG2F, LOG2P>(0), int OFS = 0>
__device__ __forceinline__ void load_twiddles(cx<T>* dst, const cx<T>* __restrict__ table, int tid) {
    if constexpr (LNS < LOG2F) {
        constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
        constexpr int L = stage_l<LOG2F, LOG2P>(LNS), R = 1 << L, V = P / R;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int j = tid + v * NT;
            const int k = j & ((1 << LNS) - 1);
#pragma unroll
            for (int t = 1; t < R; ++t) {
                // t*k/(Ns*R) turns  ->  table index t*k*F/(Ns*R)
                dst[OFS + v * (R - 1) + (t - 1)] = table[(t * k) << (LOG2F - LNS - L)];
            }
        }
        load_twiddles<T, LOG2F, LOG2P, LNS + L, OFS + V * (R - 1)>(dst, table, tid);
    }
}

// Stages >= 1 of the transform (stage 0 is done by the caller right after its fused load*window).
//   TWREGS : twiddles come from the per-thread list `twl` (registers), else they are looked up in `table`
//            (global / L1) at every use.
//   DBUF   : two exchange buffers (one barrier per exchange) or one (two barriers).
// `sync` is a functor: team-wide barrier.
// Twiddles of the stages >= 1, when they are not register resident, come from a "linearised" table: one block per stage
// laid out [t-1][k], k = j mod Ns, holding exp(-2 pi i t k/(Ns R)), so that the lanes of a warp read consecutive words
// (the natural table exp(-2 pi i k/F) would be read with stride t*F/(Ns*R): scattered sectors in L1, 8..16-way bank
// conflicts in shared memory).  Block s starts at tw_lin_offset(lns_s); the total is F - R0 entries.
//   TWLDG = true : the table is in global memory, read through the read-only path;   false: a copy in shared memory.
template <int LOG2F, int LOG2P> __host__ __device__ constexpr int tw_lin_offset(int lns) {
    int ofs = 0, cur = stage_l<LOG2F, LOG2P>(0);
    while (cur < lns) {
        int l = stage_l<LOG2F, LOG2P>(cur);
        ofs += ((1 << l) - 1) << cur;
        cur += l;
    }
    return ofs;
}

template <typename T, int LOG2F, int LOG2P, int LNS = stage_l<LOG2F, LOG2P>(0)>
__device__ __forceinline__ void build_lin_twiddles(cx<T>* dst, const cx<T>* __restrict__ table, int thread, int nthreads) {
    if constexpr (LNS < LOG2F) {
        constexpr int L = stage_l<LOG2F, LOG2P>(LNS), R = 1 << L, NS = 1 << LNS;
        constexpr int OFS = tw_lin_offset<LOG2F, LOG2P>(LNS);
        for (int i = thread; i < (R - 1) * NS; i += nthreads) {
            const int t = (i >> LNS) + 1, k = i & (NS - 1);
            dst[OFS + i] = table[(t * k) << (LOG2F - LNS - L)];
        }
        build_lin_twiddles<T, LOG2F, LOG2P, LNS + L>(dst, table, thread, nthreads);
    }
}

template <bool LDG, int LOG2F, int LOG2P, int LNS, typename C>
__device__ __forceinline__ C ld_tw(const C* __restrict__ table, int t, int k) {
    const C* p = &table[tw_lin_offset<LOG2F, LOG2P>(LNS) + ((t - 1) << LNS) + k];
    if constexpr (LDG) return __ldg(p);
    else return *p;
}

// exp(-2 pi i n / 16), n = 0..15: the rotations that relate the twiddles of the butterflies of one thread in the LAST stage
// (virtual threads j = tid + v*NT:  W_F^(t*(tid + v*NT)) = W_F^(t*tid) * W_P^(t*v),  P <= 16 points per thread)
__device__ __forceinline__ constexpr double root16_re(int n) {
    constexpr double c[16] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173, 0.0, -0.38268343236508977173,
                              -0.70710678118654752440, -0.92387953251128675613, -1.0, -0.92387953251128675613, -0.70710678118654752440,
                              -0.38268343236508977173, 0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613};
    return c[n & 15];
}
__device__ __forceinline__ constexpr double root16_im(int n) { return -root16_re((n + 12) & 15); }   // -sin(x) = -cos(x - pi/2)

// twiddles of one tail stage for the V butterflies of a thread, from the table (not register resident)
template <typename T, int LOG2F, int LOG2P, int LNS, bool TWLDG>
__device__ __forceinline__ void fetch_stage_twiddles(cx<T>* tw, const cx<T>* __restrict__ table, int tid) {
    constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
    constexpr int L = stage_l<LOG2F, LOG2P>(LNS), R = 1 << L, V = P / R;
    constexpr bool LAST = (LNS + L == LOG2F);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int k = (tid + v * NT) & ((1 << LNS) - 1);
#pragma unroll
        for (int t = 1; t < R; ++t) {
            if constexpr (LAST && P <= 16) {
                // last stage: j = tid + v*NT never wraps, so butterfly v's factors are butterfly 0's rotated by the
                // compile-time constant W_P^(t*v): one table read per t instead of V (fewer shared-memory wavefronts)
                if (v == 0) tw[t - 1] = ld_tw<TWLDG, LOG2F, LOG2P, LNS>(table, t, k);
                else tw[v * (R - 1) + (t - 1)] = cmul(tw[t - 1], mkcx<T>((T)root16_re(t * v * (16 / P)), (T)root16_im(t * v * (16 / P))));
            } else {
                tw[v * (R - 1) + (t - 1)] = ld_tw<TWLDG, LOG2F, LOG2P, LNS>(table, t, k);
            }
        }
    }
}

template <typename T, int LOG2F, int LOG2P, bool TWREGS, bool DBUF, int LNS, int OFS, int XI, bool TWLDG = true, typename Sync>
__device__ __forceinline__ void fft_tail(cx<T> (&b)[1 << LOG2P], const cx<T>* twl, const cx<T>* __restrict__ table,
                                         cx<T>* buf0, cx<T>* buf1, int tid, Sync sync) {
    if constexpr (LNS < LOG2F) {
        constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
        constexpr int L = stage_l<LOG2F, LOG2P>(LNS), R = 1 << L, V = P / R;
        // every stage but the last has radix 2^LOG2P, so the stage that produced b[] had Ns = 2^(LNS-LOG2P)
        constexpr int LPREV = LOG2P;
        cx<T>* sm = (DBUF && (XI & 1)) ? buf1 : buf0;
        if constexpr (!DBUF) sync();
        scatter<T, P, (1 << LPREV), NT, LNS - LPREV>(b, sm, tid);
        sync();
        gather<T, P, NT>(b, sm, tid);
        if constexpr (TWREGS) {
            butterflies<T, P, R, true>(b, twl + OFS);
        } else {
            cx<T> tw[V * (R - 1)];
            fetch_stage_twiddles<T, LOG2F, LOG2P, LNS, TWLDG>(tw, table, tid);
            butterflies<T, P, R, true>(b, tw);
        }
        fft_tail<T, LOG2F, LOG2P, TWREGS, DBUF, LNS + L, OFS + V * (R - 1), XI + 1, TWLDG>(b, twl, table, buf0, buf1, tid, sync);
    }
}

// Same as fft_tail, but the barrier of the FIRST exchange is `syncFirst` (the staged kernel re-arms its TMA prefetch
// there: once every thread is past that barrier the stage buffer has been fully consumed).
template <typename T, int LOG2F, int LOG2P, bool TWREGS, bool DBUF, int LNS, bool TWLDG = true, typename Sync, typename SyncFirst>
__device__ __forceinline__ void fft_tail_first(cx<T> (&b)[1 << LOG2P], const cx<T>* twl, const cx<T>* __restrict__ table,
                                               cx<T>* buf0, cx<T>* buf1, int tid, Sync sync, SyncFirst syncFirst) {
    static_assert(LNS < LOG2F, "needs at least one exchange");
    constexpr int P = 1 << LOG2P, F = 1 << LOG2F, NT = F / P;
    constexpr int L = stage_l<LOG2F, LOG2P>(LNS), R = 1 << L, V = P / R;
    if constexpr (!DBUF) sync();
    scatter<T, P, (1 << LOG2P), NT, LNS - LOG2P>(b, buf0, tid);
    syncFirst();
    gather<T, P, NT>(b, buf0, tid);
    if constexpr (TWREGS) {
        butterflies<T, P, R, true>(b, twl);
    } else {
        cx<T> tw[V * (R - 1)];
        fetch_stage_twiddles<T, LOG2F, LOG2P, LNS, TWLDG>(tw, table, tid);
        butterflies<T, P, R, true>(b, tw);
    }
    fft_tail<T, LOG2F, LOG2P, TWREGS, DBUF, LNS + L, V * (R - 1), 1, TWLDG>(b, twl, table, buf0, buf1, tid, sync);
}

}  // namespace kspec
