// mixedradix.cu — float64 two-pass mixed-radix engine for frame lengths F = 2^a 3^b 5^c 7^d that are not powers of two
// (2 400 000 = 2^8.3.5^5 of BASELINE cfg-5, 48 000, 5 000, ...).  Row N4 of SURVEY 8f: numpy itself runs a mixed-radix
// transform at kspecanal.py:391 for these sizes; compared with the Bluestein engine (bigfft.cu) this is one F-point
// transform instead of two M >= 2F point ones (x7 less data at 2.4e6) and free of the chirp's cancellation noise.
//
//   F = N1*N2 (both <= MR_MAX_LINE), element n = n1*N2 + n2, bin k = k1 + N1*k2
//   pass 1  columns: for TC adjacent n2, N1-point transforms over n1 (fused ingest + window), times W_F^(n2*k1) -> Z[k1][n2]
//   pass 2  rows   : for TC adjacent k1, N2-point transforms over n2; |X| (or |X|^2) cumulated over the frames of a scan
//                    in the registers that own the bins (data_cumu K:124-147); stored once per scan in natural bin order.
//
// A line transform is a Stockham autosort over two shared-memory buffers, one stage per radix (odd radices first: their
// strides are conflict-free), twiddles from an exact L-entry table exp(-2 pi i t/L).
#include "bigfft_kernels.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace kspec {

constexpr int MR_MAX_STAGES = 16;
constexpr int MR_THREADS = 512;                   // default CTA (64 registers, two CTAs per SM); KSPEC_MR_THREADS=256 for tuning
constexpr int MR_MAX_LINE = 4096;                 // longest line: buffer pair 2 x 64 KB + at most 32 KB of twiddles always fits one CTA
constexpr int MR_SMEM_BIG = 220 * 1024;           // one CTA per SM: long lines
constexpr int MR_SMEM_STD = 112 * 1024;           // default tile: two CTAs per SM, so one CTA's loads overlap the other's butterflies
constexpr int MR_TW_LO_BITS = 10;                 // inter-pass twiddle W_F^t = hi[t >> 10] * lo[t & 1023]: both tables cache-resident
constexpr int mr_maxa(int smem, int nt) { return (smem / 32 + nt - 1) / nt; }       // accumulators per thread for a tile

struct MrSched {
    int L;                        // line length
    int TC, lTC;                  // lines per tile (power of two) and its log2
    int nStages;
    int tabN;                     // twiddle entries kept in shared memory: exp(-2 pi i t/L), t < L / (smallest twiddled radix)
    uint32_t mL;                  // magic reciprocal of L
    int radix[MR_MAX_STAGES];
    uint32_t mNs[MR_MAX_STAGES];  // magic reciprocals of the stage's sub-transform length Ns and butterfly count L/R
    uint32_t mNb[MR_MAX_STAGES];
};

namespace {

// n / d for n * d < 2^32 with m = ceil(2^32 / d), d >= 2
__device__ __forceinline__ int mr_div(int n, uint32_t m) { return (int)__umulhi((uint32_t)n, m); }

__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cd mul_mi(cd a) { return make_double2(a.y, -a.x); }      // a * (-i)

template <int R> __device__ __forceinline__ void mr_dft(cd* v);

template <> __device__ __forceinline__ void mr_dft<2>(cd* v) {
    const cd a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
}
template <> __device__ __forceinline__ void mr_dft<3>(cd* v) {
    const cd t = cadd(v[1], v[2]);
    const cd m = make_double2(v[0].x - 0.5 * t.x, v[0].y - 0.5 * t.y);
    const cd s = mul_mi(cscale(csub(v[1], v[2]), 0.86602540378443864676));      // -i sin(pi/3) (v1 - v2)
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, s);
    v[2] = csub(m, s);
}
template <> __device__ __forceinline__ void mr_dft<4>(cd* v) {
    const cd a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}
template <> __device__ __forceinline__ void mr_dft<5>(cd* v) {
    constexpr double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;      // cos(2pi/5), cos(4pi/5)
    constexpr double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;       // sin(2pi/5), sin(4pi/5)
    const cd a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const cd p1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const cd p2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const cd q1 = mul_mi(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const cd q2 = mul_mi(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(p1, q1);
    v[4] = csub(p1, q1);
    v[2] = cadd(p2, q2);
    v[3] = csub(p2, q2);
}
// radix 7 (rare) is a direct 7-point transform written straight to its destination, one output per trip of a rolled loop so
// that it does not set the register budget of the kernel; the seventh roots come from the line's global table (7 | L)
__device__ __forceinline__ cd mr_dft7_out(const cd* v, int a, const cd* __restrict__ gtab, int L) {
    cd s = v[0];
#pragma unroll
    for (int b = 1; b < 7; ++b) s = cadd(s, cmul(v[b], __ldg(&gtab[((a * b) % 7) * (L / 7)])));
    return s;
}

// shared-memory index of element e of line `line`: columns keep the lines of a tile interleaved (they arrive that way
// from global memory), rows keep each line contiguous
template <bool LINE_FAST> __device__ __forceinline__ int mr_idx(int e, int line, int lTC, int L) {
    return LINE_FAST ? (e << lTC) + line : line * L + e;
}

// one Stockham stage: butterfly j of a line reads elements j + m*L/R, multiplies by W_{Ns R}^(k m) (k = j mod Ns; one table
// entry, the higher powers by multiplication: float64 pipe time is free here, table loads are not), transforms, and writes
// (j - k) R + k + m Ns
template <int R, bool LINE_FAST, int NT>
__device__ __forceinline__ void mr_stage(const MrSched& sc, const cd* __restrict__ src, cd* __restrict__ dst, const cd* __restrict__ stab,
                                         const cd* __restrict__ gtab, int Ns, uint32_t mNs, uint32_t mNb) {
    const int L = sc.L, lTC = sc.lTC, nb = L / R, total = nb << lTC;
    const int twStep = L / (Ns * R);
#pragma unroll 1
    for (int b = threadIdx.x; b < total; b += NT) {
        int line, j;
        if (LINE_FAST) { line = b & ((1 << lTC) - 1); j = b >> lTC; } else { line = nb > 1 ? mr_div(b, mNb) : b; j = b - line * nb; }
        const int k = Ns > 1 ? j - mr_div(j, mNs) * Ns : 0;
        cd v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = src[mr_idx<LINE_FAST>(j + m * nb, line, lTC, L)];
        if (Ns > 1) {
            const cd w1 = stab[k * twStep];
            cd w = w1;
            v[1] = cmul(v[1], w1);
#pragma unroll
            for (int m = 2; m < R; ++m) {
                w = cmul(w, w1);
                v[m] = cmul(v[m], w);
            }
        }
        const int j0 = (j - k) * R + k;
        if constexpr (R == 7) {
#pragma unroll 1
            for (int m = 0; m < R; ++m) dst[mr_idx<LINE_FAST>(j0 + m * Ns, line, lTC, L)] = mr_dft7_out(v, m, gtab, L);
        } else {
            mr_dft<R>(v);
#pragma unroll
            for (int m = 0; m < R; ++m) dst[mr_idx<LINE_FAST>(j0 + m * Ns, line, lTC, L)] = v[m];
        }
    }
}

// all stages of the tile held in `a`; returns the buffer that holds the (natural order) result
template <bool LINE_FAST, int NT>
__device__ __forceinline__ cd* mr_transform(const MrSched& sc, cd* a, cd* b, const cd* __restrict__ stab, const cd* __restrict__ gtab) {
    int Ns = 1;
    for (int s = 0; s < sc.nStages; ++s) {
        const int R = sc.radix[s];
        const uint32_t mNs = sc.mNs[s], mNb = sc.mNb[s];
        switch (R) {
            case 2: mr_stage<2, LINE_FAST, NT>(sc, a, b, stab, gtab, Ns, mNs, mNb); break;
            case 3: mr_stage<3, LINE_FAST, NT>(sc, a, b, stab, gtab, Ns, mNs, mNb); break;
            case 4: mr_stage<4, LINE_FAST, NT>(sc, a, b, stab, gtab, Ns, mNs, mNb); break;
            case 5: mr_stage<5, LINE_FAST, NT>(sc, a, b, stab, gtab, Ns, mNs, mNb); break;
            default: mr_stage<7, LINE_FAST, NT>(sc, a, b, stab, gtab, Ns, mNs, mNb); break;
        }
        __syncthreads();
        cd* t = a; a = b; b = t;
        Ns *= R;
    }
    return a;
}

struct MrColsParams {
    MrSched sc;                   // L = N1
    const void* samples; int64_t scanStride; const int64_t* offs; int nFrames;
    const double* win; const cd* tabL; const cd* tabHi; const cd* tabLo; cd* Z;
    int64_t F; int N1, N2; int64_t nfs;
    double u8off, u8scale;
};

template <int INFMT, int NT>
__global__ void __launch_bounds__(NT, 2) mr_cols_kernel(const MrColsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TC = p.sc.TC, lTC = p.sc.lTC, N1 = p.N1, N2 = p.N2;
    cd* bufA = reinterpret_cast<cd*>(smem_raw);
    cd* bufB = bufA + (size_t)TC * N1;
    cd* stab = bufB + (size_t)TC * N1;
    for (int i = threadIdx.x; i < p.sc.tabN; i += NT) stab[i] = p.tabL[i];
    const int tpf = (N2 + TC - 1) / TC;                     // tiles per frame
    const int64_t tiles = p.nfs * tpf;
    const int items = TC * N1;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t fs = tile / tpf;
        const int c0 = (int)(tile - fs * tpf) * TC;
        const int64_t s = fs / p.nFrames;
        const int64_t base = s * p.scanStride + __ldg(&p.offs[fs - s * p.nFrames]);
#pragma unroll 4
        for (int i = threadIdx.x; i < items; i += NT) {
            // branch-free, so that the unrolled iterations' loads issue back to back (a ragged last tile re-reads column 0)
            const int line = i & (TC - 1), e = i >> lTC, n2 = c0 + line;
            const bool in = n2 < N2;
            const int64_t n = (int64_t)e * N2 + (in ? n2 : 0);
            const cd v = Ingest<double, INFMT>::load(p.samples, base + n, __ldg(&p.win[n]), p.u8off, p.u8scale);
            bufA[i] = in ? v : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const cd* res = mr_transform<true, NT>(p.sc, bufA, bufB, stab, p.tabL);
#pragma unroll 4
        for (int i = threadIdx.x; i < items; i += NT) {
            const int line = i & (TC - 1), k1 = i >> lTC, n2 = c0 + line;
            const bool in = n2 < N2;
            const int t = (in ? n2 : 0) * k1;                // < F <= MR_MAX_LINE^2 < 2^31
            const cd w = cmul(__ldg(&p.tabHi[t >> MR_TW_LO_BITS]), __ldg(&p.tabLo[t & ((1 << MR_TW_LO_BITS) - 1)]));
            if (in) p.Z[fs * p.F + (int64_t)k1 * N2 + n2] = cmul(res[i], w);
        }
        __syncthreads();
    }
}

struct MrRowsParams {
    MrSched sc;                   // L = N2
    const cd* Z; const cd* tabL; double* acc;
    int64_t F; int N1, N2; int64_t nScans; int nFrames; int cumuMode;
};

template <int MR_MAXA, int NT>
__global__ void __launch_bounds__(NT, (MR_MAXA <= mr_maxa(MR_SMEM_STD, NT) ? 2 : 1)) mr_rows_acc_kernel(const MrRowsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TC = p.sc.TC, N1 = p.N1, N2 = p.N2;
    cd* bufA = reinterpret_cast<cd*>(smem_raw);
    cd* bufB = bufA + (size_t)TC * N2;
    cd* stab = bufB + (size_t)TC * N2;
    for (int i = threadIdx.x; i < p.sc.tabN; i += NT) stab[i] = p.tabL[i];
    const int tps = (N1 + TC - 1) / TC;                     // tiles per scan
    const int64_t tiles = p.nScans * tps;
    const int items = TC * N2;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s = tile / tps;
        const int r0 = (int)(tile - s * tps) * TC;
        double acc[MR_MAXA];
        for (int f = 0; f < p.nFrames; ++f) {
            const cd* __restrict__ zf = p.Z + (s * p.nFrames + f) * p.F;
#pragma unroll 4
            for (int i = threadIdx.x; i < items; i += NT) {
                const int line = mr_div(i, p.sc.mL), e = i - line * N2, k1 = r0 + line;
                const bool in = k1 < N1;
                const cd v = __ldg(&zf[(int64_t)(in ? k1 : 0) * N2 + e]);
                bufA[i] = in ? v : make_double2(0.0, 0.0);
            }
            __syncthreads();
            const cd* res = mr_transform<false, NT>(p.sc, bufA, bufB, stab, p.tabL);
#pragma unroll
            for (int m = 0; m < MR_MAXA; ++m) {
                const int i = threadIdx.x + NT * m;
                if (i < items) {
                    const cd x = res[i];
                    double mag = x.x * x.x + x.y * x.y;
                    if (p.cumuMode != KSPEC_CUMU_PSD) mag = sqrt(mag);
                    if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) acc[m] = mag;
                    else if (p.cumuMode == KSPEC_CUMU_AVG) acc[m] = (acc[m] + mag) / 2;
                    else if (p.cumuMode == KSPEC_CUMU_MAX) acc[m] = fmax(acc[m], mag);
                    else if (p.cumuMode == KSPEC_CUMU_MIN) acc[m] = fmin(acc[m], mag);
                    else acc[m] += mag;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int m = 0; m < MR_MAXA; ++m) {
            const int i = threadIdx.x + NT * m;
            if (i < items) {
                const int line = mr_div(i, p.sc.mL), k2 = i - line * N2, k1 = r0 + line;
                if (k1 < N1) p.acc[s * p.F + k1 + (int64_t)N1 * k2] = acc[m];
            }
        }
    }
}

// t[k] = exp(-2 pi i k mult / n)
__global__ void mr_table_kernel(cd* t, int64_t count, int64_t mult, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)((k * mult) % n) / (double)n, &s, &c);
        t[k] = make_double2(c, s);
    }
}

uint32_t mr_magic(int d) { return (uint32_t)((((uint64_t)1 << 32) + (uint64_t)d - 1) / (uint64_t)d); }

// radix schedule of a line: odd radices first (conflict-free strides), then 4s, then a single 2
bool mr_schedule(int L, MrSched* sc) {
    sc->L = L; sc->nStages = 0; sc->TC = 1; sc->lTC = 0;
    int rest = L;
    for (int r : {7, 5, 3}) while (rest % r == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = r; rest /= r; }
    while (rest % 4 == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = 4; rest /= 4; }
    if (rest % 2 == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = 2; rest /= 2; }
    if (rest != 1) return false;
    sc->mL = mr_magic(L);
    sc->tabN = 0;
    int Ns = 1;
    for (int s = 0; s < sc->nStages; ++s) {
        sc->mNs[s] = mr_magic(Ns);
        sc->mNb[s] = mr_magic(L / sc->radix[s]);
        if (s > 0 && L / sc->radix[s] > sc->tabN) sc->tabN = L / sc->radix[s];
        Ns *= sc->radix[s];
    }
    return true;
}

// lines per tile: the largest power of two whose buffer pair fits next to the twiddle table in `smem` bytes
void mr_tile_lines(MrSched* sc, int other, int smem) {
    int tc = (smem - sc->tabN * 16) / (32 * sc->L);
    if (tc > 32) tc = 32;
    if (tc > other) tc = other;
    int l = 0;
    while ((2 << l) <= tc) ++l;
    sc->lTC = l;
    sc->TC = 1 << l;
}
size_t mr_smem_bytes(const MrSched& sc) { return (size_t)2 * sc.TC * sc.L * 16 + (size_t)sc.tabN * 16; }

}  // namespace

// F = N1*N2 with both factors 7-smooth, as square as possible and within the line limit; false when F does not qualify
bool mixedradix_split(int64_t F, int* n1, int* n2) {
    if (F < 2) return false;
    int64_t rest = F;
    for (int r : {2, 3, 5, 7}) while (rest % r == 0) rest /= r;
    if (rest != 1) return false;
    int64_t best = 0;
    for (int64_t d = 1; d * d <= F; ++d)
        if (F % d == 0 && F / d <= MR_MAX_LINE) { best = d; }
    if (best < 2) return false;
    *n1 = (int)best;
    *n2 = (int)(F / best);
    return true;
}

struct MixedRadix {
    int inFmt = 0, smCount = 0;
    int threads = 0;            // CTA size of both passes (KSPEC_MR_THREADS=256 at creation: tuning knob)
    int64_t F = 0;
    int N1 = 0, N2 = 0;
    MrSched sc1{}, sc2{};
    double u8off = 0, u8scale = 0;
    cudaStream_t st = nullptr;
    double* dWin = nullptr;
    cd *dTab1 = nullptr, *dTab2 = nullptr, *dTabHi = nullptr, *dTabLo = nullptr, *dZ = nullptr;
    int64_t* dOffs = nullptr;
    int nOffs = 0;
    size_t zCap = 0;
};

void mixedradix_destroy(MixedRadix* b) {
    if (!b) return;
    for (void* p : {(void*)b->dWin, (void*)b->dTab1, (void*)b->dTab2, (void*)b->dTabHi, (void*)b->dTabLo, (void*)b->dZ, (void*)b->dOffs})
        if (p) cudaFree(p);
    delete b;
}

#define MCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(err, errLen, "%s: %s", #call, cudaGetErrorString(e_));                        \
            mixedradix_destroy(b);                                                                 \
            return nullptr;                                                                        \
        }                                                                                          \
    } while (0)

MixedRadix* mixedradix_create(int prec, int inFmt, int64_t F, const double* window, double u8off, double u8scale, cudaStream_t st,
                              char* err, size_t errLen) {
    if (prec != KSPEC_PREC_F64) {
        snprintf(err, errLen, "fftSize %lld runs on the multi-pass engines, which compute in float64: use precision auto or f64", (long long)F);
        return nullptr;
    }
    MixedRadix* b = new MixedRadix();
    b->inFmt = inFmt; b->F = F; b->u8off = u8off; b->u8scale = u8scale; b->st = st;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&b->smCount, cudaDevAttrMultiProcessorCount, dev);
    if (!mixedradix_split(F, &b->N1, &b->N2) || !mr_schedule(b->N1, &b->sc1) || !mr_schedule(b->N2, &b->sc2)) {
        snprintf(err, errLen, "fftSize %lld has no mixed-radix split", (long long)F);
        mixedradix_destroy(b);
        return nullptr;
    }
    MCK(cudaMalloc(&b->dWin, (size_t)F * 8));
    MCK(cudaMemcpyAsync(b->dWin, window, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    MCK(cudaMalloc(&b->dTab1, (size_t)b->N1 * 16));
    MCK(cudaMalloc(&b->dTab2, (size_t)b->N2 * 16));
    b->threads = MR_THREADS;
    if (const char* e = getenv("KSPEC_MR_THREADS")) { if (atoi(e) == 256) b->threads = 256; }
    int smem = MR_SMEM_STD;
    if (const char* e = getenv("KSPEC_MR_SMEM")) { const int v = atoi(e); if (v >= 16 * 1024 && v <= MR_SMEM_BIG) smem = v; }
    for (MrSched* sc : {&b->sc1, &b->sc2}) {
        mr_tile_lines(sc, sc == &b->sc1 ? b->N2 : b->N1, smem);
        if (mr_smem_bytes(*sc) > (size_t)smem) mr_tile_lines(sc, 1, MR_SMEM_BIG);       // a long line: one per CTA, one CTA per SM
    }
    const int64_t nHi = (F >> MR_TW_LO_BITS) + 1, nLo = (int64_t)1 << MR_TW_LO_BITS;
    MCK(cudaMalloc(&b->dTabHi, (size_t)nHi * 16));
    MCK(cudaMalloc(&b->dTabLo, (size_t)nLo * 16));
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTab1, b->N1, 1, b->N1);
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTab2, b->N2, 1, b->N2);
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTabHi, nHi, nLo, F);
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTabLo, nLo, 1, F);
    MCK(cudaGetLastError());
    MCK(cudaStreamSynchronize(st));
    return b;
}

void mixedradix_info(const MixedRadix* b, int* n1, int* n2) { *n1 = b->N1; *n2 = b->N2; }

int mixedradix_run(MixedRadix* b, const void* samples, int64_t scanStride, int64_t nScans, const int64_t* frameOffs, int nFrames,
                   int cumuMode, void* acc, int64_t* launches) {
    cudaStream_t st = b->st;
    if (b->nOffs != nFrames) {
        if (b->dOffs) cudaFree(b->dOffs);
        b->dOffs = nullptr;
        if (cudaMalloc(&b->dOffs, (size_t)nFrames * 8) != cudaSuccess) { set_error("frame table allocation failed"); return KSPEC_ERR_NOMEM; }
        b->nOffs = nFrames;
    }
    cudaMemcpyAsync(b->dOffs, frameOffs, (size_t)nFrames * 8, cudaMemcpyHostToDevice, st);
    // scans are processed in chunks whose work vector (one F-point slab per frame) stays within ~1 GiB
    const size_t slab = (size_t)b->F * 16;
    int64_t chunk = (int64_t)(((size_t)1 << 30) / (slab * (size_t)nFrames));
    if (chunk < 1) chunk = 1;
    if (chunk > nScans) chunk = nScans;
    const size_t need = slab * (size_t)nFrames * (size_t)chunk;
    if (b->zCap < need) {
        if (b->dZ) cudaFree(b->dZ);
        b->dZ = nullptr; b->zCap = 0;
        if (cudaMalloc(&b->dZ, need) != cudaSuccess) {
            cudaGetLastError();
            set_error("mixed-radix work buffer (%zu bytes) does not fit in device memory", need);
            return KSPEC_ERR_NOMEM;
        }
        b->zCap = need;
    }
    const size_t eb = b->inFmt == KSPEC_IN_U8_IQ ? 2 : (b->inFmt == KSPEC_IN_C64 ? 8 : 16);
    const size_t smem1 = mr_smem_bytes(b->sc1), smem2 = mr_smem_bytes(b->sc2);
    const int nt = b->threads;
    void (*kr)(const MrRowsParams);
    void (*kc)(const MrColsParams);
    const bool stdTile = smem2 <= (size_t)MR_SMEM_STD;
    if (nt == 512) {
        kr = stdTile ? mr_rows_acc_kernel<mr_maxa(MR_SMEM_STD, 512), 512> : mr_rows_acc_kernel<mr_maxa(MR_SMEM_BIG, 512), 512>;
        kc = b->inFmt == KSPEC_IN_U8_IQ ? mr_cols_kernel<KSPEC_IN_U8_IQ, 512> : (b->inFmt == KSPEC_IN_C64 ? mr_cols_kernel<KSPEC_IN_C64, 512> : mr_cols_kernel<KSPEC_IN_C128, 512>);
    } else {
        kr = stdTile ? mr_rows_acc_kernel<mr_maxa(MR_SMEM_STD, 256), 256> : mr_rows_acc_kernel<mr_maxa(MR_SMEM_BIG, 256), 256>;
        kc = b->inFmt == KSPEC_IN_U8_IQ ? mr_cols_kernel<KSPEC_IN_U8_IQ, 256> : (b->inFmt == KSPEC_IN_C64 ? mr_cols_kernel<KSPEC_IN_C64, 256> : mr_cols_kernel<KSPEC_IN_C128, 256>);
    }
    if (cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess ||
        cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) {
        set_error("mixed-radix shared memory request failed: %s", cudaGetErrorString(cudaGetLastError()));
        return KSPEC_ERR_CUDA;
    }
    int per1 = 1, per2 = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per1, kc, nt, smem1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, kr, nt, smem2);
    if (per1 < 1) per1 = 1;
    if (per2 < 1) per2 = 1;
    double* dAcc = reinterpret_cast<double*>(acc);
    for (int64_t s0 = 0; s0 < nScans; s0 += chunk) {
        const int64_t ns = (nScans - s0 < chunk) ? nScans - s0 : chunk;
        const int64_t nfs = ns * nFrames;
        const void* smp = reinterpret_cast<const unsigned char*>(samples) + (size_t)s0 * scanStride * eb;
        MrColsParams pc{b->sc1, smp, scanStride, b->dOffs, nFrames, b->dWin, b->dTab1, b->dTabHi, b->dTabLo, b->dZ, b->F, b->N1, b->N2, nfs, b->u8off, b->u8scale};
        const int64_t t1 = nfs * ((b->N2 + b->sc1.TC - 1) / b->sc1.TC);
        const int64_t cap1 = (int64_t)b->smCount * per1;
        kc<<<(int)(t1 < cap1 ? t1 : cap1), nt, smem1, st>>>(pc);
        MrRowsParams pr{b->sc2, b->dZ, b->dTab2, dAcc + s0 * b->F, b->F, b->N1, b->N2, ns, nFrames, cumuMode};
        const int64_t t2 = ns * ((b->N1 + b->sc2.TC - 1) / b->sc2.TC);
        const int64_t cap2 = (int64_t)b->smCount * per2;
        kr<<<(int)(t2 < cap2 ? t2 : cap2), nt, smem2, st>>>(pr);
        *launches += 2;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("mixed-radix FFT launch failed: %s", cudaGetErrorString(e)); return KSPEC_ERR_CUDA; }
    return KSPEC_OK;
}

}  // namespace kspec
