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
constexpr int MR_THREADS = 256;
constexpr int MR_MAX_LINE = 6144;                 // longest line: one buffer pair of 2 x 96 KB
constexpr int MR_TILE_ELEMS = 6144;               // TC * L <= this (elements per buffer)
constexpr int MR_TILE_TARGET = 3200;              // default tile: two CTAs (2 x 100 KB) per SM so loads overlap butterflies
constexpr int MR_MAXA_BIG = (MR_TILE_ELEMS + MR_THREADS - 1) / MR_THREADS;      // accumulators per thread, largest tile
constexpr int MR_MAXA_STD = (MR_TILE_TARGET + MR_THREADS - 1) / MR_THREADS;     // ... default tile (two CTAs per SM)

struct MrSched {
    int L;                        // line length
    int TC;                       // lines per tile
    int nStages;
    int radix[MR_MAX_STAGES];
};

namespace {

__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cd mul_mi(cd a) { return make_double2(a.y, -a.x); }      // a * (-i)

template <int R> __device__ __forceinline__ void mr_dft(cd* v, const cd* __restrict__ tab, int L);

template <> __device__ __forceinline__ void mr_dft<2>(cd* v, const cd*, int) {
    const cd a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
}
template <> __device__ __forceinline__ void mr_dft<3>(cd* v, const cd*, int) {
    const cd t = cadd(v[1], v[2]);
    const cd m = make_double2(v[0].x - 0.5 * t.x, v[0].y - 0.5 * t.y);
    const cd s = mul_mi(cscale(csub(v[1], v[2]), 0.86602540378443864676));      // -i sin(pi/3) (v1 - v2)
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, s);
    v[2] = csub(m, s);
}
template <> __device__ __forceinline__ void mr_dft<4>(cd* v, const cd*, int) {
    const cd a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}
template <> __device__ __forceinline__ void mr_dft<5>(cd* v, const cd*, int) {
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
template <> __device__ __forceinline__ void mr_dft<7>(cd* v, const cd* __restrict__ tab, int L) {
    // direct 7-point transform; the seventh roots come from the line's own twiddle table (7 | L)
    cd w[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) w[t] = __ldg(&tab[t * (L / 7)]);
    cd o[7];
#pragma unroll
    for (int a = 0; a < 7; ++a) {
        cd s = v[0];
#pragma unroll
        for (int b = 1; b < 7; ++b) s = cadd(s, cmul(v[b], w[(a * b) % 7]));
        o[a] = s;
    }
#pragma unroll
    for (int a = 0; a < 7; ++a) v[a] = o[a];
}

// shared-memory index of element e of line `line`: columns keep the lines of a tile interleaved (they arrive that way
// from global memory), rows keep each line contiguous
template <bool LINE_FAST> __device__ __forceinline__ int mr_idx(int e, int line, int TC, int L) {
    return LINE_FAST ? e * TC + line : line * L + e;
}

template <int R, bool LINE_FAST>
__device__ __forceinline__ void mr_stage(const MrSched& sc, const cd* __restrict__ src, cd* __restrict__ dst, const cd* __restrict__ tab, int Ns) {
    const int L = sc.L, TC = sc.TC, nb = L / R, total = nb * TC;
    const int twStep = L / (Ns * R);
    for (int b = threadIdx.x; b < total; b += MR_THREADS) {
        int line, j;
        if (LINE_FAST) { line = b % TC; j = b / TC; } else { j = b % nb; line = b / nb; }
        const int k = j % Ns;
        cd v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = src[mr_idx<LINE_FAST>(j + m * nb, line, TC, L)];
        if (Ns > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) v[m] = cmul(v[m], __ldg(&tab[k * m * twStep]));
        }
        mr_dft<R>(v, tab, L);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; ++m) dst[mr_idx<LINE_FAST>(j0 + m * Ns, line, TC, L)] = v[m];
    }
}

// all stages of the tile held in `a`; returns the buffer that holds the (natural order) result
template <bool LINE_FAST>
__device__ __forceinline__ cd* mr_transform(const MrSched& sc, cd* a, cd* b, const cd* __restrict__ tab) {
    int Ns = 1;
    for (int s = 0; s < sc.nStages; ++s) {
        const int R = sc.radix[s];
        switch (R) {
            case 2: mr_stage<2, LINE_FAST>(sc, a, b, tab, Ns); break;
            case 3: mr_stage<3, LINE_FAST>(sc, a, b, tab, Ns); break;
            case 4: mr_stage<4, LINE_FAST>(sc, a, b, tab, Ns); break;
            case 5: mr_stage<5, LINE_FAST>(sc, a, b, tab, Ns); break;
            default: mr_stage<7, LINE_FAST>(sc, a, b, tab, Ns); break;
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
    const double* win; const cd* tabL; const cd* tabF; cd* Z;
    int64_t F; int N1, N2; int64_t nfs;
    double u8off, u8scale;
};

template <int INFMT>
__global__ void __launch_bounds__(MR_THREADS, 2) mr_cols_kernel(const MrColsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TC = p.sc.TC, N1 = p.N1, N2 = p.N2;
    cd* bufA = reinterpret_cast<cd*>(smem_raw);
    cd* bufB = bufA + (size_t)TC * N1;
    const int tpf = (N2 + TC - 1) / TC;                     // tiles per frame
    const int64_t tiles = p.nfs * tpf;
    const int items = TC * N1;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t fs = tile / tpf;
        const int c0 = (int)(tile - fs * tpf) * TC;
        const int64_t s = fs / p.nFrames;
        const int64_t base = s * p.scanStride + __ldg(&p.offs[fs - s * p.nFrames]);
        for (int i = threadIdx.x; i < items; i += MR_THREADS) {
            const int line = i % TC, e = i / TC, n2 = c0 + line;
            cd v = make_double2(0.0, 0.0);
            if (n2 < N2) {
                const int64_t n = (int64_t)e * N2 + n2;
                v = Ingest<double, INFMT>::load(p.samples, base + n, __ldg(&p.win[n]), p.u8off, p.u8scale);
            }
            bufA[i] = v;
        }
        __syncthreads();
        const cd* res = mr_transform<true>(p.sc, bufA, bufB, p.tabL);
        for (int i = threadIdx.x; i < items; i += MR_THREADS) {
            const int line = i % TC, k1 = i / TC, n2 = c0 + line;
            if (n2 < N2) p.Z[fs * p.F + (int64_t)k1 * N2 + n2] = cmul(res[i], __ldg(&p.tabF[(int64_t)n2 * k1]));
        }
        __syncthreads();
    }
}

struct MrRowsParams {
    MrSched sc;                   // L = N2
    const cd* Z; const cd* tabL; double* acc;
    int64_t F; int N1, N2; int64_t nScans; int nFrames; int cumuMode;
};

template <int MR_MAXA>
__global__ void __launch_bounds__(MR_THREADS, (MR_MAXA <= MR_MAXA_STD ? 2 : 1)) mr_rows_acc_kernel(const MrRowsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TC = p.sc.TC, N1 = p.N1, N2 = p.N2;
    cd* bufA = reinterpret_cast<cd*>(smem_raw);
    cd* bufB = bufA + (size_t)TC * N2;
    const int tps = (N1 + TC - 1) / TC;                     // tiles per scan
    const int64_t tiles = p.nScans * tps;
    const int items = TC * N2;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s = tile / tps;
        const int r0 = (int)(tile - s * tps) * TC;
        double acc[MR_MAXA];
        for (int f = 0; f < p.nFrames; ++f) {
            const cd* __restrict__ zf = p.Z + (s * p.nFrames + f) * p.F;
            for (int i = threadIdx.x; i < items; i += MR_THREADS) {
                const int line = i / N2, e = i - line * N2, k1 = r0 + line;
                bufA[i] = (k1 < N1) ? zf[(int64_t)k1 * N2 + e] : make_double2(0.0, 0.0);
            }
            __syncthreads();
            const cd* res = mr_transform<false>(p.sc, bufA, bufB, p.tabL);
#pragma unroll
            for (int m = 0; m < MR_MAXA; ++m) {
                const int i = threadIdx.x + MR_THREADS * m;
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
            const int i = threadIdx.x + MR_THREADS * m;
            if (i < items) {
                const int line = i / N2, k2 = i - line * N2, k1 = r0 + line;
                if (k1 < N1) p.acc[s * p.F + k1 + (int64_t)N1 * k2] = acc[m];
            }
        }
    }
}

__global__ void mr_table_kernel(cd* t, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)n, &s, &c);
        t[k] = make_double2(c, s);
    }
}

// radix schedule of a line: odd radices first, then 4s, then a single 2
bool mr_schedule(int L, int TC, MrSched* sc) {
    sc->L = L; sc->TC = TC; sc->nStages = 0;
    int rest = L;
    for (int r : {7, 5, 3}) while (rest % r == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = r; rest /= r; }
    while (rest % 4 == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = 4; rest /= 4; }
    if (rest % 2 == 0) { if (sc->nStages == MR_MAX_STAGES) return false; sc->radix[sc->nStages++] = 2; rest /= 2; }
    return rest == 1;
}

int mr_tile_lines(int L, int other) {
    int budget = MR_TILE_TARGET;
    if (const char* e = getenv("KSPEC_MR_TILE_ELEMS")) { const int v = atoi(e); if (v >= 1 && v <= MR_TILE_ELEMS) budget = v; }
    int tc = budget / L;
    if (tc > 32) tc = 32;
    if (tc > other) tc = other;
    while (tc > 1 && (tc & (tc - 1))) --tc;           // power of two: whole sectors
    return tc < 1 ? 1 : tc;
}

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
    int64_t F = 0;
    int N1 = 0, N2 = 0;
    MrSched sc1{}, sc2{};
    double u8off = 0, u8scale = 0;
    cudaStream_t st = nullptr;
    double* dWin = nullptr;
    cd *dTab1 = nullptr, *dTab2 = nullptr, *dTabF = nullptr, *dZ = nullptr;
    int64_t* dOffs = nullptr;
    int nOffs = 0;
    size_t zCap = 0;
};

void mixedradix_destroy(MixedRadix* b) {
    if (!b) return;
    for (void* p : {(void*)b->dWin, (void*)b->dTab1, (void*)b->dTab2, (void*)b->dTabF, (void*)b->dZ, (void*)b->dOffs})
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
    if (!mixedradix_split(F, &b->N1, &b->N2) || !mr_schedule(b->N1, mr_tile_lines(b->N1, b->N2), &b->sc1) ||
        !mr_schedule(b->N2, mr_tile_lines(b->N2, b->N1), &b->sc2)) {
        snprintf(err, errLen, "fftSize %lld has no mixed-radix split", (long long)F);
        mixedradix_destroy(b);
        return nullptr;
    }
    MCK(cudaMalloc(&b->dWin, (size_t)F * 8));
    MCK(cudaMemcpyAsync(b->dWin, window, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    MCK(cudaMalloc(&b->dTab1, (size_t)b->N1 * 16));
    MCK(cudaMalloc(&b->dTab2, (size_t)b->N2 * 16));
    MCK(cudaMalloc(&b->dTabF, (size_t)F * 16));
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTab1, b->N1);
    mr_table_kernel<<<64, 256, 0, st>>>(b->dTab2, b->N2);
    mr_table_kernel<<<1024, 256, 0, st>>>(b->dTabF, F);
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
    const size_t smem1 = (size_t)2 * b->sc1.TC * b->N1 * 16, smem2 = (size_t)2 * b->sc2.TC * b->N2 * 16;
    auto kr = (b->sc2.TC * b->N2 <= MR_MAXA_STD * MR_THREADS) ? mr_rows_acc_kernel<MR_MAXA_STD> : mr_rows_acc_kernel<MR_MAXA_BIG>;
    auto kc = b->inFmt == KSPEC_IN_U8_IQ ? mr_cols_kernel<KSPEC_IN_U8_IQ> : (b->inFmt == KSPEC_IN_C64 ? mr_cols_kernel<KSPEC_IN_C64> : mr_cols_kernel<KSPEC_IN_C128>);
    if (cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess ||
        cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) {
        set_error("mixed-radix shared memory request failed: %s", cudaGetErrorString(cudaGetLastError()));
        return KSPEC_ERR_CUDA;
    }
    int per1 = 1, per2 = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per1, kc, MR_THREADS, smem1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, kr, MR_THREADS, smem2);
    if (per1 < 1) per1 = 1;
    if (per2 < 1) per2 = 1;
    double* dAcc = reinterpret_cast<double*>(acc);
    for (int64_t s0 = 0; s0 < nScans; s0 += chunk) {
        const int64_t ns = (nScans - s0 < chunk) ? nScans - s0 : chunk;
        const int64_t nfs = ns * nFrames;
        const void* smp = reinterpret_cast<const unsigned char*>(samples) + (size_t)s0 * scanStride * eb;
        MrColsParams pc{b->sc1, smp, scanStride, b->dOffs, nFrames, b->dWin, b->dTab1, b->dTabF, b->dZ, b->F, b->N1, b->N2, nfs, b->u8off, b->u8scale};
        const int64_t t1 = nfs * ((b->N2 + b->sc1.TC - 1) / b->sc1.TC);
        const int64_t cap1 = (int64_t)b->smCount * per1;
        kc<<<(int)(t1 < cap1 ? t1 : cap1), MR_THREADS, smem1, st>>>(pc);
        MrRowsParams pr{b->sc2, b->dZ, b->dTab2, dAcc + s0 * b->F, b->F, b->N1, b->N2, ns, nFrames, cumuMode};
        const int64_t t2 = ns * ((b->N1 + b->sc2.TC - 1) / b->sc2.TC);
        const int64_t cap2 = (int64_t)b->smCount * per2;
        kr<<<(int)(t2 < cap2 ? t2 : cap2), MR_THREADS, smem2, st>>>(pr);
        *launches += 2;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("mixed-radix FFT launch failed: %s", cudaGetErrorString(e)); return KSPEC_ERR_CUDA; }
    return KSPEC_OK;
}

}  // namespace kspec
