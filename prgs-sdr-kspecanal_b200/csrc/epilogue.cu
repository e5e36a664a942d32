// epilogue.cu — the small per-bin kernels around the fused scan kernel.
//
//   stats_finish     zero_span's cross-scan Max/Min/Avg (K:471-476) from per-team partials; Avg is the halving
//                    recurrence of data_cumu (K:137-139) replayed in float64 over the rows that can still matter.
//   scan_stitch      _scan_range's overlap stitch and Max/Min/Avg update (K:643-668), closed form per bin.
//   plotcompress     _data_plotcompress (K:184-200).
//   linear_epilogue  dB / stats / waterfall for engines that deliver one linear accumulation row per scan.
//   widen / narrow   T <-> float64 at the ABI boundary.
#include "kspec_internal.h"
#include "db_math.cuh"
#include <math.h>

namespace kspec {

namespace {

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// block = (FIN_X bins) x (FIN_Y slot groups): the per-team partials (888 of them after a batch of the headline shape) are reduced
// by FIN_Y threads per bin in parallel over many small CTAs (the kernel is pure latency: 14 MB out of L2), then a tree in shared
// memory and the y == 0 thread finishes the bin.  Long frames (F > FIN_WIDE_F: the multi-pass engines, a handful of scans per batch)
// have the parallelism in the bins instead: 128 bins x 4 slot groups per CTA, coalesced rows (the 8 x 64 shape spent 1.27 ms on
// 262 144 CTAs of mostly idle threads at F = 2^21).
constexpr int FIN_WIDE_F = 16384;
template <typename T, int FIN_X, int FIN_Y>
__global__ void __launch_bounds__(FIN_X * FIN_Y) stats_finish_kernel(const T* __restrict__ wsMax, const T* __restrict__ wsMin, int slots,
                                    const T* __restrict__ avgRows, int avgWin, int F, const double* __restrict__ carry,
                                    int firstIsSeed, double avgScale, double* __restrict__ out, int partialsLinear, T gain,
                                    PeerExchange px, unsigned long long seq) {
    __shared__ double shMax[FIN_Y][FIN_X + 1], shMin[FIN_Y][FIN_X + 1], shAvg[FIN_Y][FIN_X + 1];
    const int j = blockIdx.x * FIN_X + threadIdx.x;
    const bool inb = j < F;
    double mx = -d_inf(), mn = d_inf();
    // the rows of the Avg recurrence (at most AVG_WINDOW = FIN_Y of them) are fetched by the y threads side by side: the
    // sequential recurrence below then reads shared memory instead of paying one L2 round trip per row (8 x 64 shape only: the
    // wide shape has a bin per lane and reads its rows coalesced, straight from global memory)
    if (avgWin <= FIN_Y)
        for (int r = threadIdx.y; r < avgWin && inb; r += FIN_Y) shAvg[r][threadIdx.x] = (double)avgRows[(int64_t)r * F + j];
    if (inb) {
#pragma unroll 4
        for (int s = threadIdx.y; s < slots; s += FIN_Y) {
            mx = fmax(mx, (double)wsMax[(int64_t)s * F + j]);
            mn = fmin(mn, (double)wsMin[(int64_t)s * F + j]);
        }
    }
    shMax[threadIdx.y][threadIdx.x] = mx;
    shMin[threadIdx.y][threadIdx.x] = mn;
    __syncthreads();
    for (int h = FIN_Y / 2; h > 0; h >>= 1) {
        if (threadIdx.y < h) {
            shMax[threadIdx.y][threadIdx.x] = fmax(shMax[threadIdx.y][threadIdx.x], shMax[threadIdx.y + h][threadIdx.x]);
            shMin[threadIdx.y][threadIdx.x] = fmin(shMin[threadIdx.y][threadIdx.x], shMin[threadIdx.y + h][threadIdx.x]);
        }
        __syncthreads();
    }
    if (threadIdx.y == 0 && inb) {
        mx = shMax[0][threadIdx.x];
        mn = shMin[0][threadIdx.x];
        if (partialsLinear) {
            // the R32 kernels reduce the normalised LINEAR amplitudes (the dB map is monotone): the same conversion as the rows get
            mx = (double)(to_db((T)mx) - gain);
            mn = (double)(to_db((T)mn) - gain);
        }
        if (carry) {
            mx = fmax(mx, carry[j]);
            mn = fmin(mn, carry[F + j]);
        }
        double a;
        int r = 0;
        if (carry) a = carry[2 * F + j];
        else if (firstIsSeed) { a = (avgWin <= FIN_Y) ? shAvg[0][threadIdx.x] : (double)avgRows[j]; r = 1; }
        else a = 0.0;
        if (avgWin <= FIN_Y) {
            for (; r < avgWin; ++r) a = (a + shAvg[r][threadIdx.x]) / 2;
        } else {
#pragma unroll 8
            for (; r < avgWin; ++r) a = (a + (double)avgRows[(int64_t)r * F + j]) / 2;
        }
        a = (avgScale == 0.0) ? 0.0 : a * avgScale;
        out[j] = mx;
        out[F + j] = mn;
        out[2 * F + j] = a;
        // the exchange, folded into this kernel: the three values go straight into slot `rank` of every rank's symmetric buffer
        // (peer stores over NVLink; the local rank's own copy is one of them)
        if (px.nRanks > 0) {
            const size_t ofs = ((size_t)(seq & 1) * px.nRanks + px.rank) * 3 * (size_t)F;
            for (int r2 = 0; r2 < px.nRanks; ++r2) {
                double* d = px.slots[r2] + ofs;
                d[j] = mx;
                d[F + j] = mn;
                d[2 * F + j] = a;
            }
        }
    }
    if (px.nRanks > 0) {
        // last block done -> every write of this launch is visible system-wide -> raise this rank's flag on every rank
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const unsigned int done = atomicAdd(px.counter, 1u);
            if (done == gridDim.x - 1) {
                *px.counter = 0;
                __threadfence_system();
                for (int r2 = 0; r2 < px.nRanks; ++r2)
                    *reinterpret_cast<volatile unsigned long long*>(px.flags[r2] + (seq & 1) * px.nRanks + px.rank) = seq;
            }
        }
    }
}

// Second half of the peer exchange: wait until every rank's flag of this epoch carries `seq`, then MAX / MIN / SUM over the slots
// of the LOCAL symmetric buffer (the peers wrote into it).  A rank that never arrives trips a time-out (status = 1) instead of
// hanging the GPU.
__global__ void peer_combine_kernel(PeerExchange px, unsigned long long seq, double* __restrict__ out) {
    __shared__ int ok;
    const int F = px.F;
    if (threadIdx.x == 0) {
        ok = 1;
        const long long t0 = clock64();
        for (int r = 0; r < px.nRanks; ++r) {
            volatile unsigned long long* f = px.flags[px.rank] + (seq & 1) * px.nRanks + r;
            while (*f < seq) {
                if (clock64() - t0 > 4000000000LL) { ok = 0; *px.status = 1; break; }      // ~2 s
                __nanosleep(200);
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (!ok) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= F) return;
    const double* base = px.slots[px.rank] + (size_t)(seq & 1) * px.nRanks * 3 * (size_t)F;
    double mx = -d_inf(), mn = d_inf(), av = 0.0;
    for (int r = 0; r < px.nRanks; ++r) {
        const double* s = base + (size_t)r * 3 * F;
        mx = fmax(mx, __ldcg(s + j));
        mn = fmin(mn, __ldcg(s + F + j));
        av += __ldcg(s + 2 * F + j);                  // rank order: the same sum on every rank
    }
    out[j] = mx;
    out[F + j] = mn;
    out[2 * F + j] = av;
}

template <typename T> __global__ void widen_kernel(const T* __restrict__ s, double* __restrict__ d, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (double)s[i];
}
template <typename T> __global__ void narrow_kernel(const double* __restrict__ s, T* __restrict__ d, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (T)s[i];
}

template <typename T>
__global__ void scan_stitch_kernel(const T* __restrict__ rows, const uint8_t* __restrict__ ok, const int64_t* __restrict__ iStart,
                                   const int64_t* __restrict__ iDone, int nSteps, int F, int64_t total, double failValue,
                                   int baseIsRaw, int passIndex, double* __restrict__ cur, double* __restrict__ mx,
                                   double* __restrict__ mn, double* __restrict__ av) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total) return;
    // i1 = last step with iStart <= b ; i0 = first step with iStart + F > b  (iStart is non-decreasing)
    int lo = 0, hi = nSteps;            // upper_bound(iStart, b)
    while (lo < hi) { int mid = (lo + hi) >> 1; if (iStart[mid] <= b) lo = mid + 1; else hi = mid; }
    const int i1 = lo - 1;
    lo = 0; hi = nSteps;                // first with iStart > b - F
    while (lo < hi) { int mid = (lo + hi) >> 1; if (iStart[mid] + F > b) hi = mid; else lo = mid + 1; }
    const int i0 = lo;
    if (i1 < 0 || i0 > i1) return;
    auto val = [&](int i) -> double {
        if (ok && !ok[i]) return failValue;
        return (double)rows[(int64_t)i * F + (b - iStart[i])];
    };
    double c = val(i0);
    for (int i = i0 + 1; i <= i1; ++i) c = (c + val(i)) / 2;     // RAW for the first cover, halving AVG afterwards (K:643-650)
    cur[b] = c;
    if (baseIsRaw) {                                            // K:651-656: every covering step updates the stats with its own row
        double m1 = mx[b], m2 = mn[b], a = av[b];
        for (int i = i0; i <= i1; ++i) {
            const double v = val(i);
            m1 = fmax(m1, v); m2 = fmin(m2, v);
            a = (passIndex == 0) ? v : (a + v) / 2;
        }
        mx[b] = m1; mn[b] = m2; av[b] = a;
    } else if (b < iDone[i1]) {                                 // K:657-668: slice [iStart:iDone) of the finished Fft.Cur
        mx[b] = fmax(mx[b], c);
        mn[b] = fmin(mn[b], c);
        av[b] = (passIndex == 0) ? c : (av[b] + c) / 2;
    }
}

// Sharded stepped scan: this rank holds the dB rows of steps [stepBase, stepBase+nLocal).  The sequential
// stitch "first cover RAW, later covers (cur+new)/2" (K:643-650) of a bin covered by steps i0..i1 is the weighted sum
//   cur = x_i0 * 2^-(i1-i0) + sum_{i0 < i <= i1} x_i * 2^-(i1-i+1),
// so every rank writes the part of that sum its own steps contribute and a SUM over ranks gives Fft.Cur.
template <typename T>
__global__ void scan_stitch_partial_kernel(const T* __restrict__ rows, const uint8_t* __restrict__ ok, const int64_t* __restrict__ iStart,
                                           int nSteps, int stepBase, int nLocal, int F, int64_t total, double failValue,
                                           double* __restrict__ curPartial) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total) return;
    int lo = 0, hi = nSteps;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (iStart[mid] <= b) lo = mid + 1; else hi = mid; }
    const int i1 = lo - 1;
    lo = 0; hi = nSteps;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (iStart[mid] + F > b) hi = mid; else lo = mid + 1; }
    const int i0 = lo;
    double acc = 0.0;
    if (i1 >= 0 && i0 <= i1) {
        const int a = i0 > stepBase ? i0 : stepBase;
        const int z = i1 < stepBase + nLocal - 1 ? i1 : stepBase + nLocal - 1;
        for (int i = a; i <= z; ++i) {
            const double v = (ok && !ok[i - stepBase]) ? failValue : (double)rows[(int64_t)(i - stepBase) * F + (b - iStart[i])];
            const int sh = (i == i0) ? (i1 - i0) : (i1 - i + 1);
            acc += ldexp(v, -sh);
        }
    }
    curPartial[b] = acc;
}

// Max/Min/Avg update from a finished Fft.Cur on the slices [iStart_i, iDone_i) (K:657-668); bins past the last iDone keep
// their state.  Used after the SUM over ranks of the sharded stitch.
__global__ void scan_stats_update_kernel(const double* __restrict__ cur, int64_t total, int64_t lastDone, int passIndex,
                                         double* __restrict__ mx, double* __restrict__ mn, double* __restrict__ av) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total || b >= lastDone) return;
    const double c = cur[b];
    mx[b] = fmax(mx[b], c);
    mn[b] = fmin(mn[b], c);
    av[b] = (passIndex == 0) ? c : (av[b] + c) / 2;
}

__global__ void plotcompress_kernel(const double* __restrict__ y, int64_t g, int mode, double* __restrict__ out) {
    __shared__ double sh[256];
    const int64_t w = blockIdx.x;
    const double* p = y + w * g;
    double r = (mode == KSPEC_COMPRESS_MAX) ? -d_inf() : (mode == KSPEC_COMPRESS_MIN ? d_inf() : 0.0);
    for (int64_t q = threadIdx.x; q < g; q += blockDim.x) {
        const double v = p[q];
        r = (mode == KSPEC_COMPRESS_MAX) ? fmax(r, v) : (mode == KSPEC_COMPRESS_MIN ? fmin(r, v) : r + v);
    }
    sh[threadIdx.x] = r;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double a = sh[threadIdx.x], c = sh[threadIdx.x + s];
            sh[threadIdx.x] = (mode == KSPEC_COMPRESS_MAX) ? fmax(a, c) : (mode == KSPEC_COMPRESS_MIN ? fmin(a, c) : a + c);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[w] = (mode == KSPEC_COMPRESS_AVG) ? sh[0] / (double)g : sh[0];
}

__device__ __forceinline__ float db_of(float v) { return 10.0f * log10f(v); }
__device__ __forceinline__ double db_of(double v) { return 10.0 * log10(v); }

// one thread per shifted bin j; loops over the scans of the batch in order
template <typename T>
__global__ void linear_epilogue_kernel(const ScanParams p, const T* __restrict__ acc, int F) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= F) return;
    // np.fft.fftshift (K:396) rolls by F//2:  out = concat(in[ceil(F/2):], in[:ceil(F/2)])  ->  out[j] = in[(j + ceil(F/2)) mod F]
    const int c2 = (F + 1) / 2;
    const int srcBin = p.accShifted ? j : (j + c2) % F;
    const int src = p.accL1 ? (((srcBin & ((1 << p.accL1) - 1)) << p.accL2) + (srcBin >> p.accL1)) : srcBin;
    T* rows = reinterpret_cast<T*>(p.rows);
    T mx = 0, mn = 0;
    for (int64_t s = 0; s < p.nScans; ++s) {
        T lin = acc[s * F + src] * (T)p.linScale;
        if (p.rowsKind == KSPEC_ROWS_LINEAR) rows[s * F + j] = lin;
        if (p.rowsKind == KSPEC_ROWS_DB || p.wantStats) {
            if (p.dbClip) lin = fmax(lin, (T)p.minAmp);
            T db = db_of(lin) - (T)p.gain;
            if (p.infToZero && isinf(db)) db = (T)0;
            if (p.rowsKind == KSPEC_ROWS_DB) rows[s * F + j] = db;
            if (p.wantStats) {
                mx = (s == 0) ? db : fmax(mx, db);
                mn = (s == 0) ? db : fmin(mn, db);
                const int64_t ar = s - (p.nScans - p.avgWin);
                if (ar >= 0) reinterpret_cast<T*>(p.avgRows)[ar * F + j] = db;
            }
        }
    }
    if (p.wantStats) {
        reinterpret_cast<T*>(p.wsMax)[j] = mx;
        reinterpret_cast<T*>(p.wsMin)[j] = mn;
    }
}

// the same per-row work without the cross-scan statistics (stepped scans: clip, dB, inf -> 0 only; K:640-641): one thread per
// (scan, bin), so that many short rows -- 1226 steps of 64 bins in quickFullScan -- do not serialise behind 64 threads
template <typename T>
__global__ void linear_rows_kernel(const ScanParams p, const T* __restrict__ acc, int F) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nScans * F) return;
    const int64_t s = i / F;
    const int j = (int)(i - s * F);
    const int c2 = (F + 1) / 2;
    const int srcBin = p.accShifted ? j : (j + c2) % F;
    const int src = p.accL1 ? (((srcBin & ((1 << p.accL1) - 1)) << p.accL2) + (srcBin >> p.accL1)) : srcBin;
    T lin = acc[s * F + src] * (T)p.linScale;
    T* rows = reinterpret_cast<T*>(p.rows);
    if (p.rowsKind == KSPEC_ROWS_LINEAR) { rows[i] = lin; return; }
    if (p.dbClip) lin = fmax(lin, (T)p.minAmp);
    T db = db_of(lin) - (T)p.gain;
    if (p.infToZero && isinf(db)) db = (T)0;
    rows[i] = db;
}

// waterfall rows for the linear engines: block per (scan, output column)
template <typename T>
__global__ void linear_hm_kernel(const ScanParams p, const T* __restrict__ acc, int F) {
    __shared__ T sh[256];
    const int W = p.hmW, g = F / W;
    const int64_t s = blockIdx.x / W;
    const int w = blockIdx.x % W;
    const int c2 = (F + 1) / 2;
    const int mode = p.hmMode;
    const T inf = (T)d_inf();
    T r = (mode == KSPEC_COMPRESS_MAX) ? -inf : (mode == KSPEC_COMPRESS_MIN ? inf : (T)0);
    for (int q = threadIdx.x; q < g; q += blockDim.x) {
        const int j = w * g + q;
        const int srcBin = p.accShifted ? j : (j + c2) % F;
        const int src = p.accL1 ? (((srcBin & ((1 << p.accL1) - 1)) << p.accL2) + (srcBin >> p.accL1)) : srcBin;
        T lin = acc[s * F + src] * (T)p.linScale;
        if (p.dbClip) lin = fmax(lin, (T)p.minAmp);
        T db = db_of(lin) - (T)p.gain;
        if (p.infToZero && isinf(db)) db = (T)0;
        if (p.adj) db -= reinterpret_cast<const T*>(p.adj)[j];
        r = (mode == KSPEC_COMPRESS_MAX) ? fmax(r, db) : (mode == KSPEC_COMPRESS_MIN ? fmin(r, db) : r + db);
    }
    sh[threadIdx.x] = r;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            const T a = sh[threadIdx.x], c = sh[threadIdx.x + st];
            sh[threadIdx.x] = (mode == KSPEC_COMPRESS_MAX) ? fmax(a, c) : (mode == KSPEC_COMPRESS_MIN ? fmin(a, c) : a + c);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        T v = sh[0];
        if (mode == KSPEC_COMPRESS_AVG) v /= (T)g;
        reinterpret_cast<T*>(p.hm)[s * W + w] = v;
    }
}

// plot_highs (K:243-272): walk the points by descending level; mark a point unless an already marked one lies closer
// than delta in x; stop after numMarkers.  One block: numMarkers rounds of a block-wide arg-max over the still
// admissible points (equal levels: lowest index first; the reference's unstable argsort leaves that order open).
__global__ void plot_highs_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, int numMarkers, double delta,
                                  int64_t* __restrict__ idxOut, int* __restrict__ nOut) {
    __shared__ double shV[1024];
    __shared__ long long shI[1024];
    __shared__ double marked[64];
    __shared__ long long markedIdx[64];
    __shared__ int nMarked;
    __shared__ long long lowest;
    if (threadIdx.x == 0) nMarked = 0;
    // the reference's loop np.arange(-1, -len, -1) never reaches the smallest element: find it once and skip it
    double bv = d_inf();
    long long bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        if (y[i] < bv) { bv = y[i]; bi = i; }
    shV[threadIdx.x] = bv; shI[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double a = shV[threadIdx.x + s]; const long long ai = shI[threadIdx.x + s];
            if (ai >= 0 && (shI[threadIdx.x] < 0 || a < shV[threadIdx.x] || (a == shV[threadIdx.x] && ai < shI[threadIdx.x]))) { shV[threadIdx.x] = a; shI[threadIdx.x] = ai; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) lowest = shI[0];
    __syncthreads();
    for (int m = 0; m < numMarkers; ++m) {
        const int nm = nMarked;
        bv = -d_inf(); bi = -1;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            if (i == lowest) continue;
            const double xi = x[i];
            bool ok = true;
            for (int k = 0; k < nm; ++k) ok = ok && !(fabs(marked[k] - xi) < delta) && markedIdx[k] != i;   // each point is visited once
            if (ok && (bi < 0 || y[i] > bv)) { bv = y[i]; bi = i; }
        }
        shV[threadIdx.x] = bv; shI[threadIdx.x] = bi;
        __syncthreads();
        for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) {
                const double a = shV[threadIdx.x + s]; const long long ai = shI[threadIdx.x + s];
                if (ai >= 0 && (shI[threadIdx.x] < 0 || a > shV[threadIdx.x] || (a == shV[threadIdx.x] && ai < shI[threadIdx.x]))) { shV[threadIdx.x] = a; shI[threadIdx.x] = ai; }
            }
            __syncthreads();
        }
        if (shI[0] < 0) break;                     // nothing admissible left (uniform: every thread reads the same slot)
        if (threadIdx.x == 0) { idxOut[nm] = shI[0]; marked[nm] = x[shI[0]]; markedIdx[nm] = shI[0]; nMarked = nm + 1; }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) *nOut = nMarked;
}

// data_proc 'Conv' (K:113-120): np.convolve(vals, taps, 'same') then the first/last 12 points := mean of the result
__global__ void conv_same_kernel(const double* __restrict__ v, int64_t n, const double* __restrict__ taps, int m, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = i + (m - 1) / 2;            // index into the full convolution
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
        const int64_t a = j - k;
        if (a >= 0 && a < n) acc += v[a] * taps[k];
    }
    out[i] = acc;
}
__global__ void conv_edges_kernel(double* __restrict__ out, int64_t n, int edge) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += out[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int st = blockDim.x >> 1; st > 0; st >>= 1) { if ((int)threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st]; __syncthreads(); }
    const double avg = sh[0] / (double)n;
    __syncthreads();
    for (int i = threadIdx.x; i < edge && i < n; i += blockDim.x) { out[i] = avg; out[n - 1 - i] = avg; }
}

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

}  // namespace

void launch_stats_finish(int prec, const void* wsMax, const void* wsMin, int slots, const void* avgRows, int avgWin, int F,
                         const double* carry, int firstIsSeed, double avgScale, double* out, cudaStream_t st, int partialsLinear,
                         double gain, const PeerExchange* px, unsigned long long seq) {
    PeerExchange none;
    const PeerExchange& pe = px ? *px : none;
    const bool wide = F > FIN_WIDE_F;
    if (prec == KSPEC_PREC_F32) {
        auto args = [&](auto k, int fx, int fy) {
            k<<<nblk(F, fx), dim3(fx, fy), 0, st>>>((const float*)wsMax, (const float*)wsMin, slots, (const float*)avgRows, avgWin, F, carry,
                                                    firstIsSeed, avgScale, out, partialsLinear, (float)gain, pe, seq);
        };
        if (wide) args(stats_finish_kernel<float, 128, 4>, 128, 4);
        else args(stats_finish_kernel<float, 8, 64>, 8, 64);
    } else {
        auto args = [&](auto k, int fx, int fy) {
            k<<<nblk(F, fx), dim3(fx, fy), 0, st>>>((const double*)wsMax, (const double*)wsMin, slots, (const double*)avgRows, avgWin, F, carry,
                                                    firstIsSeed, avgScale, out, partialsLinear, gain, pe, seq);
        };
        if (wide) args(stats_finish_kernel<double, 128, 4>, 128, 4);
        else args(stats_finish_kernel<double, 8, 64>, 8, 64);
    }
}

void launch_peer_combine(const PeerExchange& px, unsigned long long seq, double* out, cudaStream_t st) {
    peer_combine_kernel<<<nblk(px.F, 128), 128, 0, st>>>(px, seq, out);
}

void launch_widen(int prec, const void* src, double* dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const int g = (int)(n / 256 + 1 < 4096 ? n / 256 + 1 : 4096);
    if (prec == KSPEC_PREC_F32) widen_kernel<float><<<g, 256, 0, st>>>((const float*)src, dst, n);
    else widen_kernel<double><<<g, 256, 0, st>>>((const double*)src, dst, n);
}

void launch_narrow(int prec, const double* src, void* dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const int g = (int)(n / 256 + 1 < 4096 ? n / 256 + 1 : 4096);
    if (prec == KSPEC_PREC_F32) narrow_kernel<float><<<g, 256, 0, st>>>(src, (float*)dst, n);
    else narrow_kernel<double><<<g, 256, 0, st>>>(src, (double*)dst, n);
}

void launch_scan_stitch(int prec, const void* dbRows, const uint8_t* stepOk, const int64_t* iStart, const int64_t* iDone,
                        int nSteps, int F, int64_t total, double failValue, int baseIsRaw, int passIndex, double* cur,
                        double* mx, double* mn, double* av, cudaStream_t st) {
    if (prec == KSPEC_PREC_F32)
        scan_stitch_kernel<float><<<nblk(total, 256), 256, 0, st>>>((const float*)dbRows, stepOk, iStart, iDone, nSteps, F, total,
                                                                    failValue, baseIsRaw, passIndex, cur, mx, mn, av);
    else
        scan_stitch_kernel<double><<<nblk(total, 256), 256, 0, st>>>((const double*)dbRows, stepOk, iStart, iDone, nSteps, F,
                                                                     total, failValue, baseIsRaw, passIndex, cur, mx, mn, av);
}

void launch_plot_highs(const double* x, const double* y, int64_t n, int numMarkers, double delta, int64_t* idxOut, int* nOut, cudaStream_t st) {
    plot_highs_kernel<<<1, 1024, 0, st>>>(x, y, n, numMarkers, delta, idxOut, nOut);
}

void launch_conv_same(const double* v, int64_t n, const double* taps, int m, int edge, double* out, cudaStream_t st) {
    conv_same_kernel<<<nblk(n, 256), 256, 0, st>>>(v, n, taps, m, out);
    conv_edges_kernel<<<1, 256, 0, st>>>(out, n, edge);
}

void launch_scan_stitch_partial(int prec, const void* dbRows, const uint8_t* stepOk, const int64_t* iStart, int nSteps, int stepBase,
                                int nLocal, int F, int64_t total, double failValue, double* curPartial, cudaStream_t st) {
    if (prec == KSPEC_PREC_F32)
        scan_stitch_partial_kernel<float><<<nblk(total, 256), 256, 0, st>>>((const float*)dbRows, stepOk, iStart, nSteps, stepBase, nLocal, F,
                                                                            total, failValue, curPartial);
    else
        scan_stitch_partial_kernel<double><<<nblk(total, 256), 256, 0, st>>>((const double*)dbRows, stepOk, iStart, nSteps, stepBase, nLocal,
                                                                             F, total, failValue, curPartial);
}

void launch_scan_stats_update(const double* cur, int64_t total, int64_t lastDone, int passIndex, double* mx, double* mn, double* av,
                              cudaStream_t st) {
    scan_stats_update_kernel<<<nblk(total, 256), 256, 0, st>>>(cur, total, lastDone, passIndex, mx, mn, av);
}

void launch_plotcompress(const double* y, int64_t n, int xRes, int mode, double* out, cudaStream_t st) {
    const int64_t g = n / xRes;
    plotcompress_kernel<<<xRes, 256, 0, st>>>(y, g, mode, out);
}

// frame-parallel small batches: rows[(s*nFrames + f)*F + j] holds frame f of scan s (normalised, shifted); data_cumu
// (K:124-147) over the frames in their order, one thread per (scan, bin)
template <typename T>
__global__ void frames_combine_kernel(const T* __restrict__ rows, T* __restrict__ out, int64_t nScans, int nFrames, int F, int cumuMode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nScans * F) return;
    const int64_t s = i / F;
    const int j = (int)(i - s * F);
    const T* r = rows + s * nFrames * F + j;
    T a = r[0];
    if (cumuMode == KSPEC_CUMU_RAW) a = r[(int64_t)(nFrames - 1) * F];
    else if (cumuMode == KSPEC_CUMU_AVG) { for (int f = 1; f < nFrames; ++f) a = (a + r[(int64_t)f * F]) * (T)0.5; }
    else if (cumuMode == KSPEC_CUMU_MAX) { for (int f = 1; f < nFrames; ++f) a = fmax(a, r[(int64_t)f * F]); }
    else if (cumuMode == KSPEC_CUMU_MIN) { for (int f = 1; f < nFrames; ++f) a = fmin(a, r[(int64_t)f * F]); }
    else { for (int f = 1; f < nFrames; ++f) a += r[(int64_t)f * F]; }
    out[i] = a;
}

void launch_frames_combine(int prec, const void* rows, void* out, int64_t nScans, int nFrames, int F, int cumuMode, cudaStream_t st) {
    const unsigned g = nblk(nScans * F, 256);
    if (prec == KSPEC_PREC_F32) frames_combine_kernel<float><<<g, 256, 0, st>>>((const float*)rows, (float*)out, nScans, nFrames, F, cumuMode);
    else frames_combine_kernel<double><<<g, 256, 0, st>>>((const double*)rows, (double*)out, nScans, nFrames, F, cumuMode);
}

void launch_linear_epilogue(int prec, const ScanParams& p, const void* acc, int F, int slots, cudaStream_t st) {
    (void)slots;
    if (!p.wantStats && p.rowsKind != KSPEC_ROWS_NONE) {
        const unsigned g = nblk(p.nScans * F, 256);
        if (prec == KSPEC_PREC_F32) linear_rows_kernel<float><<<g, 256, 0, st>>>(p, (const float*)acc, F);
        else linear_rows_kernel<double><<<g, 256, 0, st>>>(p, (const double*)acc, F);
    } else if (prec == KSPEC_PREC_F32) {
        linear_epilogue_kernel<float><<<nblk(F, 256), 256, 0, st>>>(p, (const float*)acc, F);
    } else {
        linear_epilogue_kernel<double><<<nblk(F, 256), 256, 0, st>>>(p, (const double*)acc, F);
    }
    if (p.hm) {
        if (prec == KSPEC_PREC_F32) linear_hm_kernel<float><<<(unsigned)(p.nScans * p.hmW), 256, 0, st>>>(p, (const float*)acc, F);
        else linear_hm_kernel<double><<<(unsigned)(p.nScans * p.hmW), 256, 0, st>>>(p, (const double*)acc, F);
    }
}

}  // namespace kspec
