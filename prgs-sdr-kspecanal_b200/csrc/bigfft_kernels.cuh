// bigfft_kernels.cuh — float64 multi-pass engines for frames that do not fit the fused shared-memory kernel.
//
//   four-step   F = F1*F2 (both powers of two <= 4096): column FFTs + twiddle, then row FFTs.  cfg-4 of BASELINE.json
//               (fftSize 2^21, "full-sample-rate window") runs here; intermediates stay L2-resident for F <= 2^22.
//   Bluestein   any F (2.4e6 of cfg-5, 1200, odd sizes): chirp-z with a power-of-two circular convolution of length
//               M >= 2F-1:  |X_k| = |FFT_M( conj( FFT_M(x.w.c) . V ) )_k| / M,  c_n = exp(-i pi n^2/F), V = FFT_M(conj c).
//               The final chirp multiply has modulus one and the path only needs |X| (K:391), so it is skipped.
//
// Both replace the single np.fft.fft call at kspecanal.py:391; numpy itself would use mixed radix for 2.4e6.
// The per-team register FFT is the one of fft_core.cuh; "ops" below describe where a team's elements come from and
// where its bins go, so one kernel template serves every pass.
#pragma once
#include "curscan_smem.cuh"

namespace kspec {

typedef double2 cd;

__device__ __forceinline__ cd cconj(cd a) { a.y = -a.y; return a; }

// geometry shared by the pass ops: transform length M = L1*L2, element n = n1*L2 + n2, bin k = k1 + L1*k2
struct BigGeom {
    int64_t M;        // transform length
    int64_t F;        // frame length (== M for four-step, < M zero padded for Bluestein)
    int l1, l2;       // log2 L1, log2 L2
};

// ---- pass ops ---------------------------------------------------------------------------------------------------
// Every op provides   Ctx begin(int64_t q)   (the per-item index arithmetic, done once: the 64-bit division by nFrames used to be
// repeated for each of a thread's 16 elements and kept their loads from being issued back to back),
// cd load(const Ctx&, int e)   and   void store(const Ctx&, int k, cd v)   for batch item q.
// One launch covers every frame of every scan of a chunk: q = fs * L2 + n2 (column passes) or fs * L1 + k1 (row passes),
// fs = scan * nFrames + frame; the work vectors Z and P hold one M-point slab per fs.

// column pass, first transform: gather the frame (fused ingest, window, optional chirp, zero padding)
// BLUE: chirp multiply and zero padding (n >= F).  Both forms are branch-free so that a thread's 16 element loads issue back
// to back: with a per-element branch every load sat in its own basic block and cost a full round trip.
template <int INFMT, bool BLUE> struct OpColsIn {
    BigGeom g;
    const void* samples; int64_t scanStride; const int64_t* offs; int nFrames;
    const double* win; const cd* chirp;   // chirp == nullptr for the plain four-step transform
    const cd* twM; cd* Z;
    double u8off, u8scale;
    struct Ctx { int64_t base, n2, zbase; };       // first sample of the frame, column, first element of the frame's slab
    __device__ __forceinline__ Ctx begin(int64_t q) const {
        const int64_t fs = q >> g.l2, n2 = q & (((int64_t)1 << g.l2) - 1);
        const int64_t s = fs / nFrames;
        return Ctx{s * scanStride + __ldg(&offs[fs - s * nFrames]), n2, fs * g.M};
    }
    __device__ __forceinline__ cd load(const Ctx& c, int e) const {
        const int64_t n = ((int64_t)e << g.l2) + c.n2;
        if constexpr (!BLUE) {
            return Ingest<double, INFMT>::load(samples, c.base + n, __ldg(&win[n]), u8off, u8scale);
        } else {
            const bool in = n < g.F;
            const int64_t m = in ? n : 0;
            const cd v = cmul(Ingest<double, INFMT>::load(samples, c.base + m, __ldg(&win[m]), u8off, u8scale), __ldg(&chirp[m]));
            return in ? v : make_double2(0.0, 0.0);
        }
    }
    __device__ __forceinline__ void store(const Ctx& c, int k, cd v) const {
        Z[c.zbase + ((int64_t)k << g.l2) + c.n2] = cmul(v, __ldg(&twM[c.n2 * k]));
    }
};

// column pass on a natural-order device vector (precomputing V; single slab)
struct OpColsPlain {
    BigGeom g; const cd* X; const cd* twM; cd* Z;
    typedef int64_t Ctx;
    __device__ __forceinline__ Ctx begin(int64_t q) const { return q; }
    __device__ __forceinline__ cd load(const Ctx& q, int e) const { return X[((int64_t)e << g.l2) + q]; }
    __device__ __forceinline__ void store(const Ctx& q, int k, cd v) const { Z[((int64_t)k << g.l2) + q] = cmul(v, __ldg(&twM[q * k])); }
};

// column pass of the SECOND Bluestein transform: its input P sits in the row pass's output layout [k1*L2 + k2]
struct OpColsMid {
    BigGeom g; const cd* P; const cd* twM; cd* Z;
    struct Ctx { int64_t zbase, n2, loc0; };
    __device__ __forceinline__ Ctx begin(int64_t q) const {
        // element n = e*L2 + n2 of the natural-order vector lives at (n mod L1)*L2 + n div L1   (L2 >= L1)
        const int64_t fs = q >> g.l2, n2 = q & (((int64_t)1 << g.l2) - 1);
        const int64_t L1m = ((int64_t)1 << g.l1) - 1;
        return Ctx{fs * g.M, n2, fs * g.M + ((n2 & L1m) << g.l2) + (n2 >> g.l1)};
    }
    __device__ __forceinline__ cd load(const Ctx& c, int e) const { return P[c.loc0 + ((int64_t)e << (g.l2 - g.l1))]; }
    __device__ __forceinline__ void store(const Ctx& c, int k, cd v) const {
        Z[c.zbase + ((int64_t)k << g.l2) + c.n2] = cmul(v, __ldg(&twM[c.n2 * k]));
    }
};

// row pass storing the spectrum as is, layout [k1*L2 + k2] (precomputing V; single slab)
struct OpRowsPlain {
    BigGeom g; const cd* Z; cd* out;
    typedef int64_t Ctx;
    __device__ __forceinline__ Ctx begin(int64_t q) const { return q << g.l2; }
    __device__ __forceinline__ cd load(const Ctx& o, int e) const { return Z[o + e]; }
    __device__ __forceinline__ void store(const Ctx& o, int k, cd v) const { out[o + k] = v; }
};

// row pass of the first Bluestein transform: P = conj(U . V)
struct OpRowsMul {
    BigGeom g; const cd* Z; const cd* V; cd* P;
    struct Ctx { int64_t zrow, vrow; };            // first element of the row in the slab / in V
    __device__ __forceinline__ Ctx begin(int64_t q) const {
        const int64_t fs = q >> g.l1, k1 = q & (((int64_t)1 << g.l1) - 1);
        return Ctx{fs * g.M + (k1 << g.l2), k1 << g.l2};
    }
    __device__ __forceinline__ cd load(const Ctx& c, int e) const { return Z[c.zrow + e]; }
    __device__ __forceinline__ void store(const Ctx& c, int k, cd v) const { P[c.zrow + k] = cconj(cmul(v, __ldg(&V[c.vrow + k]))); }
};

// contiguous batched transform on device vectors (V for the small Bluestein, self tests)
struct OpPlain {
    int l; const cd* X; cd* Y;
    typedef int64_t Ctx;
    __device__ __forceinline__ Ctx begin(int64_t q) const { return q << l; }
    __device__ __forceinline__ cd load(const Ctx& o, int e) const { return X[o + e]; }
    __device__ __forceinline__ void store(const Ctx& o, int k, cd v) const { Y[o + k] = v; }
};

// ---- one kernel for all passes: a team of NT threads transforms batch item q ---------------------------------------
template <int LOG2L, typename Op>
__global__ void __launch_bounds__(SmemCfg<double, LOG2L>::CTA, 1)
team_fft_kernel(const Op op, const cd* __restrict__ tw, int64_t nBatch) {
    using C = SmemCfg<double, LOG2L>;
    constexpr int P = C::P, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P;
    constexpr int L0 = stage_l<LOG2L, LOG2P>(0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    cd* bufA = reinterpret_cast<cd*>(smem_raw) + team * C::FPAD;
    cd* bufB = C::DBUF ? bufA + TEAMS * C::FPAD : bufA;
    auto sync = [] { __syncthreads(); };
    const int64_t perIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (nBatch + perIter - 1) / perIter;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t q = it * perIter + (int64_t)blockIdx.x * TEAMS + team;
        const bool valid = q < nBatch;
        const int64_t qc = valid ? q : nBatch - 1;
        const typename Op::Ctx ctx = op.begin(qc);
        cd b[P];
#pragma unroll
        for (int m = 0; m < P; ++m) b[m] = op.load(ctx, tid + NT * m);
        butterflies<double, P, (1 << L0), false>(b, nullptr);
        fft_tail<double, LOG2L, LOG2P, false, C::DBUF, L0, 0, 0>(b, nullptr, tw, bufA, bufB, tid, sync);
        if constexpr (C::DBUF && (C::NX & 1)) { cd* t = bufA; bufA = bufB; bufB = t; }
        if (valid) {
#pragma unroll
            for (int m = 0; m < P; ++m) op.store(ctx, tid + NT * m, b[m]);
        }
    }
}

template <int LOG2L, typename Op>
static int launch_team_fft(const Op& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    using C = SmemCfg<double, LOG2L>;
    auto k = team_fft_kernel<LOG2L, Op>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int64_t need = (nBatch + C::TEAMS - 1) / C::TEAMS;
    int64_t cap = (int64_t)smCount * 4;
    int grid = (int)(need < cap ? need : cap);
    k<<<grid, C::CTA, C::SMEM_BYTES, st>>>(op, tw, nBatch);
    return (int)cudaGetLastError();
}

// ---- column pass of the four-step transform, tiled ------------------------------------------------------------------
// team_fft_kernel<OpColsIn> reads a column element by element: the lanes of a warp hold rows e, e+1, ... of ONE column, so every
// sample, window value, twiddle and Z element of a warp request lies in its own 32-byte sector (stride L2 elements) and the L1
// tag stage, not HBM, sets the pace (profiles/r2_ncu_fourstep.txt: l1tex 74 %, DRAM 1.2 TB/s).  Here a CTA owns a tile of
// TC = 4096/L1 adjacent columns of one frame (64 KB of double2 whatever L1 is):
//   1. all threads load the tile row by row, columns fastest: runs of TC samples / window values per row, ingest and window
//      fused, products into shared memory;
//   2. each team transforms its columns out of / back into the tile (in place: a thread reads rows tid + NT*m of its column and
//      writes bins tid + NT*m of the same column), the exchange buffer of fft_tail beside it;
//      bins leave multiplied by W_M^(n2*k), k = tid + NT*m: W_M^(n2*tid) times successive powers of W_M^(n2*NT), two table
//      reads per thread and column (a per-element gather in step 3 cost a third of the kernel's stall samples);
//   3. all threads write the tile to Z row by row, runs of TC elements.
// The tile is swizzled (column c of row e sits in slot c ^ f(e)) so that both the row-wise and the column-wise accesses are free
// of bank conflicts with 16-byte elements.
template <int LOG2L> struct ColsTileCfg {
    using C = SmemCfg<double, LOG2L>;
    static constexpr int L = 1 << LOG2L;
    static constexpr int LOG2TC = 12 - LOG2L, TC = 1 << LOG2TC;
    static constexpr int ROUNDS = TC / C::TEAMS;
    static constexpr int SW_SHIFT = TC >= 8 ? 0 : (TC == 4 ? 1 : 2);
    static constexpr int SW_MASK = (TC >= 8 ? 8 : TC) - 1;
    static constexpr int TILE_ELEMS = L * TC;
    static constexpr int XCH_ELEMS = C::FPAD * C::TEAMS;                 // single exchange buffer (two barriers per exchange)
    static constexpr int SMEM_BYTES = (TILE_ELEMS + XCH_ELEMS) * (int)sizeof(cd);
    static_assert(TC >= C::TEAMS && TC % C::TEAMS == 0, "a tile is a whole number of rounds of the CTA's teams");
    static_assert(TILE_ELEMS % C::CTA == 0, "the row-wise phases have no remainder");
    static __device__ __forceinline__ int slot(int e, int c) { return e * TC + (c ^ ((e >> SW_SHIFT) & SW_MASK)); }
};

template <int LOG2L, int INFMT>
__global__ void __launch_bounds__(SmemCfg<double, LOG2L>::CTA, 2)
cols_tiled_kernel(const OpColsIn<INFMT, false> op, const cd* __restrict__ tw, int64_t nTiles) {
    using C = SmemCfg<double, LOG2L>;
    using TCfg = ColsTileCfg<LOG2L>;
    constexpr int P = C::P, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P, CTA = C::CTA;
    constexpr int L0 = stage_l<LOG2L, LOG2P>(0);
    constexpr int TC = TCfg::TC, LOG2TC = TCfg::LOG2TC;
    constexpr int PER = TCfg::TILE_ELEMS / CTA, UNR = 16;        // 32 elements per thread and phase, two bursts of 16 (x2) loads
    static_assert(PER % UNR == 0, "the row-wise phases are unrolled");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tile = reinterpret_cast<cd*>(smem_raw);
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    cd* xbuf = tile + TCfg::TILE_ELEMS + team * C::FPAD;
    auto sync = [] { __syncthreads(); };
    const int l2 = op.g.l2;
    const int tilesPerFrameLog = l2 - LOG2TC;
    for (int64_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const int64_t fs = t >> tilesPerFrameLog;
        const int64_t n2_0 = (t & (((int64_t)1 << tilesPerFrameLog) - 1)) << LOG2TC;
        const int64_t s = fs / op.nFrames;
        const int64_t base = s * op.scanStride + __ldg(&op.offs[fs - s * op.nFrames]);
        // 1. rows of the tile, columns fastest
#pragma unroll 1
        for (int i0 = 0; i0 < PER; i0 += UNR) {
            cd v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = (i0 + u) * CTA + threadIdx.x;
                const int64_t n = ((int64_t)(i >> LOG2TC) << l2) + n2_0 + (i & (TC - 1));
                v[u] = Ingest<double, INFMT>::load(op.samples, base + n, __ldg(&op.win[n]), op.u8off, op.u8scale);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = (i0 + u) * CTA + threadIdx.x;
                tile[TCfg::slot(i >> LOG2TC, i & (TC - 1))] = v[u];
            }
        }
        __syncthreads();
        // 2. column transforms, in place in the tile
#pragma unroll 1
        for (int r = 0; r < TCfg::ROUNDS; ++r) {
            const int c = r * TEAMS + team;
            cd w = __ldg(&op.twM[(n2_0 + c) * tid]);
            const cd wstep = __ldg(&op.twM[(n2_0 + c) * NT]);
            cd b[P];
#pragma unroll
            for (int m = 0; m < P; ++m) b[m] = tile[TCfg::slot(tid + NT * m, c)];
            butterflies<double, P, (1 << L0), false>(b, nullptr);
            fft_tail<double, LOG2L, LOG2P, false, false, L0, 0, 0>(b, nullptr, tw, xbuf, xbuf, tid, sync);
            // bin k = tid + NT*m of column n2 leaves multiplied by W_M^(n2*k) = W_M^(n2*tid) * (W_M^(n2*NT))^m: two table
            // entries per thread and column (fetched before the transform) instead of one gather per element
#pragma unroll
            for (int m = 0; m < P; ++m) {
                b[m] = cmul(b[m], w);
                w = cmul(w, wstep);
            }
#pragma unroll
            for (int m = 0; m < P; ++m) tile[TCfg::slot(tid + NT * m, c)] = b[m];
        }
        __syncthreads();
        // 3. rows of the tile to Z
        cd* zt = op.Z + fs * op.g.M + n2_0;
#pragma unroll 8
        for (int i = threadIdx.x; i < TCfg::TILE_ELEMS; i += CTA) {
            const int k = i >> LOG2TC, c = i & (TC - 1);
            zt[((int64_t)k << l2) + c] = tile[TCfg::slot(k, c)];
        }
        __syncthreads();            // the next tile's rows overwrite this one
    }
}

template <int LOG2L, int INFMT>
static int launch_cols_tiled(const OpColsIn<INFMT, false>& op, const cd* tw, int64_t nFrameSlabs, int smCount, cudaStream_t st) {
    using TCfg = ColsTileCfg<LOG2L>;
    auto k = cols_tiled_kernel<LOG2L, INFMT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int64_t nTiles = nFrameSlabs << (op.g.l2 - TCfg::LOG2TC);
    const int64_t cap = (int64_t)smCount * 2;
    const int grid = (int)(nTiles < cap ? nTiles : cap);
    k<<<grid, SmemCfg<double, LOG2L>::CTA, TCfg::SMEM_BYTES, st>>>(op, tw, nTiles);
    return (int)cudaGetLastError();
}

constexpr int COLS_TILED_MIN_L = 8, COLS_TILED_MAX_L = 10;     // column lengths 256..1024: tiles of 16..4 columns

// final row pass: a team owns row k1 of one scan and walks the scan's frames, |X| (scaled) cumulated in the registers
// that own the bins (data_cumu, K:124-147); one store per bin and scan.  acc layout: [k1][k2] (four-step; the epilogue
// un-permutes) or natural bin order (Bluestein, only bins < F exist).
struct RowsAccParams {
    BigGeom g; const cd* Z; double* acc; double scale; int cumuMode; int nFrames; int transposedAcc;
};

// OCC3: 128-thread CTAs three per SM (168 registers, no spills; two exchange buffers of 35 KB each) instead of two (204 registers)
template <int LOG2L, bool OCC3>
__global__ void __launch_bounds__(SmemCfg<double, LOG2L>::CTA, (OCC3 && SmemCfg<double, LOG2L>::CTA == 128 ? 3 : 1))
team_fft_acc_kernel(const RowsAccParams p, const cd* __restrict__ tw, int64_t nBatch) {
    using C = SmemCfg<double, LOG2L>;
    constexpr int P = C::P, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P;
    constexpr int L0 = stage_l<LOG2L, LOG2P>(0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    cd* bufA = reinterpret_cast<cd*>(smem_raw) + team * C::FPAD;
    cd* bufB = C::DBUF ? bufA + TEAMS * C::FPAD : bufA;
    auto sync = [] { __syncthreads(); };
    const int64_t perIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (nBatch + perIter - 1) / perIter;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t q = it * perIter + (int64_t)blockIdx.x * TEAMS + team;
        const bool valid = q < nBatch;
        const int64_t qc = valid ? q : nBatch - 1;
        const int64_t s = qc >> p.g.l1, k1 = qc & (((int64_t)1 << p.g.l1) - 1);
        double a[P];
        for (int f = 0; f < p.nFrames; ++f) {
            const cd* row = p.Z + (s * p.nFrames + f) * p.g.M + (k1 << p.g.l2);
            cd b[P];
#pragma unroll
            for (int m = 0; m < P; ++m) b[m] = row[tid + NT * m];
            butterflies<double, P, (1 << L0), false>(b, nullptr);
            fft_tail<double, LOG2L, LOG2P, false, C::DBUF, L0, 0, 0>(b, nullptr, tw, bufA, bufB, tid, sync);
            if constexpr (C::DBUF && (C::NX & 1)) { cd* t = bufA; bufA = bufB; bufB = t; }
#pragma unroll
            for (int m = 0; m < P; ++m) {
                double mag = sqrt(b[m].x * b[m].x + b[m].y * b[m].y) * p.scale;
                if (p.cumuMode == KSPEC_CUMU_PSD) mag *= mag;                 // bUsePSD: power, summed (K:374-384)
                if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) a[m] = mag;
                else if (p.cumuMode == KSPEC_CUMU_AVG) a[m] = (a[m] + mag) / 2;
                else if (p.cumuMode == KSPEC_CUMU_MAX) a[m] = fmax(a[m], mag);
                else if (p.cumuMode == KSPEC_CUMU_MIN) a[m] = fmin(a[m], mag);
                else a[m] += mag;
            }
        }
        if (valid) {
#pragma unroll
            for (int m = 0; m < P; ++m) {
                const int k = tid + NT * m;
                const int64_t bin = k1 + ((int64_t)k << p.g.l1);
                if (bin < p.g.F) p.acc[s * p.g.F + (p.transposedAcc ? (k1 << p.g.l2) + k : bin)] = a[m];
            }
        }
    }
}

template <int LOG2L, bool OCC3>
static int launch_team_fft_acc(const RowsAccParams& p, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st) {
    using C = SmemCfg<double, LOG2L>;
    auto k = team_fft_acc_kernel<LOG2L, OCC3>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, C::CTA, C::SMEM_BYTES) != cudaSuccess || occ < 1) occ = 1;
    const int64_t need = (nBatch + C::TEAMS - 1) / C::TEAMS;
    const int64_t cap = (int64_t)smCount * occ * (OCC3 ? 1 : 2);     // whole waves of resident CTAs
    const int grid = (int)(need < cap ? need : cap);
    k<<<grid, C::CTA, C::SMEM_BYTES, st>>>(p, tw, nBatch);
    return (int)cudaGetLastError();
}

#define KSPEC_SWITCH_L(L, LO, HI, EXPR)                       \
    switch (L) {                                              \
        case 4:  if constexpr (4  >= LO && 4  <= HI) { constexpr int LL = 4;  return EXPR; } break;  \
        case 5:  if constexpr (5  >= LO && 5  <= HI) { constexpr int LL = 5;  return EXPR; } break;  \
        case 6:  if constexpr (6  >= LO && 6  <= HI) { constexpr int LL = 6;  return EXPR; } break;  \
        case 7:  if constexpr (7  >= LO && 7  <= HI) { constexpr int LL = 7;  return EXPR; } break;  \
        case 8:  if constexpr (8  >= LO && 8  <= HI) { constexpr int LL = 8;  return EXPR; } break;  \
        case 9:  if constexpr (9  >= LO && 9  <= HI) { constexpr int LL = 9;  return EXPR; } break;  \
        case 10: if constexpr (10 >= LO && 10 <= HI) { constexpr int LL = 10; return EXPR; } break;  \
        case 11: if constexpr (11 >= LO && 11 <= HI) { constexpr int LL = 11; return EXPR; } break;  \
        case 12: if constexpr (12 >= LO && 12 <= HI) { constexpr int LL = 12; return EXPR; } break;  \
        case 13: if constexpr (13 >= LO && 13 <= HI) { constexpr int LL = 13; return EXPR; } break;  \
        default: break;                                       \
    }

// ---- Bluestein with M small enough for one team: both transforms, the product and |.| fused in one kernel -----------
struct BlueSmallParams {
    const void* samples; int64_t scanStride; int64_t nScans;
    const int64_t* frameOffs; int nFrames;
    const double* win; const cd* chirp; const cd* V; const cd* tw;
    int F; int cumuMode; double u8off, u8scale;
    double* acc;      // [nScans][F], natural bin order, un-normalised (the epilogue applies 2*winAdj/F)
};

template <int INFMT, int LOG2M>
__global__ void __launch_bounds__(SmemCfg<double, LOG2M>::CTA, 1)
bluestein_smem_kernel(const BlueSmallParams p) {
    using C = SmemCfg<double, LOG2M>;
    constexpr int P = C::P, NT = C::NT, TEAMS = C::TEAMS, LOG2P = C::LOG2P, M = 1 << LOG2M;
    constexpr int L0 = stage_l<LOG2M, LOG2P>(0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = (TEAMS > 1) ? (threadIdx.x / NT) : 0;
    const int tid = (TEAMS > 1) ? (threadIdx.x % NT) : threadIdx.x;
    cd* bufA = reinterpret_cast<cd*>(smem_raw) + team * C::FPAD;
    cd* bufB = C::DBUF ? bufA + TEAMS * C::FPAD : bufA;
    auto sync = [] { __syncthreads(); };
    const int64_t perIter = (int64_t)gridDim.x * TEAMS;
    const int64_t iters = (p.nScans + perIter - 1) / perIter;
    const double invM = 1.0 / (double)M;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t scan = it * perIter + (int64_t)blockIdx.x * TEAMS + team;
        const bool valid = scan < p.nScans;
        const int64_t sbase = (valid ? scan : p.nScans - 1) * p.scanStride;
        double acc[P];
        for (int f = 0; f < p.nFrames; ++f) {
            const int64_t fbase = sbase + p.frameOffs[f];
            cd b[P];
#pragma unroll
            for (int m = 0; m < P; ++m) {
                const int n = tid + NT * m;
                if (n < p.F) b[m] = cmul(Ingest<double, INFMT>::load(p.samples, fbase + n, __ldg(&p.win[n]), p.u8off, p.u8scale), __ldg(&p.chirp[n]));
                else b[m] = make_double2(0.0, 0.0);
            }
            butterflies<double, P, (1 << L0), false>(b, nullptr);
            fft_tail<double, LOG2M, LOG2P, false, C::DBUF, L0, 0, 0>(b, nullptr, p.tw, bufA, bufB, tid, sync);
#pragma unroll
            for (int m = 0; m < P; ++m) b[m] = cconj(cmul(b[m], __ldg(&p.V[tid + NT * m])));
            butterflies<double, P, (1 << L0), false>(b, nullptr);
            // the second transform continues the buffer alternation where the first one stopped
            fft_tail<double, LOG2M, LOG2P, false, C::DBUF, L0, 0, C::NX>(b, nullptr, p.tw, bufA, bufB, tid, sync);
#pragma unroll
            for (int m = 0; m < P; ++m) {
                double mag = sqrt(b[m].x * b[m].x + b[m].y * b[m].y) * invM;
                if (p.cumuMode == KSPEC_CUMU_PSD) mag *= mag;
                if (f == 0 || p.cumuMode == KSPEC_CUMU_RAW) acc[m] = mag;
                else if (p.cumuMode == KSPEC_CUMU_AVG) acc[m] = (acc[m] + mag) / 2;
                else if (p.cumuMode == KSPEC_CUMU_MAX) acc[m] = fmax(acc[m], mag);
                else if (p.cumuMode == KSPEC_CUMU_MIN) acc[m] = fmin(acc[m], mag);
                else acc[m] += mag;
            }
        }
        if (valid) {
#pragma unroll
            for (int m = 0; m < P; ++m) {
                const int k = tid + NT * m;
                if (k < p.F) p.acc[scan * p.F + k] = acc[m];
            }
        }
    }
}

template <int INFMT, int LOG2M>
static int launch_bluestein_smem(const BlueSmallParams& p, int smCount, cudaStream_t st) {
    using C = SmemCfg<double, LOG2M>;
    auto k = bluestein_smem_kernel<INFMT, LOG2M>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int64_t need = (p.nScans + C::TEAMS - 1) / C::TEAMS;
    int64_t cap = (int64_t)smCount * 2;
    int grid = (int)(need < cap ? need : cap);
    k<<<grid, C::CTA, C::SMEM_BYTES, st>>>(p);
    return (int)cudaGetLastError();
}

// entry points of the separately compiled instantiation units
int big_cols_in(int inFmt, int blue, int l1, const void* op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_cols_tiled(int inFmt, int l1, const void* op, const cd* tw, int64_t nFrameSlabs, int smCount, cudaStream_t st);
int big_cols_plain(int l1, const OpColsPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_cols_mid(int l1, const OpColsMid& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_rows_plain(int l2, const OpRowsPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_rows_mul(int l2, const OpRowsMul& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_rows_acc(int l2, int occ3, const RowsAccParams& p, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_plain(int l, const OpPlain& op, const cd* tw, int64_t nBatch, int smCount, cudaStream_t st);
int big_blue_small(int inFmt, int logM, const BlueSmallParams& p, int smCount, cudaStream_t st);

constexpr int BIG_MIN_L = 7, BIG_MAX_L = 12;      // sub-transform lengths 128..4096 -> M up to 2^24
constexpr int BLUE_SMALL_MAX_LOGM = 13;

}  // namespace kspec
