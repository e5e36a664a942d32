// bigfft.cu — host orchestration of the float64 multi-pass engines (kernels: bigfft_kernels.cuh).
//
//   KSPEC_PATH_FOURSTEP   power-of-two frames above the shared-memory limit (2^14 in f64 .. 2^24): per frame one column
//                         pass (fused ingest + window, twiddle) and one row pass (|X| + cumulate).
//   KSPEC_PATH_BLUESTEIN  any other frame length.  M = pow2 >= 2F-1.  M <= 8192: one fully fused kernel per batch of
//                         scans (both transforms in registers).  Larger M: four passes per frame.
//
// Replaces np.fft.fft at kspecanal.py:391 for those sizes; frame loop K:385-395; cumulate K:124-147.
#include "bigfft_kernels.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace kspec {

namespace {

__global__ void twiddle_init_kernel(cd* t, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)n, &s, &c);
        t[k] = make_double2(c, s);
    }
}

// chirp c_n = exp(-i pi n^2 / F); n^2 is reduced mod 2F in integers so the angle stays exact for F up to 2^31
__global__ void chirp_init_kernel(cd* c, int64_t F) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < F; n += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long r = ((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * F);
        double s, co;
        sincospi(-(double)r / (double)F, &s, &co);
        c[n] = make_double2(co, s);
    }
}

// v = conj(chirp) wrapped onto the M-point circle: v[n] = v[M-n] = conj(c_n), n < F; zero elsewhere
__global__ void chirp_wrap_kernel(const cd* c, cd* v, int64_t F, int64_t M) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        cd r = make_double2(0.0, 0.0);
        if (i < F) r = cconj(c[i]);
        else if (M - i < F) r = cconj(c[M - i]);
        v[i] = r;
    }
}

}  // namespace

struct BigFft {
    int inFmt = 0, path = 0, smCount = 0;
    int64_t F = 0, M = 0;
    int logM = 0, l1 = 0, l2 = 0;          // l1 == 0: M fits one team (small Bluestein)
    double u8off = 0, u8scale = 0;
    cudaStream_t st = nullptr;
    double* dWin = nullptr;
    cd *dTwM = nullptr, *dTw1 = nullptr, *dTw2 = nullptr, *dChirp = nullptr, *dV = nullptr, *dZ = nullptr, *dP = nullptr;
    int64_t* dOffs = nullptr;
    int nOffs = 0;
    bool rowsOcc3 = false;      // last row pass with three CTAs per SM
    bool tiledCols = false;     // four-step column pass through shared-memory tiles (cols_tiled_kernel)
    cd* dPw = nullptr;          // Bluestein: product slabs, one per frame of a chunk (dP is the single slab used at plan time)
    size_t zCap = 0;            // bytes allocated for dZ (and dPw)
};

static int ilog2(int64_t v) { int l = 0; while (((int64_t)1 << l) < v) ++l; return l; }

void bigfft_destroy(BigFft* b) {
    if (!b) return;
    for (void* p : {(void*)b->dWin, (void*)b->dTwM, (void*)b->dTw1, (void*)b->dTw2, (void*)b->dChirp, (void*)b->dV, (void*)b->dZ,
                    (void*)b->dP, (void*)b->dOffs, (void*)b->dPw})
        if (p) cudaFree(p);
    delete b;
}

#define BCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(err, errLen, "%s: %s", #call, cudaGetErrorString(e_));                        \
            bigfft_destroy(b);                                                                     \
            return nullptr;                                                                        \
        }                                                                                          \
    } while (0)

BigFft* bigfft_create(int prec, int inFmt, int64_t F, int path, int64_t* convSize, const double* window, double u8off,
                      double u8scale, cudaStream_t st, char* err, size_t errLen) {
    if (prec != KSPEC_PREC_F64) {
        snprintf(err, errLen, "fftSize %lld runs on the multi-pass engines, which compute in float64: use precision auto or f64", (long long)F);
        return nullptr;
    }
    BigFft* b = new BigFft();
    b->inFmt = inFmt; b->path = path; b->F = F; b->u8off = u8off; b->u8scale = u8scale; b->st = st;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&b->smCount, cudaDevAttrMultiProcessorCount, dev);
    if (path == KSPEC_PATH_FOURSTEP) {
        b->M = F;
    } else {
        int64_t m = 16;
        while (m < 2 * F - 1) m <<= 1;
        b->M = m;
    }
    b->logM = ilog2(b->M);
    if (path == KSPEC_PATH_BLUESTEIN && b->logM <= BLUE_SMALL_MAX_LOGM) {
        b->l1 = 0; b->l2 = b->logM;
    } else {
        b->l1 = b->logM / 2; b->l2 = b->logM - b->l1;
        if (b->l1 < BIG_MIN_L || b->l2 > BIG_MAX_L) {
            snprintf(err, errLen, "transform length 2^%d is outside the multi-pass range 2^%d..2^%d", b->logM, 2 * BIG_MIN_L, 2 * BIG_MAX_L);
            bigfft_destroy(b);
            return nullptr;
        }
    }
    *convSize = path == KSPEC_PATH_BLUESTEIN ? b->M : 0;
    {   // KSPEC_FOURSTEP_TILED=0 keeps the element-wise column pass (A/B measurements); read once, here
        const char* e = getenv("KSPEC_FOURSTEP_TILED");
        // last row pass: three CTAs per SM for rows of 2048 points and more (2^21: 0.414 -> 0.349 ms per 22 frames), two below
        // (2^14, 128-point rows: 13.0 vs 14.2 ms per batch with three); KSPEC_ROWS_OCC3=0/1 overrides
        const char* o = getenv("KSPEC_ROWS_OCC3");
        b->rowsOcc3 = o ? o[0] == '1' : b->l2 >= 11;
        b->tiledCols = path == KSPEC_PATH_FOURSTEP && b->l1 >= COLS_TILED_MIN_L && b->l1 <= COLS_TILED_MAX_L && !(e && e[0] == '0');
    }
    const int64_t M = b->M;
    BCK(cudaMalloc(&b->dWin, (size_t)F * 8));
    BCK(cudaMemcpyAsync(b->dWin, window, (size_t)F * 8, cudaMemcpyHostToDevice, st));
    const int64_t L2 = (int64_t)1 << b->l2;
    {   // per-team transforms read the linearised twiddle layout (fft_core.cuh)
        const std::vector<double> lin2 = host_lin_twiddles(b->l2);
        BCK(cudaMalloc(&b->dTw2, lin2.size() * 8 + 16));
        BCK(cudaMemcpyAsync(b->dTw2, lin2.data(), lin2.size() * 8, cudaMemcpyHostToDevice, st));
        BCK(cudaStreamSynchronize(st));
    }
    if (b->l1 > 0) {
        {
            const std::vector<double> lin1 = host_lin_twiddles(b->l1);
            BCK(cudaMalloc(&b->dTw1, lin1.size() * 8 + 16));
            BCK(cudaMemcpyAsync(b->dTw1, lin1.data(), lin1.size() * 8, cudaMemcpyHostToDevice, st));
            BCK(cudaStreamSynchronize(st));
        }
        BCK(cudaMalloc(&b->dTwM, (size_t)M * 16));
        twiddle_init_kernel<<<1024, 256, 0, st>>>(b->dTwM, M);
        BCK(cudaMalloc(&b->dZ, (size_t)M * 16));
        b->zCap = 0;            // (re)sized per batch in bigfft_run; this slab serves the chirp spectrum below
    }
    if (path == KSPEC_PATH_BLUESTEIN) {
        BCK(cudaMalloc(&b->dChirp, (size_t)F * 16));
        chirp_init_kernel<<<1024, 256, 0, st>>>(b->dChirp, F);
        BCK(cudaMalloc(&b->dV, (size_t)M * 16));
        BCK(cudaMalloc(&b->dP, (size_t)M * 16));
        chirp_wrap_kernel<<<1024, 256, 0, st>>>(b->dChirp, b->dP, F, M);       // v in natural order, staged in P
        BigGeom g{M, M, b->l1, b->l2};
        int e;
        if (b->l1 == 0) {
            OpPlain op{b->l2, b->dP, b->dV};
            e = big_plain(b->l2, op, b->dTw2, 1, b->smCount, st);
        } else {
            OpColsPlain oc{g, b->dP, b->dTwM, b->dZ};
            e = big_cols_plain(b->l1, oc, b->dTw1, L2, b->smCount, st);
            if (!e) {
                OpRowsPlain orr{g, b->dZ, b->dV};
                e = big_rows_plain(b->l2, orr, b->dTw2, (int64_t)1 << b->l1, b->smCount, st);
            }
        }
        if (e) { snprintf(err, errLen, "chirp spectrum launch failed: %s", cudaGetErrorString((cudaError_t)e)); bigfft_destroy(b); return nullptr; }
    }
    BCK(cudaStreamSynchronize(st));
    return b;
}

int bigfft_run(BigFft* b, const void* samples, int64_t scanStride, int64_t nScans, const int64_t* frameOffs, int nFrames,
               int cumuMode, void* acc, int64_t* launches) {
    cudaStream_t st = b->st;
    double* dAcc = reinterpret_cast<double*>(acc);
    int e = 0;
    if (b->l1 == 0) {
        // small Bluestein: everything in one launch
        if (b->nOffs != nFrames) {
            if (b->dOffs) cudaFree(b->dOffs);
            b->dOffs = nullptr;
            if (cudaMalloc(&b->dOffs, (size_t)nFrames * 8) != cudaSuccess) { set_error("frame table allocation failed"); return KSPEC_ERR_NOMEM; }
            b->nOffs = nFrames;
        }
        cudaMemcpyAsync(b->dOffs, frameOffs, (size_t)nFrames * 8, cudaMemcpyHostToDevice, st);
        BlueSmallParams p{samples, scanStride, nScans, b->dOffs, nFrames, b->dWin, b->dChirp, b->dV, b->dTw2, (int)b->F, cumuMode,
                          b->u8off, b->u8scale, dAcc};
        e = big_blue_small(b->inFmt, b->logM, p, b->smCount, st);
        *launches += 1;
        if (e) { set_error("Bluestein kernel launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KSPEC_ERR_CUDA; }
        return KSPEC_OK;
    }
    const int64_t L1 = (int64_t)1 << b->l1, L2 = (int64_t)1 << b->l2;
    BigGeom g{b->M, b->F, b->l1, b->l2};
    const bool blue = b->path == KSPEC_PATH_BLUESTEIN;
    if (b->nOffs != nFrames) {
        if (b->dOffs) cudaFree(b->dOffs);
        b->dOffs = nullptr;
        if (cudaMalloc(&b->dOffs, (size_t)nFrames * 8) != cudaSuccess) { set_error("frame table allocation failed"); return KSPEC_ERR_NOMEM; }
        b->nOffs = nFrames;
    }
    cudaMemcpyAsync(b->dOffs, frameOffs, (size_t)nFrames * 8, cudaMemcpyHostToDevice, st);
    // scans are processed in chunks whose work vectors (one M-point slab per frame) stay within ~1 GiB each
    const size_t slab = (size_t)b->M * 16;
    int64_t chunk = (int64_t)(((size_t)1 << 30) / (slab * (size_t)nFrames));
    if (chunk < 1) chunk = 1;
    if (chunk > nScans) chunk = nScans;
    const size_t need = slab * (size_t)nFrames * (size_t)chunk;
    if (b->zCap < need) {
        if (b->dZ) cudaFree(b->dZ);
        if (b->dPw) cudaFree(b->dPw);
        b->dZ = b->dPw = nullptr; b->zCap = 0;
        if (cudaMalloc(&b->dZ, need) != cudaSuccess || (blue && cudaMalloc(&b->dPw, need) != cudaSuccess)) {
            cudaGetLastError();
            set_error("multi-pass work buffers (%zu bytes) do not fit in device memory", need);
            return KSPEC_ERR_NOMEM;
        }
        b->zCap = need;
    }
    const size_t eb = b->inFmt == KSPEC_IN_U8_IQ ? 2 : (b->inFmt == KSPEC_IN_C64 ? 8 : 16);
    for (int64_t s0 = 0; s0 < nScans && !e; s0 += chunk) {
        const int64_t ns = (nScans - s0 < chunk) ? nScans - s0 : chunk;
        const int64_t nfs = ns * nFrames;
        const void* smp = reinterpret_cast<const unsigned char*>(samples) + (size_t)s0 * scanStride * eb;
        // pass 1: columns of every (zero padded) frame of the chunk
        if (b->inFmt == KSPEC_IN_U8_IQ) {
            OpColsIn<KSPEC_IN_U8_IQ, false> op{g, smp, scanStride, b->dOffs, nFrames, b->dWin, blue ? b->dChirp : nullptr, b->dTwM, b->dZ, b->u8off, b->u8scale};
            e = b->tiledCols ? big_cols_tiled(b->inFmt, b->l1, &op, b->dTw1, nfs, b->smCount, st)
                             : big_cols_in(b->inFmt, blue ? 1 : 0, b->l1, &op, b->dTw1, nfs * L2, b->smCount, st);
        } else if (b->inFmt == KSPEC_IN_C64) {
            OpColsIn<KSPEC_IN_C64, false> op{g, smp, scanStride, b->dOffs, nFrames, b->dWin, blue ? b->dChirp : nullptr, b->dTwM, b->dZ, b->u8off, b->u8scale};
            e = b->tiledCols ? big_cols_tiled(b->inFmt, b->l1, &op, b->dTw1, nfs, b->smCount, st)
                             : big_cols_in(b->inFmt, blue ? 1 : 0, b->l1, &op, b->dTw1, nfs * L2, b->smCount, st);
        } else {
            OpColsIn<KSPEC_IN_C128, false> op{g, smp, scanStride, b->dOffs, nFrames, b->dWin, blue ? b->dChirp : nullptr, b->dTwM, b->dZ, b->u8off, b->u8scale};
            e = b->tiledCols ? big_cols_tiled(b->inFmt, b->l1, &op, b->dTw1, nfs, b->smCount, st)
                             : big_cols_in(b->inFmt, blue ? 1 : 0, b->l1, &op, b->dTw1, nfs * L2, b->smCount, st);
        }
        *launches += 1;
        if (e) break;
        if (blue) {
            OpRowsMul om{g, b->dZ, b->dV, b->dPw};
            e = big_rows_mul(b->l2, om, b->dTw2, nfs * L1, b->smCount, st);
            if (e) break;
            OpColsMid oc{g, b->dPw, b->dTwM, b->dZ};
            e = big_cols_mid(b->l1, oc, b->dTw1, nfs * L2, b->smCount, st);
            if (e) break;
            *launches += 2;
        }
        // last pass: rows, |X|, cumulate over the frames of each scan in registers
        RowsAccParams ra{g, b->dZ, dAcc + s0 * b->F, blue ? 1.0 / (double)b->M : 1.0, cumuMode, nFrames, blue ? 0 : 1};
        e = big_rows_acc(b->l2, b->rowsOcc3 ? 1 : 0, ra, b->dTw2, ns * L1, b->smCount, st);
        *launches += 1;
    }
    if (e) { set_error("multi-pass FFT launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KSPEC_ERR_CUDA; }
    return KSPEC_OK;
}

// log2 of the column count when acc rows are stored [k1][k2] (four-step), 0 when they are in natural bin order
int bigfft_acc_l1(const BigFft* b) { return (b->path == KSPEC_PATH_FOURSTEP) ? b->l1 : 0; }
int bigfft_acc_l2(const BigFft* b) { return b->l2; }

}  // namespace kspec
