// kspec_api.cu — the C ABI of libkspec.so (include/kspec.h): plans, batches, device buffers, timers.
// Host-side orchestration only; the arithmetic lives in curscan_smem.cuh, bigfft.cu and epilogue.cu.
#include "kspec_internal.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <new>

namespace kspec {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return e_ == cudaErrorMemoryAllocation ? KSPEC_ERR_NOMEM : KSPEC_ERR_CUDA;               \
        }                                                                                            \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return KSPEC_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3);      // grow-only with slack: repeated batches reuse the allocation
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return KSPEC_ERR_NOMEM;
        }
        cap = want;
        return KSPEC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace kspec

using namespace kspec;

struct kspec_plan {
    int F = 0;
    int64_t S = 0;
    double r = 0;
    int cumu = 0, inFmt = 0, prec = 0, device = 0, path = 0, log2F = -1;
    double u8off = 0, u8scale = 0, winAdj = 0, linScale = 0;
    std::vector<int64_t> offs;
    cudaStream_t st = nullptr;
    cudaStream_t stCopy = nullptr;                 // host->device chunks of a pipelined host batch
    std::vector<cudaEvent_t> evChunk;              // "chunk k has arrived"
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int smCount = 0;
    int smReserve = 0;
    SmemKernelInfo ki{};          // base variant of the fused kernel
    SmemKernelInfo kiMulti{};     // multi-team variant (ctasPerSm == 0: not available for this shape)
    SmemKernelInfo kiR32{};       // 32 x 2 x 32 layout (fftSize 2048, float32, uint8 / complex64 ingest)
    bool frameParallelOff = false; // KSPEC_FRAME_PARALLEL=0 at plan creation
    size_t chunkBytes = (size_t)256 << 20;   // pipelined host batches; KSPEC_PIPELINE_CHUNK_BYTES at plan creation
    int64_t statsSeq = 0;          // bumped by every batch that rewrites `stats` (kspec_comm_join checks it)
    kspec::PeerExchange* peer = nullptr;   // peer-memory exchange attached by kspec_comm_peer_setup (owned by the communicator)
    bool shardedHint = false;      // the current batch is a shard of a larger capture (scanIndexBase / nScansTotal)
    bool r32Off = false;          // KSPEC_NO_R32=1 at plan creation: keep the 16/16/8 layouts (A/B runs, tests)
    bool r32Pipe = false;         // KSPEC_R32_PIPE=1 at plan creation: the two-role pipeline (curscan_r32p.cuh) instead of the one-role kernel; kiR32 describes it
    int64_t convSize = 0;
    BigFft* big = nullptr;
    MixedRadix* mixed = nullptr;
    int64_t launches = 0;
    // event pairs around the most recent engine launches (bench: per-kernel duration for the roofline)
    static constexpr int KT = 64;
    cudaEvent_t kev[KT][2] = {};
    int64_t kcount = 0;
    // device tables
    int32_t* dOffs = nullptr;
    void* dWin = nullptr;
    void* dTw = nullptr;
    void* dTwLin = nullptr;
    // grow-only workspaces
    DevBuf in, rows, hm, wsMax, wsMin, avgRows, adj, adj64, carry, stats, wide, acc, l2, misc, frameRows, vbase, scanState, scanGeo, sched;
    int64_t scanTotal = 0;                         // entries of the device-resident stepped-scan state (0: none)
    std::vector<int64_t> geoStart, geoDone;        // step geometry currently on the device (kspec_scan_pass re-uploads it only when it changes)
    std::vector<uint8_t> geoOk;
    bool geoHasOk = false;
    int64_t vbaseScans = 0;                        // scans covered by the frame-parallel base table in vbase
    // what the last *_dev batch left behind (for fetch)
    int64_t lastScans = 0;
    int lastRowsKind = 0, lastW = 0;
    bool lastHm = false, haveBatch = false;
};

namespace {

constexpr size_t TAIL_PAD = 64;   // bytes a staged bulk copy may read past the last frame (rounded to 16 B)
size_t in_elem_bytes(int fmt) { return fmt == KSPEC_IN_U8_IQ ? 2 : (fmt == KSPEC_IN_C64 ? 8 : 16); }
size_t real_bytes(int prec) { return prec == KSPEC_PREC_F32 ? 4 : 8; }

int launch_smem(const kspec_plan* pl, int variant, const ScanParams& p, int grid, SmemKernelInfo* info) {
    const bool f32 = pl->prec == KSPEC_PREC_F32;
    const int l = pl->log2F;
    switch (pl->inFmt) {
        case KSPEC_IN_U8_IQ: return f32 ? launch_smem_f32_u8(l, variant, p, grid, pl->st, info) : launch_smem_f64_u8(l, variant, p, grid, pl->st, info);
        case KSPEC_IN_C64:   return f32 ? launch_smem_f32_c64(l, variant, p, grid, pl->st, info) : launch_smem_f64_c64(l, variant, p, grid, pl->st, info);
        default:             return f32 ? launch_smem_f32_c128(l, variant, p, grid, pl->st, info) : launch_smem_f64_c128(l, variant, p, grid, pl->st, info);
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

ScanParams base_params(const kspec_plan* pl, const void* dSamples, int64_t nScans) {
    ScanParams p{};
    p.samples = dSamples;
    p.scanStride = pl->S;
    p.nScans = nScans;
    p.frameOffs = pl->dOffs;
    p.nFrames = (int)pl->offs.size();
    p.win = pl->dWin;
    p.tw = pl->dTw;
    p.twLin = pl->dTwLin;
    p.cumuMode = pl->cumu;
    p.linScale = pl->linScale;
    p.u8Offset = pl->u8off;
    p.u8Scale = pl->u8scale;
    p.hmMode = KSPEC_COMPRESS_RAW;
    return p;
}

// run the FFT engine + per-scan epilogue described by p (rows/stats/hm pointers already set); slots out
int run_engine(kspec_plan* pl, ScanParams& p, int* slotsOut, int* statsLinear = nullptr) {
    if (statsLinear) *statsLinear = 0;
    const size_t rb = real_bytes(pl->prec);
    if (pl->path == KSPEC_PATH_SMEM) {
        const int nFrames = (int)pl->offs.size();
        {   // Small batches (the reference's own use: ONE scan per sdr_curscan call) leave most teams idle while each scan walks
            // its 15..71 frames on one team: launch every frame as a one-frame scan of its own and cumulate afterwards.
            const int64_t teamsAvail = (int64_t)pl->smCount * (pl->ki.ctasPerSm > 0 ? pl->ki.ctasPerSm : 1) * pl->ki.teams;
            const int64_t nv = p.nScans * nFrames;
            const bool off = pl->frameParallelOff;                 // always walk the frames of a scan on one team
            // (with statistics: up to 64 scans, because the per-scan epilogue that follows then walks the scans of a batch in order,
            // one thread per bin; rows only -- the stepped scans -- any count that leaves teams idle: 1226 steps of 71 64-point
            // frames are 87 046 one-frame scans instead of 1226 teams that each walk 71 frames in sequence)
            const bool fewScans = p.nScans <= 64 || (!p.wantStats && p.hm == nullptr && p.rowsKind != KSPEC_ROWS_NONE);
            if (!off && nFrames > 1 && fewScans && p.nScans * 2 <= teamsAvail && (size_t)nv * pl->F * rb <= ((size_t)256 << 20)) {
                int rc;
                if (pl->vbaseScans < p.nScans) {
                    int64_t cap = 64;
                    while (cap < p.nScans) cap *= 2;
                    std::vector<int64_t> tab((size_t)cap * nFrames);
                    for (int64_t s = 0; s < cap; ++s)
                        for (int f = 0; f < nFrames; ++f) tab[(size_t)s * nFrames + f] = s * pl->S + pl->offs[f];
                    if ((rc = pl->vbase.reserve(tab.size() * 8))) return rc;
                    CK(cudaMemcpyAsync(pl->vbase.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, pl->st));
                    CK(cudaStreamSynchronize(pl->st));           // tab goes out of scope
                    pl->vbaseScans = cap;
                }
                if ((rc = pl->frameRows.reserve((size_t)nv * pl->F * rb)) || (rc = pl->acc.reserve((size_t)p.nScans * pl->F * rb))) return rc;
                ScanParams v = base_params(pl, p.samples, nv);
                v.scanBase = (const int64_t*)pl->vbase.p;
                v.totalElems = p.nScans * pl->S;
                v.nFrames = 1;                                   // frameOffs[0] == 0
                v.cumuMode = pl->cumu == KSPEC_CUMU_PSD ? KSPEC_CUMU_PSD : KSPEC_CUMU_RAW;
                v.rowsKind = KSPEC_ROWS_LINEAR;
                v.rows = pl->frameRows.p;
                const int teams = pl->ki.teams;
                const int64_t need = (nv + teams - 1) / teams;
                const int64_t cap = (int64_t)pl->smCount * (pl->ki.ctasPerSm > 0 ? pl->ki.ctasPerSm : 1);
                const int grid = (int)(need < cap ? need : cap);
                const int ks = (int)(pl->kcount % kspec_plan::KT);
                cudaEventRecord(pl->kev[ks][0], pl->st);
                const int e = launch_smem(pl, SMEM_VARIANT_FRAMES, v, grid < 1 ? 1 : grid, nullptr);
                cudaEventRecord(pl->kev[ks][1], pl->st);
                pl->kcount += 1;
                if (e != 0) { set_error("frame-parallel scan kernel launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KSPEC_ERR_CUDA; }
                launch_frames_combine(pl->prec, pl->frameRows.p, pl->acc.p, p.nScans, nFrames, pl->F, pl->cumu, pl->st);
                if (p.wantStats) {
                    if ((rc = pl->wsMax.reserve((size_t)pl->F * rb)) || (rc = pl->wsMin.reserve((size_t)pl->F * rb))) return rc;
                    p.wsMax = pl->wsMax.p;
                    p.wsMin = pl->wsMin.p;
                }
                p.accL1 = p.accL2 = 0;
                p.accShifted = 1;
                p.linScale = 1.0;
                launch_linear_epilogue(pl->prec, p, pl->acc.p, pl->F, 1, pl->st);
                pl->launches += p.hm ? 4 : 3;
                *slotsOut = 1;
                return KSPEC_OK;
            }
        }
        // large batches of the headline shape run several independent teams per CTA (one CTA per SM); small ones the base layout
        const bool r32 = pl->kiR32.ctasPerSm > 0 && !pl->r32Off && nFrames <= 1024 && p.nScans < (int64_t)INT32_MAX / 2 &&
                         p.nScans >= (int64_t)2 * pl->smCount * pl->kiR32.teams;
        const bool multi = !r32 && pl->kiMulti.ctasPerSm > 0 && p.nScans >= (int64_t)2 * pl->smCount * pl->kiMulti.teams;
        const SmemKernelInfo& ki = r32 ? pl->kiR32 : (multi ? pl->kiMulti : pl->ki);
        const int variant = multi ? SMEM_VARIANT_MULTI : SMEM_VARIANT_BASE;
        const int teams = ki.teams;
        int64_t need = (p.nScans + teams - 1) / teams;
        // smReserve SMs are left to concurrent kernels (the NCCL exchange of the previous batch, kspec_comm_allreduce_plan)
        const int sms = pl->smCount - pl->smReserve > 0 ? pl->smCount - pl->smReserve : 1;
        int64_t cap = (int64_t)sms * (ki.ctasPerSm > 0 ? ki.ctasPerSm : 1);
        int grid = (int)(need < cap ? need : cap);
        if (grid < 1) grid = 1;
        const int slots = grid * teams;
        if (p.wantStats) {
            int rc;
            if ((rc = pl->wsMax.reserve((size_t)slots * pl->F * rb))) return rc;
            if ((rc = pl->wsMin.reserve((size_t)slots * pl->F * rb))) return rc;
            p.wsMax = pl->wsMax.p;
            p.wsMin = pl->wsMin.p;
        }
        const int ks = (int)(pl->kcount % kspec_plan::KT);
        cudaEventRecord(pl->kev[ks][0], pl->st);
        if (r32) {
            // ring staging: uniform hop of half a frame (50 % overlap) and 16-byte aligned scans
            bool ring = ((size_t)pl->S * in_elem_bytes(pl->inFmt)) % 16 == 0 && ((uintptr_t)p.samples % 16) == 0 && !pl->r32Pipe;
            for (size_t f = 0; ring && f < pl->offs.size(); ++f) ring = pl->offs[f] == (int64_t)f * (pl->F / 2);
            p.hopRing = ring ? 1 : 0;
            // dynamic scan tickets when a static partition would leave a long tail: a shard of a multi-GPU capture (the previous
            // batch's NCCL exchange may still hold SMs when this grid starts) or a last wave that is mostly empty
            const double waves = (double)p.nScans / (double)slots;
            const bool ragged = ceil(waves) / waves > 1.05;
            p.scanCounter = nullptr;
            if (pl->shardedHint || ragged) {
                int rc2;
                if ((rc2 = pl->sched.reserve(64))) return rc2;
                CK(cudaMemsetAsync(pl->sched.p, 0, 4, pl->st));       // the scan ticket counter of this launch
                p.scanCounter = (unsigned int*)pl->sched.p;
            }
        }
        int e;
        if (r32 && pl->r32Pipe) e = pl->inFmt == KSPEC_IN_U8_IQ ? launch_r32p_u8(p, grid, pl->st, nullptr) : launch_r32p_c64(p, grid, pl->st, nullptr);
        else if (r32) e = pl->inFmt == KSPEC_IN_U8_IQ ? launch_r32_u8(p, grid, pl->st, nullptr) : launch_r32_c64(p, grid, pl->st, nullptr);
        else e = launch_smem(pl, variant, p, grid, nullptr);
        cudaEventRecord(pl->kev[ks][1], pl->st);
        pl->kcount += 1;
        if (e != 0) { set_error("scan kernel launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KSPEC_ERR_CUDA; }
        pl->launches += 1;
        *slotsOut = slots;
        if (statsLinear) *statsLinear = 1;                    // the fused batch kernels keep their Max/Min partials in the linear domain
        return KSPEC_OK;
    }
    // big engines: un-normalised accumulation rows, then a shared epilogue
    int rc;
    if ((rc = pl->acc.reserve((size_t)p.nScans * pl->F * rb))) return rc;
    if (p.wantStats) {
        if ((rc = pl->wsMax.reserve((size_t)pl->F * rb))) return rc;
        if ((rc = pl->wsMin.reserve((size_t)pl->F * rb))) return rc;
        p.wsMax = pl->wsMax.p;
        p.wsMin = pl->wsMin.p;
    }
    if (pl->mixed) {
        rc = mixedradix_run(pl->mixed, p.samples, pl->S, p.nScans, pl->offs.data(), (int)pl->offs.size(), pl->cumu, pl->acc.p, &pl->launches);
        if (rc) return rc;
        p.accL1 = p.accL2 = 0;
    } else {
        rc = bigfft_run(pl->big, p.samples, pl->S, p.nScans, pl->offs.data(), (int)pl->offs.size(), pl->cumu, pl->acc.p, &pl->launches);
        if (rc) return rc;
        p.accL1 = bigfft_acc_l1(pl->big);
        p.accL2 = bigfft_acc_l2(pl->big);
    }
    launch_linear_epilogue(pl->prec, p, pl->acc.p, pl->F, 1, pl->st);
    pl->launches += p.hm ? 2 : 1;
    *slotsOut = 1;
    return KSPEC_OK;
}

int check_plan(const kspec_plan* pl) {
    if (!pl) { set_error("null plan"); return KSPEC_ERR_ARG; }
    return KSPEC_OK;
}

int hm_width(int F, int xRes, int hmMode) { return (hmMode != KSPEC_COMPRESS_RAW && F > xRes) ? xRes : F; }

}  // namespace

namespace kspec {
std::vector<double> host_lin_twiddles(int log2F) {
    // same schedule as fft_core.cuh: radix 2^LOG2P stages, the remainder last; block per stage >= 1 laid out [t-1][k]
    const int log2P = log2F >= 7 ? 4 : (log2F >= 5 ? 3 : 2);
    std::vector<double> out;
    int lns = log2P < log2F ? log2P : log2F;
    while (lns < log2F) {
        const int l = (log2P < log2F - lns) ? log2P : (log2F - lns);
        const int R = 1 << l, NS = 1 << lns;
        for (int t = 1; t < R; ++t)
            for (int k = 0; k < NS; ++k) {
                // exact reduction of t*k/(NS*R) to an octant-symmetric angle is not needed at float64: |err| ~ 1e-16
                const double a = -2.0 * M_PI * (double)t * (double)k / ((double)NS * (double)R);
                out.push_back(cos(a));
                out.push_back(sin(a));
            }
        lns += l;
    }
    return out;
}
int plan_attach_peer(kspec_plan* pl, PeerExchange* px) {
    if (!pl) { set_error("null plan"); return KSPEC_ERR_ARG; }
    if (px && px->F != pl->F) { set_error("peer exchange was set up for fftSize %d, plan has %d", px->F, pl->F); return KSPEC_ERR_ARG; }
    pl->peer = px;
    return KSPEC_OK;
}
bool plan_stats_view(kspec_plan* pl, double** stats3F, int* F, cudaStream_t* st, int64_t* seq) {
    if (!pl || !pl->haveBatch || !pl->stats.p) return false;
    *stats3F = (double*)pl->stats.p;
    *F = pl->F;
    *st = pl->st;
    if (seq) *seq = pl->statsSeq;
    return true;
}
}  // namespace kspec

extern "C" {

int kspec_version(void) { return KSPEC_VERSION; }
const char* kspec_last_error(void) { return g_err; }

int kspec_device_count(int* n) {
    if (!n) { set_error("null argument"); return KSPEC_ERR_ARG; }
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); *n = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return KSPEC_ERR_CUDA; }
    *n = c;
    return KSPEC_OK;
}

int kspec_plan_create(kspec_plan** out, int fftSize, int64_t fullSize, double nonOverlap, int cumuMode, const double* window,
                      int inFmt, double u8_offset, double u8_scale, int precision, int device) {
    if (!out || !window) { set_error("null argument"); return KSPEC_ERR_ARG; }
    *out = nullptr;
    if (fftSize < 1 || fullSize < fftSize) { set_error("fftSize %d / fullSize %lld invalid", fftSize, (long long)fullSize); return KSPEC_ERR_ARG; }
    if (!(nonOverlap > 0.0)) { set_error("curScanNonOverlap must be > 0"); return KSPEC_ERR_ARG; }
    if (cumuMode < KSPEC_CUMU_RAW || cumuMode > KSPEC_CUMU_PSD) { set_error("unknown cumuMode %d", cumuMode); return KSPEC_ERR_ARG; }
    if (inFmt < KSPEC_IN_U8_IQ || inFmt > KSPEC_IN_C128) { set_error("unknown ingest format %d", inFmt); return KSPEC_ERR_ARG; }
    if (precision < KSPEC_PREC_AUTO || precision > KSPEC_PREC_F64) { set_error("unknown precision %d", precision); return KSPEC_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libkspec has no CPU fallback");
        return KSPEC_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return KSPEC_ERR_ARG; }

    kspec_plan* pl = new (std::nothrow) kspec_plan();
    if (!pl) { set_error("out of host memory"); return KSPEC_ERR_NOMEM; }
    pl->F = fftSize; pl->S = fullSize; pl->r = nonOverlap; pl->cumu = cumuMode; pl->inFmt = inFmt; pl->device = device;
    pl->u8off = u8_offset; pl->u8scale = u8_scale;
    {   // tuning / test knobs are read once, here (none is needed in normal use: DESIGN.md 3.8)
        const char* fp = getenv("KSPEC_FRAME_PARALLEL");
        pl->frameParallelOff = fp && fp[0] == '0';
        if (const char* e = getenv("KSPEC_PIPELINE_CHUNK_BYTES")) { const long long v = atoll(e); if (v > 0) pl->chunkBytes = (size_t)v; }
    }
    const bool pow2 = (fftSize & (fftSize - 1)) == 0;
    if (pow2) { int l = 0; while ((1 << l) < fftSize) ++l; pl->log2F = l; }
    if (precision == KSPEC_PREC_AUTO) precision = KSPEC_PREC_F64;     // parity first; F32 is the explicit fast mode
    if (cumuMode == KSPEC_CUMU_PSD && precision != KSPEC_PREC_F64) { delete pl; set_error("bUsePSD (KSPEC_CUMU_PSD) runs on the float64 engines only"); return KSPEC_ERR_ARG; }
    pl->prec = precision;
    const int smemMax = precision == KSPEC_PREC_F32 ? SMEM_MAX_LOG2F_F32 : SMEM_MAX_LOG2F_F64;
    if (pow2 && pl->log2F >= SMEM_MIN_LOG2F && pl->log2F <= smemMax) pl->path = KSPEC_PATH_SMEM;
    else if (pow2 && pl->log2F > smemMax) pl->path = KSPEC_PATH_FOURSTEP;
    else pl->path = KSPEC_PATH_BLUESTEIN;
    {   // 7-smooth lengths beyond the one-kernel Bluestein range: a direct two-pass mixed-radix transform (what numpy does at
        // K:391 for such sizes).  KSPEC_FORCE_BLUESTEIN=1 keeps the chirp-z engine (BASELINE cfg-5 names it).
        int n1 = 0, n2 = 0;
        const char* fb = getenv("KSPEC_FORCE_BLUESTEIN");
        if (pl->path == KSPEC_PATH_BLUESTEIN && fftSize > 4096 && !(fb && fb[0] == '1') && mixedradix_split(fftSize, &n1, &n2))
            pl->path = KSPEC_PATH_MIXEDRADIX;
        const char* fm = getenv("KSPEC_FORCE_MIXED");
        if (pl->path == KSPEC_PATH_FOURSTEP && fm && fm[0] == '1' && mixedradix_split(fftSize, &n1, &n2)) pl->path = KSPEC_PATH_MIXEDRADIX;
    }

    // frame offsets: int(i*F*r) with the product evaluated left to right in float64 (K:368, K:386-390)
    if (cumuMode == KSPEC_CUMU_PSD) {
        // K:375, K:381: noverlap = fftSize*(1-curScanNonOverlap) (float64), truncated by matplotlib's segmenting; a segment
        // every fftSize - noverlap samples, (fullSize - noverlap) // step segments
        const int64_t nover = (int64_t)((double)fftSize * (1.0 - nonOverlap));
        const int64_t step = (int64_t)fftSize - nover;
        if (nover < 0 || step < 1) { delete pl; set_error("bUsePSD: curScanNonOverlap %g leaves no segment step", nonOverlap); return KSPEC_ERR_ARG; }
        const int64_t nSeg = (fullSize - nover) / step;
        for (int64_t i = 0; i < nSeg; ++i) pl->offs.push_back(i * step);
    } else {
        const int64_t nLoops = (int64_t)((double)fullSize / ((double)fftSize * nonOverlap));
        for (int64_t i = 0; i < nLoops; ++i) {
            const int64_t start = (int64_t)(((double)i * (double)fftSize) * nonOverlap);
            if (start + fftSize > fullSize) break;
            pl->offs.push_back(start);
        }
    }
    if (pl->offs.empty()) { delete pl; set_error("no complete frame fits (fullSize %lld, fftSize %d, nonOverlap %g)", (long long)fullSize, fftSize, nonOverlap); return KSPEC_ERR_ARG; }
    double sum = 0.0;                                        // np.sum is pairwise; plain summation differs by < 1e-13 relative
    {   // pairwise summation to stay within an ulp or two of numpy's np.sum (K:373)
        std::vector<double> t(window, window + fftSize);
        size_t n = t.size();
        while (n > 1) { size_t h = n / 2; for (size_t i = 0; i < h; ++i) t[i] = t[2 * i] + t[2 * i + 1]; if (n & 1) { t[h] = t[n - 1]; n = h + 1; } else n = h; }
        sum = t[0];
    }
    pl->winAdj = (double)fftSize / sum;
    pl->linScale = pl->winAdj * 2.0 / (double)fftSize;
    if (cumuMode == KSPEC_CUMU_PSD) {
        // mean over the segments of |X|^2 / (Fs * sum(w^2)), Fs = 2 (matplotlib default; the reference passes none, K:381)
        std::vector<double> t(fftSize);
        for (int i = 0; i < fftSize; ++i) t[i] = window[i] * window[i];
        size_t n = t.size();
        while (n > 1) { size_t h = n / 2; for (size_t i = 0; i < h; ++i) t[i] = t[2 * i] + t[2 * i + 1]; if (n & 1) { t[h] = t[n - 1]; n = h + 1; } else n = h; }
        pl->linScale = 1.0 / ((double)pl->offs.size() * 2.0 * t[0]);
    }

    DeviceGuard guard(device);
    auto fail = [&](int rc) { kspec_plan_destroy(pl); return rc; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return fail(KSPEC_ERR_CUDA); }
    pl->smCount = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&pl->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&pl->ev0) != cudaSuccess ||
        cudaEventCreate(&pl->ev1) != cudaSuccess) {
        set_error("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(KSPEC_ERR_CUDA);
    }
    for (int i = 0; i < kspec_plan::KT; ++i)
        if (cudaEventCreate(&pl->kev[i][0]) != cudaSuccess || cudaEventCreate(&pl->kev[i][1]) != cudaSuccess) {
            set_error("event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(KSPEC_ERR_CUDA);
        }
    const size_t rb = real_bytes(precision);
    if (pl->path == KSPEC_PATH_SMEM) {
        if (fullSize > (int64_t)INT32_MAX) { set_error("fullSize %lld: the fused kernels keep frame offsets in 32 bits", (long long)fullSize); return fail(KSPEC_ERR_UNSUPPORTED); }
        std::vector<int32_t> o32(pl->offs.begin(), pl->offs.end());
        if (cudaMalloc(&pl->dOffs, o32.size() * 4) != cudaSuccess || cudaMalloc(&pl->dWin, fftSize * rb) != cudaSuccess ||
            cudaMalloc(&pl->dTw, (size_t)fftSize * 2 * rb) != cudaSuccess) {
            set_error("table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(KSPEC_ERR_NOMEM);
        }
        cudaMemcpy(pl->dOffs, o32.data(), o32.size() * 4, cudaMemcpyHostToDevice);
        // twiddles exp(-2 pi i k / F): computed in float64 with exact octant symmetry, rounded once to T
        std::vector<double> tw(2 * (size_t)fftSize);
        for (int k = 0; k < fftSize; ++k) {
            const double a = -2.0 * M_PI * (double)k / (double)fftSize;
            tw[2 * k] = cos(a);
            tw[2 * k + 1] = sin(a);
        }
        if (fftSize >= 4) { tw[2 * (fftSize / 4)] = 0.0; tw[2 * (fftSize / 4) + 1] = -1.0; tw[2 * (fftSize / 2)] = -1.0; tw[2 * (fftSize / 2) + 1] = 0.0;
                            tw[2 * (3 * fftSize / 4)] = 0.0; tw[2 * (3 * fftSize / 4) + 1] = 1.0; }
        const std::vector<double> lin = host_lin_twiddles(pl->log2F);
        if (cudaMalloc(&pl->dTwLin, (lin.size() + 2) * rb) != cudaSuccess) { set_error("table allocation failed"); return fail(KSPEC_ERR_NOMEM); }
        if (precision == KSPEC_PREC_F32) {
            std::vector<float> l32(lin.begin(), lin.end());
            cudaMemcpy(pl->dTwLin, l32.data(), l32.size() * 4, cudaMemcpyHostToDevice);
        } else if (!lin.empty()) {
            cudaMemcpy(pl->dTwLin, lin.data(), lin.size() * 8, cudaMemcpyHostToDevice);
        }
        if (precision == KSPEC_PREC_F32) {
            std::vector<float> w32(fftSize), t32(2 * (size_t)fftSize);
            for (int i = 0; i < fftSize; ++i) w32[i] = (float)window[i];
            for (size_t i = 0; i < t32.size(); ++i) t32[i] = (float)tw[i];
            cudaMemcpy(pl->dWin, w32.data(), fftSize * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(pl->dTw, t32.data(), t32.size() * 4, cudaMemcpyHostToDevice);
        } else {
            cudaMemcpy(pl->dWin, window, fftSize * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(pl->dTw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
        }
        ScanParams dummy{};
        if (launch_smem(pl, SMEM_VARIANT_MULTI, dummy, 0, &pl->kiMulti) != 0) { cudaGetLastError(); pl->kiMulti = SmemKernelInfo{}; }
        if (launch_smem(pl, SMEM_VARIANT_BASE, dummy, 0, &pl->ki) != 0) { set_error("kernel attribute query failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(KSPEC_ERR_CUDA); }
        if (pl->ki.ctasPerSm < 1) { set_error("fused kernel for fftSize %d does not fit on this device", fftSize); return fail(KSPEC_ERR_UNSUPPORTED); }
        if (precision == KSPEC_PREC_F32 && pl->log2F == 11 && inFmt != KSPEC_IN_C128) {
            const char* no = getenv("KSPEC_NO_R32");
            const char* pipe = getenv("KSPEC_R32_PIPE");
            pl->r32Off = no && no[0] == '1';
            pl->r32Pipe = pipe && pipe[0] == '1';
            int e = pl->r32Pipe ? (inFmt == KSPEC_IN_U8_IQ ? launch_r32p_u8(dummy, 0, pl->st, &pl->kiR32) : launch_r32p_c64(dummy, 0, pl->st, &pl->kiR32)) : 1;
            if (e != 0 || pl->kiR32.ctasPerSm < 1) {
                cudaGetLastError();
                pl->r32Pipe = false;
                e = inFmt == KSPEC_IN_U8_IQ ? launch_r32_u8(dummy, 0, pl->st, &pl->kiR32) : launch_r32_c64(dummy, 0, pl->st, &pl->kiR32);
                if (e != 0) { cudaGetLastError(); pl->kiR32 = SmemKernelInfo{}; }
            }
        }
    } else if (pl->path == KSPEC_PATH_MIXEDRADIX) {
        char err[256] = "";
        pl->mixed = mixedradix_create(precision, inFmt, fftSize, window, u8_offset, u8_scale, pl->st, err, sizeof(err));
        if (!pl->mixed) { set_error("mixed-radix engine: %s", err); return fail(KSPEC_ERR_UNSUPPORTED); }
    } else {
        char err[256] = "";
        pl->big = bigfft_create(precision, inFmt, fftSize, pl->path, &pl->convSize, window, u8_offset, u8_scale, pl->st, err, sizeof(err));
        if (!pl->big) { set_error("big-FFT engine: %s", err); return fail(KSPEC_ERR_UNSUPPORTED); }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("plan setup failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(KSPEC_ERR_CUDA); }
    *out = pl;
    return KSPEC_OK;
}

int kspec_plan_destroy(kspec_plan* pl) {
    if (!pl) return KSPEC_OK;
    DeviceGuard guard(pl->device);
    if (pl->st) cudaStreamSynchronize(pl->st);
    if (pl->big) bigfft_destroy(pl->big);
    if (pl->mixed) mixedradix_destroy(pl->mixed);
    for (DevBuf* b : {&pl->in, &pl->rows, &pl->hm, &pl->wsMax, &pl->wsMin, &pl->avgRows, &pl->adj, &pl->adj64, &pl->carry, &pl->stats,
                      &pl->wide, &pl->acc, &pl->l2, &pl->misc, &pl->frameRows, &pl->vbase, &pl->scanState, &pl->scanGeo, &pl->sched}) b->release();
    if (pl->dOffs) cudaFree(pl->dOffs);
    if (pl->dWin) cudaFree(pl->dWin);
    if (pl->dTw) cudaFree(pl->dTw);
    if (pl->dTwLin) cudaFree(pl->dTwLin);
    for (int i = 0; i < kspec_plan::KT; ++i) { if (pl->kev[i][0]) cudaEventDestroy(pl->kev[i][0]); if (pl->kev[i][1]) cudaEventDestroy(pl->kev[i][1]); }
    if (pl->ev0) cudaEventDestroy(pl->ev0);
    if (pl->ev1) cudaEventDestroy(pl->ev1);
    for (cudaEvent_t e : pl->evChunk) cudaEventDestroy(e);
    if (pl->stCopy) cudaStreamDestroy(pl->stCopy);
    if (pl->st) cudaStreamDestroy(pl->st);
    delete pl;
    return KSPEC_OK;
}

int kspec_plan_frames(const kspec_plan* pl, int64_t* offsets, int* n) {
    if (check_plan(pl) || !n) { set_error("null argument"); return KSPEC_ERR_ARG; }
    if (offsets) memcpy(offsets, pl->offs.data(), pl->offs.size() * sizeof(int64_t));
    *n = (int)pl->offs.size();
    return KSPEC_OK;
}

int kspec_plan_info(const kspec_plan* pl, kspec_plan_info_t* info) {
    if (check_plan(pl) || !info) { set_error("null argument"); return KSPEC_ERR_ARG; }
    memset(info, 0, sizeof(*info));
    info->fft_size = pl->F; info->full_size = pl->S; info->n_frames = (int)pl->offs.size(); info->precision = pl->prec;
    info->path = pl->path; info->in_fmt = pl->inFmt; info->device = pl->device; info->sm_count = pl->smCount;
    const SmemKernelInfo& ki = (pl->kiR32.ctasPerSm > 0 && !pl->r32Off) ? pl->kiR32 : (pl->kiMulti.ctasPerSm > 0 ? pl->kiMulti : pl->ki);     // what a large batch runs
    info->cta_threads = ki.ctaThreads; info->ctas_per_sm = ki.ctasPerSm; info->smem_bytes = ki.smemBytes;
    info->scans_per_cta = ki.teams; info->tma_stages = ki.stages; info->conv_size = pl->convSize; info->win_adj = pl->winAdj;
    return KSPEC_OK;
}

// ---- device-resident batch ---------------------------------------------------------------------------------------
namespace {

// One engine launch + stats for scans [scanOfs, scanOfs + nScans) of a batch whose samples start at dSamples (already
// offset) and whose row buffers (pl->rows / pl->hm, reserved by the caller for the whole batch) receive this part at row
// scanOfs.  dCarry: device [max|min|avg] to continue from, or nullptr.  Leaves [max|min|avg] in pl->stats.
int zerospan_part(kspec_plan* pl, const void* dSamples, int64_t nScans, int64_t scanOfs, double gain, bool haveAdj, int hmMode, int W,
                  int rowsKind, bool wantHm, const double* dCarry, int firstIsSeed, double avgScale, bool exchange = false) {
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    int rc;
    ScanParams p = base_params(pl, dSamples, nScans);
    p.gain = gain;
    p.rowsKind = rowsKind;
    if (rowsKind != KSPEC_ROWS_NONE) p.rows = (char*)pl->rows.p + (size_t)scanOfs * F * rb;
    if (wantHm) { p.hm = (char*)pl->hm.p + (size_t)scanOfs * W * rb; p.hmMode = hmMode; p.hmW = W; }
    p.wantStats = 1;
    p.avgWin = (int)(nScans < AVG_WINDOW ? nScans : AVG_WINDOW);
    if ((rc = pl->avgRows.reserve((size_t)p.avgWin * F * rb))) return rc;
    p.avgRows = pl->avgRows.p;
    if (haveAdj) p.adj = pl->adj.p;
    int slots = 0, statsLinear = 0;
    if ((rc = run_engine(pl, p, &slots, &statsLinear))) return rc;
    // a shard of a multi-GPU capture with the peer exchange attached: the statistics kernel also writes this rank's vectors
    // into every rank's symmetric buffer, and a small second kernel reduces over the ranks: `stats` then holds Fft.Max/Min/Avg
    const bool px = exchange && pl->peer && pl->peer->F == F;
    const unsigned long long seq = px ? ++pl->peer->seq : 0;
    launch_stats_finish(pl->prec, pl->wsMax.p, pl->wsMin.p, slots, pl->avgRows.p, p.avgWin, F, dCarry, firstIsSeed, avgScale,
                        (double*)pl->stats.p, pl->st, statsLinear, gain, px ? pl->peer : nullptr, seq);
    pl->launches += 1;
    if (px) { launch_peer_combine(*pl->peer, seq, (double*)pl->stats.p, pl->st); pl->launches += 1; }
    pl->statsSeq += 1;
    CK(cudaGetLastError());
    return KSPEC_OK;
}

int zerospan_check(kspec_plan* pl, const void* samples, int64_t nScans, int hmMode, int xRes, int rowsKind, int wantHm, const double* mx,
                   const double* mn, const double* av, int carry, int64_t scanIndexBase, int64_t nScansTotal, int* W) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples || nScans < 1) { set_error("no scans"); return KSPEC_ERR_ARG; }
    if (rowsKind < KSPEC_ROWS_NONE || rowsKind > KSPEC_ROWS_DB) { set_error("unknown rowsKind %d", rowsKind); return KSPEC_ERR_ARG; }
    if (hmMode < KSPEC_COMPRESS_RAW || hmMode > KSPEC_COMPRESS_MIN) { set_error("unknown pltCompressHM %d", hmMode); return KSPEC_ERR_ARG; }
    if (carry && (!mx || !mn || !av)) { set_error("carry requested without max/min/avg state"); return KSPEC_ERR_ARG; }
    if (nScansTotal < scanIndexBase + nScans || scanIndexBase < 0) { set_error("shard [%lld,+%lld) outside capture of %lld scans", (long long)scanIndexBase, (long long)nScans, (long long)nScansTotal); return KSPEC_ERR_ARG; }
    if (pl->path == KSPEC_PATH_SMEM) {
        // the fused kernels stage frames with 16-byte granular bulk copies relative to the sample base
        if (((uintptr_t)samples & 15) != 0) { set_error("device sample buffer must be 16-byte aligned (kspec_dev_alloc is)"); return KSPEC_ERR_ARG; }
    }
    *W = hm_width(pl->F, xRes, hmMode);
    if (wantHm && (xRes < 1 || pl->F % *W != 0)) { set_error("fftSize %d is not a multiple of the waterfall width %d (xRes must divide fftSize, K:941-949)", pl->F, *W); return KSPEC_ERR_ARG; }
    return KSPEC_OK;
}

// buffers and uploads shared by every part of a batch: row buffers, adj, host carry -> pl->carry
int zerospan_prepare(kspec_plan* pl, int64_t nScans, int W, int rowsKind, bool wantHm, const double* adj, const double* mx, const double* mn,
                     const double* av, int carry) {
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    int rc;
    if (rowsKind != KSPEC_ROWS_NONE && (rc = pl->rows.reserve((size_t)nScans * F * rb))) return rc;
    if (wantHm && (rc = pl->hm.reserve((size_t)nScans * W * rb))) return rc;
    if (adj) {
        if ((rc = pl->adj64.reserve((size_t)F * 8)) || (rc = pl->adj.reserve((size_t)F * rb))) return rc;
        CK(cudaMemcpyAsync(pl->adj64.p, adj, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        launch_narrow(pl->prec, (const double*)pl->adj64.p, pl->adj.p, F, pl->st);
        pl->launches += 1;
    }
    if ((rc = pl->stats.reserve((size_t)3 * F * 8)) || (rc = pl->carry.reserve((size_t)3 * F * 8))) return rc;
    if (carry) {
        CK(cudaMemcpyAsync(pl->carry.p, mx, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        CK(cudaMemcpyAsync((double*)pl->carry.p + F, mn, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        CK(cudaMemcpyAsync((double*)pl->carry.p + 2 * F, av, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
    }
    return KSPEC_OK;
}

// weight of this shard's Avg partial so that a SUM over shards equals the sequential halving recurrence
double shard_avg_scale(int64_t scanIndexBase, int64_t nScans, int64_t nScansTotal) {
    const int64_t after = nScansTotal - (scanIndexBase + nScans);
    return after == 0 ? 1.0 : ldexp(1.0, (int)(after > 2000 ? -2000 : -after));
}

}  // namespace

int kspec_zerospan_batch_dev(kspec_plan* pl, const void* dSamples, int64_t nScans, double gain, const double* adj, int hmMode,
                             int xRes, int rowsKind, int wantHm, const double* mx, const double* mn, const double* av, int carry,
                             int64_t scanIndexBase, int64_t nScansTotal) {
    int W = 0, rc;
    if ((rc = zerospan_check(pl, dSamples, nScans, hmMode, xRes, rowsKind, wantHm, mx, mn, av, carry, scanIndexBase, nScansTotal, &W))) return rc;
    DeviceGuard guard(pl->device);
    pl->shardedHint = nScansTotal != nScans;
    if ((rc = zerospan_prepare(pl, nScans, W, rowsKind, wantHm != 0, adj, mx, mn, av, carry))) return rc;
    if ((rc = zerospan_part(pl, dSamples, nScans, 0, gain, adj != nullptr, hmMode, W, rowsKind, wantHm != 0,
                            carry ? (const double*)pl->carry.p : nullptr, (!carry && scanIndexBase == 0) ? 1 : 0,
                            shard_avg_scale(scanIndexBase, nScans, nScansTotal), /*exchange=*/!carry && nScansTotal != nScans)))
        return rc;
    pl->lastScans = nScans; pl->lastRowsKind = rowsKind; pl->lastW = W; pl->lastHm = wantHm != 0; pl->haveBatch = true;
    return KSPEC_OK;
}

// zero_span loop body on spectra that already exist (zeroSpanPlay): dB, Max/Min/Avg, waterfall rows on the GPU
int kspec_zerospan_rows_batch(kspec_plan* pl, const double* linRows, int64_t nScans, double gain, const double* adj, int hmMode,
                              int xRes, double* dbRows, double* hm_rows, double* mx, double* mn, double* av, int carry) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!linRows || nScans < 1 || !mx || !mn || !av) { set_error("bad rows batch arguments"); return KSPEC_ERR_ARG; }
    if (hmMode < KSPEC_COMPRESS_RAW || hmMode > KSPEC_COMPRESS_MIN) { set_error("unknown pltCompressHM %d", hmMode); return KSPEC_ERR_ARG; }
    const int F = pl->F;
    const int W = hm_width(F, xRes, hmMode);
    if (hm_rows && (xRes < 1 || F % W != 0)) { set_error("fftSize %d is not a multiple of the waterfall width %d", F, W); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const size_t rb = real_bytes(pl->prec);
    const int64_t n = nScans * F;
    int rc;
    if ((rc = pl->wide.reserve((size_t)n * 8)) || (rc = pl->acc.reserve((size_t)n * rb))) return rc;
    CK(cudaMemcpyAsync(pl->wide.p, linRows, (size_t)n * 8, cudaMemcpyHostToDevice, pl->st));
    const void* dAcc = pl->wide.p;
    if (pl->prec == KSPEC_PREC_F32) { launch_narrow(pl->prec, (const double*)pl->wide.p, pl->acc.p, n, pl->st); pl->launches += 1; dAcc = pl->acc.p; }
    ScanParams p{};
    p.nScans = nScans; p.linScale = 1.0; p.gain = gain; p.accShifted = 1; p.wantStats = 1;
    p.rowsKind = dbRows ? KSPEC_ROWS_DB : KSPEC_ROWS_NONE;
    if (dbRows) { if ((rc = pl->rows.reserve((size_t)n * rb))) return rc; p.rows = pl->rows.p; }
    if (hm_rows) { if ((rc = pl->hm.reserve((size_t)nScans * W * rb))) return rc; p.hm = pl->hm.p; p.hmMode = hmMode; p.hmW = W; }
    p.avgWin = (int)(nScans < AVG_WINDOW ? nScans : AVG_WINDOW);
    if ((rc = pl->avgRows.reserve((size_t)p.avgWin * F * rb)) || (rc = pl->wsMax.reserve((size_t)F * rb)) ||
        (rc = pl->wsMin.reserve((size_t)F * rb)) || (rc = pl->stats.reserve((size_t)3 * F * 8)))
        return rc;
    p.avgRows = pl->avgRows.p; p.wsMax = pl->wsMax.p; p.wsMin = pl->wsMin.p;
    if (adj) {
        if ((rc = pl->adj64.reserve((size_t)F * 8)) || (rc = pl->adj.reserve((size_t)F * rb))) return rc;
        CK(cudaMemcpyAsync(pl->adj64.p, adj, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        launch_narrow(pl->prec, (const double*)pl->adj64.p, pl->adj.p, F, pl->st);
        pl->launches += 1;
        p.adj = pl->adj.p;
    }
    const double* dCarry = nullptr;
    if (carry) {
        if ((rc = pl->carry.reserve((size_t)3 * F * 8))) return rc;
        CK(cudaMemcpyAsync(pl->carry.p, mx, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        CK(cudaMemcpyAsync((double*)pl->carry.p + F, mn, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        CK(cudaMemcpyAsync((double*)pl->carry.p + 2 * F, av, (size_t)F * 8, cudaMemcpyHostToDevice, pl->st));
        dCarry = (const double*)pl->carry.p;
    }
    launch_linear_epilogue(pl->prec, p, dAcc, F, 1, pl->st);
    launch_stats_finish(pl->prec, pl->wsMax.p, pl->wsMin.p, 1, pl->avgRows.p, p.avgWin, F, dCarry, carry ? 0 : 1, 1.0, (double*)pl->stats.p, pl->st);
    pl->launches += hm_rows ? 3 : 2;
    pl->statsSeq += 1;
    CK(cudaGetLastError());
    pl->lastScans = nScans; pl->lastRowsKind = p.rowsKind; pl->lastW = W; pl->lastHm = hm_rows != nullptr; pl->haveBatch = true;
    return kspec_zerospan_fetch(pl, dbRows, hm_rows, mx, mn, av);
}

int kspec_zerospan_fetch(kspec_plan* pl, double* rows, double* hm_rows, double* mx, double* mn, double* av) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!pl->haveBatch) { set_error("kspec_zerospan_fetch before any batch"); return KSPEC_ERR_STATE; }
    DeviceGuard guard(pl->device);
    const int F = pl->F;
    int rc;
    auto copy_rows = [&](const DevBuf& src, int64_t n, double* dst) -> int {
        if (pl->prec == KSPEC_PREC_F64) {
            CK(cudaMemcpyAsync(dst, src.p, (size_t)n * 8, cudaMemcpyDeviceToHost, pl->st));
        } else {
            if ((rc = pl->wide.reserve((size_t)n * 8))) return rc;
            launch_widen(pl->prec, src.p, (double*)pl->wide.p, n, pl->st);
            pl->launches += 1;
            CK(cudaMemcpyAsync(dst, pl->wide.p, (size_t)n * 8, cudaMemcpyDeviceToHost, pl->st));
            CK(cudaStreamSynchronize(pl->st));      // `wide` is reused by the next copy
        }
        return KSPEC_OK;
    };
    if (rows) {
        if (pl->lastRowsKind == KSPEC_ROWS_NONE) { set_error("rows requested but the batch emitted none"); return KSPEC_ERR_STATE; }
        if ((rc = copy_rows(pl->rows, pl->lastScans * F, rows))) return rc;
    }
    if (hm_rows) {
        if (!pl->lastHm) { set_error("waterfall rows requested but the batch emitted none"); return KSPEC_ERR_STATE; }
        if ((rc = copy_rows(pl->hm, pl->lastScans * pl->lastW, hm_rows))) return rc;
    }
    const double* s = (const double*)pl->stats.p;
    if (mx) CK(cudaMemcpyAsync(mx, s, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    if (mn) CK(cudaMemcpyAsync(mn, s + F, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    if (av) CK(cudaMemcpyAsync(av, s + 2 * F, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------------------------
int kspec_zerospan_batch(kspec_plan* pl, const void* samples, int64_t nScans, double gain, const double* adj, int hmMode, int xRes,
                         int rowsKind, double* rows, double* hm_rows, double* mx, double* mn, double* av, int carry,
                         int64_t scanIndexBase, int64_t nScansTotal) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples || nScans < 1) { set_error("no scans"); return KSPEC_ERR_ARG; }
    if (rows == nullptr) rowsKind = KSPEC_ROWS_NONE;
    DeviceGuard guard(pl->device);
    const size_t scanBytes = (size_t)pl->S * in_elem_bytes(pl->inFmt);
    const size_t bytes = (size_t)nScans * scanBytes;
    int rc;
    if ((rc = pl->in.reserve(bytes + TAIL_PAD))) return rc;
    // Large host batches are pipelined: the samples cross PCIe in chunks on a copy stream while the engine works on the
    // chunks that have arrived (each chunk continues from the previous one's Max/Min/Avg, like consecutive batches with
    // carry) and the finished rows flow back; the exposed time is one chunk's copy plus one chunk's compute.
    int64_t chunkScans = (int64_t)(pl->chunkBytes / scanBytes);
    if (chunkScans < AVG_WINDOW) chunkScans = AVG_WINDOW;        // every part must hold the whole Avg window of its own rows
    // every chunk must start on a 16-byte boundary of the device buffer (bulk copies of the staged kernels): with an odd
    // fullSize a scan is not a whole number of 16-byte granules, so chunks are made of whole groups of scans that are
    {
        int64_t grp = 1;
        while ((grp * (int64_t)scanBytes) % 16 != 0) grp *= 2;   // scanBytes >= 2: at most 8
        chunkScans = (chunkScans + grp - 1) / grp * grp;
    }
    if (nScans < 2 * chunkScans) {
        CK(cudaMemcpyAsync(pl->in.p, samples, bytes, cudaMemcpyHostToDevice, pl->st));
        if ((rc = kspec_zerospan_batch_dev(pl, pl->in.p, nScans, gain, adj, hmMode, xRes, rowsKind, hm_rows != nullptr, mx, mn, av, carry,
                                           scanIndexBase, nScansTotal)))
            return rc;
        return kspec_zerospan_fetch(pl, rows, hm_rows, mx, mn, av);
    }
    int W = 0;
    const bool wantHm = hm_rows != nullptr;
    if ((rc = zerospan_check(pl, samples, nScans, hmMode, xRes, rowsKind, wantHm, mx, mn, av, carry, scanIndexBase, nScansTotal, &W))) return rc;
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    const int64_t nChunks = (nScans + chunkScans - 1) / chunkScans;
    if (!pl->stCopy) CK(cudaStreamCreateWithFlags(&pl->stCopy, cudaStreamNonBlocking));
    while ((int64_t)pl->evChunk.size() < nChunks + 1) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        pl->evChunk.push_back(e);
    }
    if ((rc = zerospan_prepare(pl, nScans, W, rowsKind, wantHm, adj, mx, mn, av, carry))) return rc;
    const bool f32 = pl->prec == KSPEC_PREC_F32;
    const size_t rowElems = rowsKind != KSPEC_ROWS_NONE ? (size_t)nScans * F : 0, hmElems = wantHm ? (size_t)nScans * W : 0;
    if (f32 && (rc = pl->wide.reserve((rowElems + hmElems) * 8))) return rc;
    // the copy stream may not overwrite samples an earlier batch is still reading
    CK(cudaEventRecord(pl->evChunk[nChunks], pl->st));
    CK(cudaStreamWaitEvent(pl->stCopy, pl->evChunk[nChunks], 0));
    for (int64_t c = 0; c < nChunks; ++c) {
        const int64_t s0 = c * chunkScans, ns = (nScans - s0 < chunkScans) ? nScans - s0 : chunkScans;
        CK(cudaMemcpyAsync((char*)pl->in.p + (size_t)s0 * scanBytes, (const char*)samples + (size_t)s0 * scanBytes, (size_t)ns * scanBytes,
                           cudaMemcpyHostToDevice, pl->stCopy));
        CK(cudaEventRecord(pl->evChunk[c], pl->stCopy));
    }
    for (int64_t c = 0; c < nChunks; ++c) {
        const int64_t s0 = c * chunkScans, ns = (nScans - s0 < chunkScans) ? nScans - s0 : chunkScans;
        const bool last = c == nChunks - 1;
        CK(cudaStreamWaitEvent(pl->st, pl->evChunk[c], 0));
        const bool cont = carry || c > 0;
        if (c > 0) CK(cudaMemcpyAsync(pl->carry.p, pl->stats.p, (size_t)3 * F * 8, cudaMemcpyDeviceToDevice, pl->st));
        if ((rc = zerospan_part(pl, (const char*)pl->in.p + (size_t)s0 * scanBytes, ns, s0, gain, adj != nullptr, hmMode, W, rowsKind, wantHm,
                                cont ? (const double*)pl->carry.p : nullptr, (!cont && scanIndexBase == 0) ? 1 : 0,
                                last ? shard_avg_scale(scanIndexBase, nScans, nScansTotal) : 1.0)))
            return rc;
        // rows of this chunk back to the caller
        if (rowsKind != KSPEC_ROWS_NONE) {
            const size_t o = (size_t)s0 * F, n = (size_t)ns * F;
            const void* src = (const char*)pl->rows.p + o * rb;
            if (f32) { launch_widen(pl->prec, src, (double*)pl->wide.p + o, (int64_t)n, pl->st); pl->launches += 1; src = (double*)pl->wide.p + o; }
            CK(cudaMemcpyAsync(rows + o, src, n * 8, cudaMemcpyDeviceToHost, pl->st));
        }
        if (wantHm) {
            const size_t o = (size_t)s0 * W, n = (size_t)ns * W;
            const void* src = (const char*)pl->hm.p + o * rb;
            if (f32) { launch_widen(pl->prec, src, (double*)pl->wide.p + rowElems + o, (int64_t)n, pl->st); pl->launches += 1; src = (double*)pl->wide.p + rowElems + o; }
            CK(cudaMemcpyAsync(hm_rows + o, src, n * 8, cudaMemcpyDeviceToHost, pl->st));
        }
    }
    pl->lastScans = nScans; pl->lastRowsKind = rowsKind; pl->lastW = W; pl->lastHm = wantHm; pl->haveBatch = true;
    const double* st3 = (const double*)pl->stats.p;
    if (mx) CK(cudaMemcpyAsync(mx, st3, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    if (mn) CK(cudaMemcpyAsync(mn, st3 + F, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    if (av) CK(cudaMemcpyAsync(av, st3 + 2 * F, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

int kspec_curscan(kspec_plan* pl, const void* samples, double* out) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples || !out) { set_error("null argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    const size_t bytes = (size_t)pl->S * in_elem_bytes(pl->inFmt);
    int rc;
    if ((rc = pl->in.reserve(bytes + TAIL_PAD)) || (rc = pl->rows.reserve((size_t)F * rb))) return rc;
    CK(cudaMemcpyAsync(pl->in.p, samples, bytes, cudaMemcpyHostToDevice, pl->st));
    ScanParams p = base_params(pl, pl->in.p, 1);
    p.rowsKind = KSPEC_ROWS_LINEAR;
    p.rows = pl->rows.p;
    int slots = 0;
    if ((rc = run_engine(pl, p, &slots))) return rc;
    if (pl->prec == KSPEC_PREC_F64) {
        CK(cudaMemcpyAsync(out, pl->rows.p, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    } else {
        if ((rc = pl->wide.reserve((size_t)F * 8))) return rc;
        launch_widen(pl->prec, pl->rows.p, (double*)pl->wide.p, F, pl->st);
        pl->launches += 1;
        CK(cudaMemcpyAsync(out, pl->wide.p, (size_t)F * 8, cudaMemcpyDeviceToHost, pl->st));
    }
    CK(cudaStreamSynchronize(pl->st));
    pl->haveBatch = false;
    return KSPEC_OK;
}

int kspec_scan_batch(kspec_plan* pl, const void* samples, int nSteps, const uint8_t* stepOk, const int64_t* iStart, const int64_t* iDone,
                     int64_t totalEntries, double minAmp4Clip, double gain, int baseIsRaw, int passIndex, double* cur, double* mx,
                     double* mn, double* av) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples || nSteps < 1 || !iStart || !iDone || !cur || !mx || !mn || !av || totalEntries < 1) { set_error("bad scan batch arguments"); return KSPEC_ERR_ARG; }
    for (int i = 1; i < nSteps; ++i)
        if (iStart[i] < iStart[i - 1] || iStart[i] > iStart[i - 1] + pl->F) { set_error("scanRangeNonOverlap must be in (0,1]: step %d starts at %lld after %lld", i, (long long)iStart[i], (long long)iStart[i - 1]); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    const size_t bytes = (size_t)nSteps * pl->S * in_elem_bytes(pl->inFmt);
    int rc;
    if ((rc = pl->in.reserve(bytes + TAIL_PAD)) || (rc = pl->rows.reserve((size_t)nSteps * F * rb))) return rc;
    CK(cudaMemcpyAsync(pl->in.p, samples, bytes, cudaMemcpyHostToDevice, pl->st));
    ScanParams p = base_params(pl, pl->in.p, nSteps);
    p.rowsKind = KSPEC_ROWS_DB;
    p.rows = pl->rows.p;
    p.dbClip = 1; p.minAmp = minAmp4Clip; p.infToZero = 1; p.gain = gain;
    int slots = 0;
    if ((rc = run_engine(pl, p, &slots))) return rc;
    // geometry + state vectors
    const size_t geoBytes = (size_t)nSteps * (8 + 8 + 1);
    const size_t stBytes = (size_t)totalEntries * 8;
    if ((rc = pl->misc.reserve(geoBytes + 64 + 4 * stBytes))) return rc;
    char* base = (char*)pl->misc.p;
    double* dCur = (double*)base;
    double* dMx = dCur + totalEntries; double* dMn = dMx + totalEntries; double* dAv = dMn + totalEntries;
    int64_t* dStart = (int64_t*)(dAv + totalEntries);
    int64_t* dDone = dStart + nSteps;
    uint8_t* dOk = (uint8_t*)(dDone + nSteps);
    CK(cudaMemcpyAsync(dCur, cur, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dMx, mx, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dMn, mn, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dAv, av, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dStart, iStart, (size_t)nSteps * 8, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dDone, iDone, (size_t)nSteps * 8, cudaMemcpyHostToDevice, pl->st));
    if (stepOk) CK(cudaMemcpyAsync(dOk, stepOk, (size_t)nSteps, cudaMemcpyHostToDevice, pl->st));
    // tune failure: the reference feeds ones(F) through clip + dB (K:637-641)
    double one = 1.0 > minAmp4Clip ? 1.0 : minAmp4Clip;
    double failValue = 10.0 * log10(one) - gain;
    if (isinf(failValue)) failValue = 0.0;
    launch_scan_stitch(pl->prec, pl->rows.p, stepOk ? dOk : nullptr, dStart, dDone, nSteps, F, totalEntries, failValue, baseIsRaw,
                       passIndex, dCur, dMx, dMn, dAv, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cur, dCur, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaMemcpyAsync(mx, dMx, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaMemcpyAsync(mn, dMn, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaMemcpyAsync(av, dAv, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    pl->haveBatch = false;
    return KSPEC_OK;
}

// ---- stepped scan with the state resident on the device (K:602-668 over many passes) -----------------------------------------
namespace {

int scan_geometry_check(const kspec_plan* pl, int nSteps, const int64_t* iStart, const int64_t* iDone) {
    if (nSteps < 1 || !iStart || !iDone) { set_error("bad scan geometry"); return KSPEC_ERR_ARG; }
    for (int i = 1; i < nSteps; ++i)
        if (iStart[i] < iStart[i - 1] || iStart[i] > iStart[i - 1] + pl->F) { set_error("scanRangeNonOverlap must be in (0,1]: step %d starts at %lld after %lld", i, (long long)iStart[i], (long long)iStart[i - 1]); return KSPEC_ERR_ARG; }
    return KSPEC_OK;
}

// one pass: engine over the steps (chunked and overlapped with the host->device copies when the samples are on the host),
// then the stitch + Max/Min/Avg kernel on the device-resident state
int scan_pass_common(kspec_plan* pl, const void* samples, bool onDevice, int nSteps, const uint8_t* stepOk, const int64_t* iStart,
                     const int64_t* iDone, double minAmp4Clip, double gain, int baseIsRaw, int passIndex) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples) { set_error("no samples"); return KSPEC_ERR_ARG; }
    if (pl->scanTotal < 1) { set_error("kspec_scan_pass before kspec_scan_state_init"); return KSPEC_ERR_STATE; }
    int rc;
    if ((rc = scan_geometry_check(pl, nSteps, iStart, iDone))) return rc;
    DeviceGuard guard(pl->device);
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    const size_t stepBytes = (size_t)pl->S * in_elem_bytes(pl->inFmt);
    if ((rc = pl->rows.reserve((size_t)nSteps * F * rb))) return rc;
    // geometry
    const size_t geoBytes = (size_t)nSteps * (8 + 8 + 1) + 64;
    if ((rc = pl->scanGeo.reserve(geoBytes))) return rc;
    int64_t* dStart = (int64_t*)pl->scanGeo.p;
    int64_t* dDone = dStart + nSteps;
    uint8_t* dOk = (uint8_t*)(dDone + nSteps);
    // the geometry of a stepped scan is the same on every pass (scan_range's loop, K:719-732): three small pageable uploads per
    // pass were a third of a quickFullScan pass
    const bool sameGeo = (int)pl->geoStart.size() == nSteps && memcmp(pl->geoStart.data(), iStart, (size_t)nSteps * 8) == 0 &&
                         memcmp(pl->geoDone.data(), iDone, (size_t)nSteps * 8) == 0 && pl->geoHasOk == (stepOk != nullptr) &&
                         (!stepOk || memcmp(pl->geoOk.data(), stepOk, (size_t)nSteps) == 0);
    if (!sameGeo) {
        pl->geoStart.assign(iStart, iStart + nSteps);
        pl->geoDone.assign(iDone, iDone + nSteps);
        pl->geoHasOk = stepOk != nullptr;
        if (stepOk) pl->geoOk.assign(stepOk, stepOk + nSteps); else pl->geoOk.clear();
        // the vectors in the plan outlive the asynchronous copies
        CK(cudaMemcpyAsync(dStart, pl->geoStart.data(), (size_t)nSteps * 8, cudaMemcpyHostToDevice, pl->st));
        CK(cudaMemcpyAsync(dDone, pl->geoDone.data(), (size_t)nSteps * 8, cudaMemcpyHostToDevice, pl->st));
        if (stepOk) CK(cudaMemcpyAsync(dOk, pl->geoOk.data(), (size_t)nSteps, cudaMemcpyHostToDevice, pl->st));
    }

    auto engine = [&](const void* dSamples, int first, int count) -> int {
        ScanParams p = base_params(pl, dSamples, count);
        p.rowsKind = KSPEC_ROWS_DB;
        p.rows = (char*)pl->rows.p + (size_t)first * F * rb;
        p.dbClip = 1; p.minAmp = minAmp4Clip; p.infToZero = 1; p.gain = gain;
        int slots = 0;
        return run_engine(pl, p, &slots);
    };
    if (onDevice) {
        if (pl->path == KSPEC_PATH_SMEM && ((uintptr_t)samples & 15) != 0) { set_error("device sample buffer must be 16-byte aligned"); return KSPEC_ERR_ARG; }
        if ((rc = engine(samples, 0, nSteps))) return rc;
    } else {
        if ((rc = pl->in.reserve((size_t)nSteps * stepBytes + TAIL_PAD))) return rc;
        // chunks of whole steps, each starting on a 16-byte boundary; at least four chunks when the pass is large enough to matter
        int64_t chunkSteps = (int64_t)(((size_t)64 << 20) / stepBytes);
        if (chunkSteps < 1) chunkSteps = 1;
        if ((size_t)nSteps * stepBytes < ((size_t)8 << 20)) chunkSteps = nSteps;          // small pass: one copy
        int64_t grp = 1;
        while ((grp * (int64_t)stepBytes) % 16 != 0) grp *= 2;
        chunkSteps = (chunkSteps + grp - 1) / grp * grp;
        const int64_t nChunks = (nSteps + chunkSteps - 1) / chunkSteps;
        if (nChunks == 1) {
            CK(cudaMemcpyAsync(pl->in.p, samples, (size_t)nSteps * stepBytes, cudaMemcpyHostToDevice, pl->st));
            if ((rc = engine(pl->in.p, 0, nSteps))) return rc;
        } else {
            if (!pl->stCopy) CK(cudaStreamCreateWithFlags(&pl->stCopy, cudaStreamNonBlocking));
            while ((int64_t)pl->evChunk.size() < nChunks + 1) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                pl->evChunk.push_back(e);
            }
            CK(cudaEventRecord(pl->evChunk[nChunks], pl->st));          // the copy stream may not overwrite samples still in use
            CK(cudaStreamWaitEvent(pl->stCopy, pl->evChunk[nChunks], 0));
            for (int64_t c = 0; c < nChunks; ++c) {
                const int64_t s0 = c * chunkSteps, ns = (nSteps - s0 < chunkSteps) ? nSteps - s0 : chunkSteps;
                CK(cudaMemcpyAsync((char*)pl->in.p + (size_t)s0 * stepBytes, (const char*)samples + (size_t)s0 * stepBytes, (size_t)ns * stepBytes,
                                   cudaMemcpyHostToDevice, pl->stCopy));
                CK(cudaEventRecord(pl->evChunk[c], pl->stCopy));
            }
            for (int64_t c = 0; c < nChunks; ++c) {
                const int64_t s0 = c * chunkSteps, ns = (nSteps - s0 < chunkSteps) ? nSteps - s0 : chunkSteps;
                CK(cudaStreamWaitEvent(pl->st, pl->evChunk[c], 0));
                if ((rc = engine((const char*)pl->in.p + (size_t)s0 * stepBytes, (int)s0, (int)ns))) return rc;
            }
        }
    }
    double one = 1.0 > minAmp4Clip ? 1.0 : minAmp4Clip;             // tune failure: ones(F) through clip + dB (K:637-641)
    double failValue = 10.0 * log10(one) - gain;
    if (isinf(failValue)) failValue = 0.0;
    double* dCur = (double*)pl->scanState.p;
    const int64_t T = pl->scanTotal;
    launch_scan_stitch(pl->prec, pl->rows.p, stepOk ? dOk : nullptr, dStart, dDone, nSteps, F, T, failValue, baseIsRaw, passIndex,
                       dCur, dCur + T, dCur + 2 * T, dCur + 3 * T, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    if (!onDevice) CK(cudaStreamSynchronize(pl->st));               // the caller's sample buffer is free again
    pl->haveBatch = false;
    return KSPEC_OK;
}

}  // namespace

int kspec_scan_state_init(kspec_plan* pl, int64_t totalEntries, const double* cur, const double* mx, const double* mn, const double* av) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (totalEntries < 1 || !cur || !mx || !mn || !av) { set_error("bad scan state arguments"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const size_t b = (size_t)totalEntries * 8;
    int rc;
    if ((rc = pl->scanState.reserve(4 * b))) return rc;
    double* d = (double*)pl->scanState.p;
    CK(cudaMemcpyAsync(d, cur, b, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(d + totalEntries, mx, b, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(d + 2 * totalEntries, mn, b, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(d + 3 * totalEntries, av, b, cudaMemcpyHostToDevice, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    pl->scanTotal = totalEntries;
    return KSPEC_OK;
}

int kspec_scan_state_fetch(kspec_plan* pl, double* cur, double* mx, double* mn, double* av) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (pl->scanTotal < 1) { set_error("kspec_scan_state_fetch before kspec_scan_state_init"); return KSPEC_ERR_STATE; }
    DeviceGuard guard(pl->device);
    const int64_t T = pl->scanTotal;
    const size_t b = (size_t)T * 8;
    const double* d = (const double*)pl->scanState.p;
    if (cur) CK(cudaMemcpyAsync(cur, d, b, cudaMemcpyDeviceToHost, pl->st));
    if (mx) CK(cudaMemcpyAsync(mx, d + T, b, cudaMemcpyDeviceToHost, pl->st));
    if (mn) CK(cudaMemcpyAsync(mn, d + 2 * T, b, cudaMemcpyDeviceToHost, pl->st));
    if (av) CK(cudaMemcpyAsync(av, d + 3 * T, b, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

int kspec_scan_pass(kspec_plan* pl, const void* samples, int nSteps, const uint8_t* stepOk, const int64_t* iStart, const int64_t* iDone,
                    double minAmp4Clip, double gain, int baseIsRaw, int passIndex) {
    return scan_pass_common(pl, samples, false, nSteps, stepOk, iStart, iDone, minAmp4Clip, gain, baseIsRaw, passIndex);
}

int kspec_scan_pass_dev(kspec_plan* pl, const void* dSamples, int nSteps, const uint8_t* stepOk, const int64_t* iStart, const int64_t* iDone,
                        double minAmp4Clip, double gain, int baseIsRaw, int passIndex) {
    return scan_pass_common(pl, dSamples, true, nSteps, stepOk, iStart, iDone, minAmp4Clip, gain, baseIsRaw, passIndex);
}

// ---- stepped scan sharded by frequency step -----------------------------------------------------------------
int kspec_scan_shard(kspec_plan* pl, const void* samples, int nStepsLocal, int stepBase, int nStepsTotal, const uint8_t* stepOk,
                     const int64_t* iStart, int64_t totalEntries, double minAmp4Clip, double gain, double* curPartial) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!samples || nStepsLocal < 1 || stepBase < 0 || stepBase + nStepsLocal > nStepsTotal || !iStart || !curPartial || totalEntries < 1) {
        set_error("bad scan shard arguments");
        return KSPEC_ERR_ARG;
    }
    for (int i = 1; i < nStepsTotal; ++i)
        if (iStart[i] < iStart[i - 1] || iStart[i] > iStart[i - 1] + pl->F) { set_error("scanRangeNonOverlap must be in (0,1]"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const int F = pl->F;
    const size_t rb = real_bytes(pl->prec);
    const size_t bytes = (size_t)nStepsLocal * pl->S * in_elem_bytes(pl->inFmt);
    int rc;
    if ((rc = pl->in.reserve(bytes + TAIL_PAD)) || (rc = pl->rows.reserve((size_t)nStepsLocal * F * rb))) return rc;
    CK(cudaMemcpyAsync(pl->in.p, samples, bytes, cudaMemcpyHostToDevice, pl->st));
    ScanParams p = base_params(pl, pl->in.p, nStepsLocal);
    p.rowsKind = KSPEC_ROWS_DB;
    p.rows = pl->rows.p;
    p.dbClip = 1; p.minAmp = minAmp4Clip; p.infToZero = 1; p.gain = gain;
    int slots = 0;
    if ((rc = run_engine(pl, p, &slots))) return rc;
    const size_t stBytes = (size_t)totalEntries * 8;
    if ((rc = pl->misc.reserve(stBytes + (size_t)nStepsTotal * 8 + (size_t)nStepsLocal + 64))) return rc;
    double* dCur = (double*)pl->misc.p;
    int64_t* dStart = (int64_t*)(dCur + totalEntries);
    uint8_t* dOk = (uint8_t*)(dStart + nStepsTotal);
    CK(cudaMemcpyAsync(dStart, iStart, (size_t)nStepsTotal * 8, cudaMemcpyHostToDevice, pl->st));
    if (stepOk) CK(cudaMemcpyAsync(dOk, stepOk, (size_t)nStepsLocal, cudaMemcpyHostToDevice, pl->st));
    double one = 1.0 > minAmp4Clip ? 1.0 : minAmp4Clip;
    double failValue = 10.0 * log10(one) - gain;
    if (isinf(failValue)) failValue = 0.0;
    launch_scan_stitch_partial(pl->prec, pl->rows.p, stepOk ? dOk : nullptr, dStart, nStepsTotal, stepBase, nStepsLocal, F, totalEntries,
                               failValue, dCur, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(curPartial, dCur, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    pl->haveBatch = false;
    return KSPEC_OK;
}

int kspec_scan_stats_update(kspec_plan* pl, const double* cur, int64_t totalEntries, int64_t lastDone, int passIndex, double* mx,
                            double* mn, double* av) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!cur || !mx || !mn || !av || totalEntries < 1) { set_error("bad scan stats arguments"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    const size_t stBytes = (size_t)totalEntries * 8;
    int rc;
    if ((rc = pl->misc.reserve(4 * stBytes))) return rc;
    double* dCur = (double*)pl->misc.p;
    double* dMx = dCur + totalEntries; double* dMn = dMx + totalEntries; double* dAv = dMn + totalEntries;
    CK(cudaMemcpyAsync(dCur, cur, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dMx, mx, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dMn, mn, stBytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dAv, av, stBytes, cudaMemcpyHostToDevice, pl->st));
    launch_scan_stats_update(dCur, totalEntries, lastDone, passIndex, dMx, dMn, dAv, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(mx, dMx, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaMemcpyAsync(mn, dMn, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaMemcpyAsync(av, dAv, stBytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

int kspec_plotcompress(kspec_plan* pl, const double* y, int64_t n, int xRes, int mode, double* out) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!y || !out || n < 1 || xRes < 1) { set_error("bad plotcompress arguments"); return KSPEC_ERR_ARG; }
    if (mode < KSPEC_COMPRESS_RAW || mode > KSPEC_COMPRESS_MIN) { set_error("unknown pltCompress mode %d", mode); return KSPEC_ERR_ARG; }
    const int64_t cols = n / xRes;
    if (mode == KSPEC_COMPRESS_RAW || cols == 0) { memcpy(out, y, (size_t)n * 8); return KSPEC_OK; }   // K:182-183, K:191-192: identity
    DeviceGuard guard(pl->device);
    int rc;
    if ((rc = pl->misc.reserve((size_t)(n + xRes) * 8))) return rc;
    double* dY = (double*)pl->misc.p;
    double* dO = dY + n;
    CK(cudaMemcpyAsync(dY, y, (size_t)n * 8, cudaMemcpyHostToDevice, pl->st));
    launch_plotcompress(dY, (int64_t)xRes * cols, xRes, mode, dO, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dO, (size_t)xRes * 8, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

// ---- "next" rows of SURVEY 8f: plot_highs peak picking and the Conv display mode ----------------------------------------
int kspec_plot_highs(kspec_plan* pl, const double* freqs, const double* levels, int64_t n, int numMarkers, double delta4Marking,
                     int64_t* idxOut, int* nOut) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!freqs || !levels || !idxOut || !nOut || n < 1 || numMarkers < 0 || numMarkers > 64) { set_error("bad plot_highs arguments (numMarkers <= 64)"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    int rc;
    if ((rc = pl->misc.reserve((size_t)n * 16 + 64 * 8 + 64))) return rc;
    double* dX = (double*)pl->misc.p;
    double* dY = dX + n;
    int64_t* dI = (int64_t*)(dY + n);
    int* dN = (int*)(dI + 64);
    CK(cudaMemcpyAsync(dX, freqs, (size_t)n * 8, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dY, levels, (size_t)n * 8, cudaMemcpyHostToDevice, pl->st));
    const double delta = delta4Marking * (freqs[n - 1] - freqs[0]);                // K:248-249
    launch_plot_highs(dX, dY, n, numMarkers, delta, dI, dN, pl->st);
    pl->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(nOut, dN, sizeof(int), cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    if (*nOut > 0) CK(cudaMemcpy(idxOut, dI, (size_t)*nOut * 8, cudaMemcpyDeviceToHost));
    return KSPEC_OK;
}

int kspec_conv_smooth(kspec_plan* pl, const double* vals, int64_t n, const double* taps, int nTaps, int edge, double* out) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    if (!vals || !taps || !out || n < 1 || nTaps < 1 || n < nTaps || edge < 0) { set_error("bad conv arguments (needs n >= nTaps)"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    int rc;
    if ((rc = pl->misc.reserve((size_t)(2 * n + nTaps) * 8))) return rc;
    double* dV = (double*)pl->misc.p;
    double* dO = dV + n;
    double* dT = dO + n;
    CK(cudaMemcpyAsync(dV, vals, (size_t)n * 8, cudaMemcpyHostToDevice, pl->st));
    CK(cudaMemcpyAsync(dT, taps, (size_t)nTaps * 8, cudaMemcpyHostToDevice, pl->st));
    launch_conv_same(dV, n, dT, nTaps, edge, dO, pl->st);
    pl->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dO, (size_t)n * 8, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}

// ---- device buffers, pinned memory, timers -----------------------------------------------------------------------
int kspec_dev_alloc(kspec_plan* pl, int64_t bytes, void** dptr) {
    if (check_plan(pl) || !dptr || bytes < 1) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    cudaError_t e = cudaMalloc(dptr, (size_t)bytes + TAIL_PAD);     // tail padding: see StageCfg (16-byte granular bulk copies)
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e)); return KSPEC_ERR_NOMEM; }
    return KSPEC_OK;
}
int kspec_dev_free(kspec_plan* pl, void* dptr) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    DeviceGuard guard(pl->device);
    CK(cudaStreamSynchronize(pl->st));
    CK(cudaFree(dptr));
    return KSPEC_OK;
}
int kspec_dev_upload(kspec_plan* pl, void* dptr, const void* host, int64_t bytes) {
    if (check_plan(pl) || !dptr || !host) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    CK(cudaMemcpyAsync(dptr, host, (size_t)bytes, cudaMemcpyHostToDevice, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}
int kspec_dev_download(kspec_plan* pl, void* host, const void* dptr, int64_t bytes) {
    if (check_plan(pl) || !dptr || !host) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    CK(cudaMemcpyAsync(host, dptr, (size_t)bytes, cudaMemcpyDeviceToHost, pl->st));
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}
int kspec_dev_fill_l2(kspec_plan* pl) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    DeviceGuard guard(pl->device);
    const size_t bytes = (size_t)256 << 20;      // > 126 MB L2
    int rc;
    if ((rc = pl->l2.reserve(bytes))) return rc;
    CK(cudaMemsetAsync(pl->l2.p, 0x5a, bytes, pl->st));
    return KSPEC_OK;
}
int kspec_host_alloc(int64_t bytes, void** hptr) {
    if (!hptr || bytes < 1) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    cudaError_t e = cudaHostAlloc(hptr, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e)); return KSPEC_ERR_NOMEM; }
    return KSPEC_OK;
}
int kspec_host_free(void* hptr) {
    if (hptr) cudaFreeHost(hptr);
    return KSPEC_OK;
}
int kspec_sync(kspec_plan* pl) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    DeviceGuard guard(pl->device);
    CK(cudaStreamSynchronize(pl->st));
    return KSPEC_OK;
}
int kspec_timer_start(kspec_plan* pl) {
    if (check_plan(pl)) return KSPEC_ERR_ARG;
    DeviceGuard guard(pl->device);
    CK(cudaEventRecord(pl->ev0, pl->st));
    return KSPEC_OK;
}
int kspec_timer_stop(kspec_plan* pl, float* ms) {
    if (check_plan(pl) || !ms) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    CK(cudaEventRecord(pl->ev1, pl->st));
    CK(cudaEventSynchronize(pl->ev1));
    CK(cudaEventElapsedTime(ms, pl->ev0, pl->ev1));
    return KSPEC_OK;
}
int kspec_kernel_times(kspec_plan* pl, float* ms, int cap, int* n) {
    if (check_plan(pl) || !ms || !n || cap < 1) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    DeviceGuard guard(pl->device);
    CK(cudaStreamSynchronize(pl->st));
    int64_t have = pl->kcount < kspec_plan::KT ? pl->kcount : kspec_plan::KT;
    if (have > cap) have = cap;
    for (int64_t i = 0; i < have; ++i) {
        const int ks = (int)((pl->kcount - have + i) % kspec_plan::KT);
        CK(cudaEventElapsedTime(&ms[i], pl->kev[ks][0], pl->kev[ks][1]));
    }
    *n = (int)have;
    return KSPEC_OK;
}
int kspec_plan_reserve_sms(kspec_plan* pl, int nSMs) {
    if (check_plan(pl) || nSMs < 0 || nSMs >= pl->smCount) { set_error("bad SM reserve"); return KSPEC_ERR_ARG; }
    pl->smReserve = nSMs;
    return KSPEC_OK;
}
int kspec_launch_count(const kspec_plan* pl, int64_t* n) {
    if (check_plan(pl) || !n) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    *n = pl->launches;
    return KSPEC_OK;
}

}  // extern "C"
