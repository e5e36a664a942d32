// kspec_internal.h — shared declarations of libkspec.so's translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/kspec.h"

namespace kspec {

// Arguments of the fused scan kernels (curscan_smem.cuh) and of the linear-row epilogue (epilogue.cu).
struct ScanParams {
    const void* samples;       // device, nScans * scanStride elements of the ingest format
    int64_t scanStride;        // elements between consecutive scans (= fullSize)
    int64_t nScans;
    const int64_t* scanBase;   // frame-parallel launches only (SMEM_VARIANT_FRAMES): first element of every one-frame "scan"
    int64_t totalElems;        // ... and the length of the sample buffer, for the clipped bulk copies
    const int32_t* frameOffs;  // device, nFrames frame start offsets inside a scan (K:386)
    int32_t nFrames;
    const void* win;           // device T[F]
    const void* tw;            // device cx<T>[F], exp(-2 pi i k / F) rounded from float64
    const void* twLin;         // device cx<T>[F - R0], the same factors in the per-stage linearised layout (fft_core.cuh)
    int32_t cumuMode;
    double linScale;           // 2 * winAdj / F  (K:391)
    double u8Offset, u8Scale;
    // per-scan epilogue
    int32_t rowsKind;          // KSPEC_ROWS_*
    void* rows;                // device T[nScans][F] or null
    int32_t dbClip;            // clip to minAmp before the log (scan mode, K:640)
    double minAmp;
    int32_t infToZero;         // +-inf -> 0 after the log (scan mode, K:641)
    double gain;
    int32_t wantStats;         // emit Max/Min partials and the last avgWin dB rows
    void* wsMax;               // device T[slots][F]
    void* wsMin;
    void* avgRows;             // device T[avgWin][F], row r = scan nScans-avgWin+r
    int32_t avgWin;
    const void* adj;           // device T[F] (shifted order) or null, subtracted for the waterfall only (K:400-411)
    int32_t hmMode;            // KSPEC_COMPRESS_*
    int32_t hmW;               // waterfall row width
    void* hm;                  // device T[nScans][hmW] or null
    // linear-row engines only: accumulation rows stored [k1][k2] with bin k = k1 + 2^accL1 * k2 (0: natural order)
    int32_t accL1, accL2;
    int32_t accShifted;        // acc rows are already fftshift-ed and normalised (zeroSpanPlay records)
    unsigned int* scanCounter; // R32 kernel: device counter (zeroed before the launch) that hands out scans beyond the first per team
    int32_t hopRing;           // R32 kernels: every frame starts fftSize/2 after the previous one and scans are 16-byte aligned
};

struct SmemKernelInfo { int ctaThreads, smemBytes, teams, ctasPerSm, stages; };

// one per (precision, ingest format); defined in smem_inst_*.cu.  info != nullptr: query only, no launch.
// variant: SMEM_VARIANT_BASE, SMEM_VARIANT_MULTI (fftSize 2048, float32: four independent teams per CTA, for large batches)
// or SMEM_VARIANT_FRAMES (the base layout reading per-"scan" bases from ScanParams::scanBase: frame-parallel small batches)
int launch_smem_f32_u8(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_smem_f32_c64(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_smem_f32_c128(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_smem_f64_u8(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_smem_f64_c64(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_smem_f64_c128(int log2F, int variant, const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);

constexpr int SMEM_VARIANT_BASE = 0, SMEM_VARIANT_MULTI = 1, SMEM_VARIANT_FRAMES = 2;
// fftSize 2048, float32, large batches: the 32 x 2 x 32 layout with one shared-memory exchange per frame (curscan_r32.cuh)
int launch_r32_u8(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_r32_c64(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
// ... and the same layout as a two-role pipeline: four teams of 2 + 2 warps (curscan_r32p.cuh)
int launch_r32p_u8(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
int launch_r32p_c64(const ScanParams& p, int grid, cudaStream_t st, SmemKernelInfo* info);
constexpr int SMEM_MAX_LOG2F_F32 = 14;
constexpr int SMEM_MAX_LOG2F_F64 = 13;
constexpr int SMEM_MIN_LOG2F = 4;
constexpr int AVG_WINDOW = 64;   // rows that can still influence the float64 halving average (2^-63 cut-off)

// ---- peer-memory exchange of the per-bin statistics (comm.cu sets it up, epilogue.cu uses it) ---------------------------------
// Every rank owns a "symmetric" buffer: two epochs x nRanks slots x 3F float64 + two epochs x nRanks flags.  stats_finish_kernel
// writes this rank's [max | min | pre-weighted avg] into slot `rank` of EVERY rank's buffer (stores over NVLink through CUDA IPC
// mappings), and its last block raises flag `rank` on every rank; peer_combine_kernel waits for all flags of the epoch and reduces
// MAX / MIN / SUM over the slots of the local buffer: the exchange is the tail of the compute kernel, no collective library call.
constexpr int KSPEC_MAX_PEERS = 8;
struct PeerExchange {
    int nRanks = 0, rank = 0, F = 0;
    double* slots[KSPEC_MAX_PEERS] = {};             // base of every rank's symmetric buffer, as mapped into THIS process
    unsigned long long* flags[KSPEC_MAX_PEERS] = {}; // flag arrays inside those buffers
    unsigned int* counter = nullptr;                 // local: blocks of stats_finish_kernel that have finished (last one raises the flags)
    int* status = nullptr;                           // local: set to 1 by peer_combine_kernel on a wait time-out
    unsigned long long seq = 0;                      // exchanges issued so far (epoch = seq & 1)
};
void launch_peer_combine(const PeerExchange& px, unsigned long long seq, double* out /*3F, local stats overwritten by the reduction*/, cudaStream_t st);

// ---- epilogue.cu -------------------------------------------------------------------------------------------------
// Max/Min over the per-team partials (+ carry), Avg recurrence over the last rows (+ carry), scaled for sharding.
void launch_stats_finish(int prec, const void* wsMax, const void* wsMin, int slots, const void* avgRows, int avgWin,
                         int F, const double* carry /*3F or null*/, int firstIsSeed, double avgScale,
                         double* out /*3F*/, cudaStream_t st, int partialsLinear = 0 /*partials hold linear amplitudes*/,
                         double gain = 0.0, const PeerExchange* px = nullptr /*also write `out` into every rank's slot*/,
                         unsigned long long seq = 0);
// T -> float64 widening of result rows
void launch_widen(int prec, const void* src, double* dst, int64_t n, cudaStream_t st);
// float64 host-side vectors -> T
void launch_narrow(int prec, const double* src, void* dst, int64_t n, cudaStream_t st);
// stepped-scan stitch + Max/Min/Avg (K:643-668)
void launch_scan_stitch(int prec, const void* dbRows, const uint8_t* stepOk, const int64_t* iStart, const int64_t* iDone,
                        int nSteps, int F, int64_t total, double failValue, int baseIsRaw, int passIndex,
                        double* cur, double* mx, double* mn, double* av, cudaStream_t st);
void launch_plot_highs(const double* x, const double* y, int64_t n, int numMarkers, double delta, int64_t* idxOut, int* nOut, cudaStream_t st);
void launch_conv_same(const double* v, int64_t n, const double* taps, int m, int edge, double* out, cudaStream_t st);
void launch_scan_stitch_partial(int prec, const void* dbRows, const uint8_t* stepOk, const int64_t* iStart, int nSteps, int stepBase,
                                int nLocal, int F, int64_t total, double failValue, double* curPartial, cudaStream_t st);
void launch_scan_stats_update(const double* cur, int64_t total, int64_t lastDone, int passIndex, double* mx, double* mn, double* av,
                              cudaStream_t st);
void launch_plotcompress(const double* y, int64_t n, int xRes, int mode, double* out, cudaStream_t st);
// epilogue for engines that deliver linear, un-shifted, un-normalised |X| accumulations per scan (big FFT paths)
void launch_linear_epilogue(int prec, const ScanParams& p, const void* acc /*T[nScans][F] natural bin order*/,
                            int F, int slots, cudaStream_t st);

void launch_frames_combine(int prec, const void* rows, void* out, int64_t nScans, int nFrames, int F, int cumuMode, cudaStream_t st);

// ---- bigfft.cu ---------------------------------------------------------------------------------------------------
struct BigFft;   // four-step power-of-two engine + Bluestein wrapper
BigFft* bigfft_create(int prec, int inFmt, int64_t F, int path, int64_t* convSize, const double* window, double u8off,
                      double u8scale, cudaStream_t st, char* err, size_t errLen);
void bigfft_destroy(BigFft*);
// acc[scan][bin] (T, natural order) <- cumulate over frames of |FFT(frame * window)|
int bigfft_run(BigFft*, const void* samples, int64_t scanStride, int64_t nScans, const int64_t* frameOffs, int nFrames,
               int cumuMode, void* acc, int64_t* launches);

int bigfft_acc_l1(const BigFft*);
int bigfft_acc_l2(const BigFft*);

// ---- mixedradix.cu ---------------------------------------------------------------------------------------------------
struct MixedRadix;   // two-pass engine for 7-smooth, non power-of-two frame lengths (2.4e6 = 2^8.3.5^5)
bool mixedradix_split(int64_t F, int* n1, int* n2);
MixedRadix* mixedradix_create(int prec, int inFmt, int64_t F, const double* window, double u8off, double u8scale, cudaStream_t st,
                              char* err, size_t errLen);
void mixedradix_destroy(MixedRadix*);
// acc[scan][bin] (float64, natural bin order)
int mixedradix_run(MixedRadix*, const void* samples, int64_t scanStride, int64_t nScans, const int64_t* frameOffs, int nFrames,
                   int cumuMode, void* acc, int64_t* launches);

// float64 (re, im) pairs of the linearised twiddle table of an F = 2^log2F transform with the kernels' stage schedule
std::vector<double> host_lin_twiddles(int log2F);

void set_error(const char* fmt, ...);

// comm.cu <- kspec_api.cu: device view of the [max | min | avg] vectors the last zeroSpan batch left in the plan
bool plan_stats_view(kspec_plan* plan, double** stats3F, int* F, cudaStream_t* st, int64_t* seq = nullptr);
// comm.cu -> kspec_api.cu: attach / detach the peer-memory exchange to a plan (sharded zeroSpan batches then leave REDUCED statistics)
int plan_attach_peer(kspec_plan* plan, PeerExchange* px);

}  // namespace kspec
