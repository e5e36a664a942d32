/*
 * kspec.h — C ABI of libkspec.so: the B200 (sm_100a) implementation of kSpecAnal's spectrum hot path.
 *
 * The reference (hanishkvc/prgs-sdr-kspecanal) has no FFI; its seams are Python functions inside
 * python/kspecanal.py ("K:" below).  Each entry point names the reference lines it replaces.  The
 * host side stays Python and binds these with ctypes (prgs-sdr-kspecanal_b200/kspec/_ffi.py,
 * INTEGRATION.md).  Plain pointers and sizes only; no CUDA, torch or numpy types.
 *
 * Conventions
 *   - every function returns 0 (KSPEC_OK) or a negative kspec_status; nothing exits or throws;
 *     kspec_last_error() gives the text of the last failure on the calling thread.
 *   - the caller owns every host buffer; the library owns device memory behind opaque handles.
 *   - host-visible results are float64 (so zeroSpanSave pickles keep dtype/shape, K:524-525);
 *     rows are C-contiguous [row][bin]; spectra are fftshift-ed (K:396).
 *   - one plan = one device, one CUDA stream; a plan is not thread safe; different plans are independent.
 *   - there is NO CPU fallback: without a CUDA device kspec_plan_create fails with KSPEC_ERR_CUDA.
 */
#ifndef KSPEC_H
#define KSPEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSPEC_VERSION 100 /* 0.1.0 */

typedef struct kspec_plan kspec_plan;
typedef struct kspec_comm kspec_comm;

typedef enum {
    KSPEC_OK = 0,
    KSPEC_ERR_ARG = -1,         /* bad argument (the reference would prg_quit, K:967-972) */
    KSPEC_ERR_CUDA = -2,        /* CUDA runtime failure / no device */
    KSPEC_ERR_NOMEM = -3,
    KSPEC_ERR_UNSUPPORTED = -4,
    KSPEC_ERR_NCCL = -5,
    KSPEC_ERR_STATE = -6        /* call order (e.g. fetch before run) */
} kspec_status;

/* curScanCumuMode, K:30-33, data_cumu K:124-147 */
enum { KSPEC_CUMU_RAW = 0, KSPEC_CUMU_AVG = 1, KSPEC_CUMU_MAX = 2, KSPEC_CUMU_MIN = 3,
       /* bUsePSD (K:374-384): the reference hands the scan to matplotlib's Welch PSD instead of its own loop:
        * segments every fftSize - int(fftSize*(1-curScanNonOverlap)) samples, mean over segments of |FFT(x*w)|^2,
        * divided by Fs*sum(w^2) with matplotlib's default Fs = 2, centred like fftshift.  float64 engines only. */
       KSPEC_CUMU_PSD = 4 };
/* IQ ingest formats: rtl_sdr raw interleaved uint8 I,Q (octave/load_rtlsdr.m:8-12), numpy complex64, numpy complex128 (K:335) */
enum { KSPEC_IN_U8_IQ = 0, KSPEC_IN_C64 = 1, KSPEC_IN_C128 = 2 };
/* arithmetic of the FFT/magnitude/cumulate chain.
 *   AUTO = F64: within 1e-8 dB of the reference's float64 on EVERY bin of every engine (the parity mode; tests hold it
 *     to that on all five BASELINE configurations).  AUTO never selects float32.
 *   F32 (power-of-two fftSize <= 16384 only; an explicit request, never a default) is the fast mode BASELINE.json's
 *     north_star describes.  Its contract, checked by tests/test_gpu_parity.py::test_f32_contract on the synthetic
 *     inputs of every BASELINE configuration where float32 is selectable:
 *       - every bin of a scan is within 3e-7 of that scan's strongest bin (absolute, linear amplitude): float32
 *         rounding noise of a transform whose largest output is the peak;
 *       - hence every bin above 2e-3 of the scan's peak (-27 dB in this program's 10*log10 convention) is within
 *         1e-3 dB, and weaker bins are NOT guaranteed to be: a caller that needs 1e-3 dB on deep nulls next to strong
 *         carriers uses AUTO;
 *       - frame count, frame offsets, bin order and the peak-bin argmax are exact;
 *       - on the BASELINE cfg-1 workload (fftSize 2048, hanning, 50 % overlap, cumulate AVG, tones + noise) every bin
 *         of every output (rows, waterfall rows, Max/Min/Avg) is within 1e-3 dB: the noise floor of that input sits
 *         above 2e-3 of the peak after 15 averaged frames. */
enum { KSPEC_PREC_AUTO = 0, KSPEC_PREC_F32 = 1, KSPEC_PREC_F64 = 2 };
/* pltCompress / pltCompressHM, K:25-29, _data_plotcompress K:168-202 (MIN: documented, unreachable in the reference) */
enum { KSPEC_COMPRESS_RAW = 0, KSPEC_COMPRESS_MAX = 1, KSPEC_COMPRESS_AVG = 2, KSPEC_COMPRESS_MIN = 3 };
/* which FFT engine a plan resolved to (kspec_plan_info) */
enum { KSPEC_PATH_SMEM = 0, KSPEC_PATH_FOURSTEP = 1, KSPEC_PATH_BLUESTEIN = 2, KSPEC_PATH_MIXEDRADIX = 3 };
/* kind of per-scan rows a batch emits */
enum { KSPEC_ROWS_NONE = 0, KSPEC_ROWS_LINEAR = 1, KSPEC_ROWS_DB = 2 };

typedef struct {
    int32_t fft_size;
    int64_t full_size;
    int32_t n_frames;        /* frames sdr_curscan transforms per scan (K:385-390) */
    int32_t precision;       /* resolved KSPEC_PREC_F32 / F64 */
    int32_t path;            /* KSPEC_PATH_* */
    int32_t in_fmt;
    int32_t device;
    int32_t sm_count;
    int32_t cta_threads;     /* smem path: threads per CTA */
    int32_t ctas_per_sm;     /* smem path: resident CTAs per SM (occupancy query) */
    int32_t smem_bytes;      /* smem path: dynamic shared memory per CTA */
    int32_t scans_per_cta;   /* smem path: scans processed side by side in one CTA (tiny fftSize) */
    int32_t tma_stages;      /* smem path: frames prefetched ahead by cp.async.bulk into shared memory (0 = direct loads) */
    int64_t conv_size;       /* Bluestein: convolution length M (power of two >= 2F-1), else 0 */
    double  win_adj;         /* F / sum(window), K:373 */
} kspec_plan_info_t;

/* ---- diagnostics -------------------------------------------------------------------------------------------- */
int kspec_version(void);
const char* kspec_last_error(void);
int kspec_device_count(int* n);

/* ---- plan: derived state of handle_args' tail (K:926-936) + sdr_curscan's per-call setup (K:368-373) ------------ */
/* window: fftSize float64 values computed by the host with numpy (np.hanning / np.kaiser(F,64) / ..., K:932-935).
 * nonOverlap: curScanNonOverlap (K:45).  Frame offsets are (int64)((double)(i*fftSize)*nonOverlap), K:386.
 * u8_offset/u8_scale: component = (byte - u8_offset) * u8_scale (pyrtlsdr: 127.5, 1/127.5); ignored for complex input. */
int kspec_plan_create(kspec_plan** out, int fftSize, int64_t fullSize, double nonOverlap, int cumuMode,
                      const double* window, int inFmt, double u8_offset, double u8_scale, int precision, int device);
int kspec_plan_destroy(kspec_plan* plan);
/* frame start offsets (for the bit-exact check).  offsets may be NULL to query *n only. */
int kspec_plan_frames(const kspec_plan* plan, int64_t* offsets, int* n);
int kspec_plan_info(const kspec_plan* plan, kspec_plan_info_t* info);

/* ---- sdr_curscan (K:351-397): fullSize samples -> float64[fftSize], linear, shifted --------------------------- */
int kspec_curscan(kspec_plan* plan, const void* samples, double* out);

/* ---- zero_span loop body (K:464-484) over nScans consecutive scans ------------------------------------------------
 * samples: nScans*fullSize elements (host).  Per scan: sdr_curscan -> 10*log10(.)-gain (no low clip, K:469) ->
 * Max/Min/Avg in the dB domain (K:471-476) -> waterfall row compress(dB - adj, hmMode) (K:478-480).
 * rowsKind/rows: optional nScans x fftSize output, KSPEC_ROWS_LINEAR = sdr_curscan outputs (what zero_span_save
 *   pickles, K:523-525), KSPEC_ROWS_DB = Fft.Cur per scan.  rows may be NULL (KSPEC_ROWS_NONE).
 * hm_rows: nScans x W, W = xRes if (hmMode != RAW and fftSize > xRes) else fftSize (K:449-457); may be NULL.
 * max/min/avg: fftSize each, in-out.  carry != 0: they hold the state after earlier scans (K:438-441 = None otherwise).
 * scanIndexBase/nScansTotal: position of this batch inside a capture sharded over several plans/GPUs; with
 *   base+nScans < nScansTotal the avg written is this shard's PARTIAL of the halving recurrence, pre-weighted with
 *   2^-(nScansTotal-base-nScans), so that a SUM over shards (kspec_comm_allreduce_stats) gives Fft.Avg.  Single
 *   plan: pass base=0, total=nScans.  Terms weighted below 2^-63 are dropped (< 1 ulp of float64). */
int kspec_zerospan_batch(kspec_plan* plan, const void* samples, int64_t nScans, double gain, const double* adj,
                         int hmMode, int xRes, int rowsKind, double* rows, double* hm_rows,
                         double* max, double* min, double* avg, int carry,
                         int64_t scanIndexBase, int64_t nScansTotal);

/* ---- zero_span loop body when the per-scan spectra already exist: zeroSpanPlay (K:547-564 feeding K:469-484) ------
 * lin_rows: nScans x fftSize float64, linear and shifted -- what sdr_curscan returns and zero_span_save pickled.
 * db_rows / hm_rows may be NULL; max/min/avg/carry as in kspec_zerospan_batch. */
int kspec_zerospan_rows_batch(kspec_plan* plan, const double* lin_rows, int64_t nScans, double gain, const double* adj,
                              int hmMode, int xRes, double* db_rows, double* hm_rows,
                              double* max, double* min, double* avg, int carry);

/* ---- _scan_range step loop (K:619-668) for one full pass --------------------------------------------------------
 * samples: nSteps*fullSize elements, step i = capture taken after tuning to startFreq + fS/2 + i*fS*R.
 * stepOk[i]==0: tune failed, the reference substitutes ones(fftSize) (K:635-639); may be NULL (all ok).
 * iStart/iDone: nSteps entries each computed by the host with the reference's float64 expressions (K:622-624);
 * cur/max/min/avg: totalEntries each, in-out (first pass: initialise per K:602-608 on the host).
 * passIndex 0 -> Avg is overwritten, otherwise halving-averaged (K:615-618). */
int kspec_scan_batch(kspec_plan* plan, const void* samples, int nSteps, const uint8_t* stepOk,
                     const int64_t* iStart, const int64_t* iDone, int64_t totalEntries,
                     double minAmp4Clip, double gain, int baseIsRaw, int passIndex,
                     double* cur, double* max, double* min, double* avg);

/* ---- the same pass with Fft.Cur/Max/Min/Avg RESIDENT ON THE DEVICE across passes (K:602-668; scan_range's loop K:719-732) -----
 * kspec_scan_batch moves four totalEntries-long float64 vectors to the device and back on every pass (21.6 M entries at
 * fftSize 2.4e6: 691 MB each way).  Here the state lives in the plan: kspec_scan_state_init uploads it once (the host
 * initialises it per K:602-608), kspec_scan_pass runs one pass -- the steps' samples cross PCIe in chunks on a copy stream
 * while the engine transforms the chunks that have arrived, then one stitch + Max/Min/Avg kernel updates the state in HBM --
 * and kspec_scan_state_fetch copies out whichever vectors the host wants to look at (NULL = skip).  kspec_scan_pass_dev
 * takes samples that are already on the device (16-byte aligned).  Arguments as in kspec_scan_batch. */
int kspec_scan_state_init(kspec_plan* plan, int64_t totalEntries, const double* cur, const double* max, const double* min,
                          const double* avg);
int kspec_scan_pass(kspec_plan* plan, const void* samples, int nSteps, const uint8_t* stepOk, const int64_t* iStart,
                    const int64_t* iDone, double minAmp4Clip, double gain, int baseIsRaw, int passIndex);
int kspec_scan_pass_dev(kspec_plan* plan, const void* d_samples, int nSteps, const uint8_t* stepOk, const int64_t* iStart,
                        const int64_t* iDone, double minAmp4Clip, double gain, int baseIsRaw, int passIndex);
int kspec_scan_state_fetch(kspec_plan* plan, double* cur, double* max, double* min, double* avg);

/* ---- the same pass sharded by frequency step over several plans / GPUs -----------------------------------
 * kspec_scan_shard: this plan holds the captures of steps [stepBase, stepBase+nStepsLocal) of nStepsTotal; iStart has
 * nStepsTotal entries (the global geometry).  curPartial (totalEntries) receives this shard's share of the stitched
 * Fft.Cur -- the halving recurrence written as a weighted sum -- so that a SUM over shards (kspec_comm_allreduce_sum)
 * equals the sequential result (bins that no step covers -- none in the reference's geometries, K:598-600 -- come out
 * as 0 in every shard: the caller keeps its previous Fft.Cur there, as the single-plan kspec_scan_batch does).
 * kspec_scan_stats_update then applies K:657-668 (Max/Min/Avg from the finished Fft.Cur on
 * the bins below lastDone = iDone of the last step; bScanRangeBaseDataIsRaw is not available in sharded mode). */
int kspec_scan_shard(kspec_plan* plan, const void* samples, int nStepsLocal, int stepBase, int nStepsTotal,
                     const uint8_t* stepOk, const int64_t* iStart, int64_t totalEntries,
                     double minAmp4Clip, double gain, double* curPartial);
int kspec_scan_stats_update(kspec_plan* plan, const double* cur, int64_t totalEntries, int64_t lastDone, int passIndex,
                            double* max, double* min, double* avg);

/* ---- _data_plotcompress (K:168-202) on a float64 vector -------------------------------------------------------- */
int kspec_plotcompress(kspec_plan* plan, const double* y, int64_t n, int xRes, int mode, double* out);

/* ---- SURVEY 8f "next" rows -------------------------------------------------------------------------------------------
 * plot_highs (K:243-272): indices of the numMarkers (<= 64) strongest points of a level curve that are at least
 * delta4Marking*(freqs[n-1]-freqs[0]) apart, strongest first; the weakest point is never marked (K:254). */
int kspec_plot_highs(kspec_plan* plan, const double* freqs, const double* levels, int64_t n, int numMarkers,
                     double delta4Marking, int64_t* idxOut, int* nOut);
/* data_proc 'Conv' / pltCompress conv (K:113-120): np.convolve(vals, taps, 'same'), then the first and last `edge`
 * (12 in the reference) points are set to the mean of the result.  taps = np.kaiser(128, 64) in the reference (K:87). */
int kspec_conv_smooth(kspec_plan* plan, const double* vals, int64_t n, const double* taps, int nTaps, int edge, double* out);

/* ---- device-resident variants (zero-copy pipelines; what bench.py times for the roofline) ---------------------
 * kspec_dev_* manage device buffers on the plan's device.  Sample buffers handed to kspec_zerospan_batch_dev must be
 * 16-byte aligned (every CUDA allocation is; an offset into one need not be: KSPEC_ERR_ARG otherwise) and, when
 * nScans*fullSize*elementSize is not a multiple of 16, readable up to the next multiple (kspec_dev_alloc pads): the fused
 * kernels stage frames with 16-byte granular bulk copies.  kspec_zerospan_batch_dev consumes samples already in HBM
 * and leaves rows / hm rows / stats in plan-owned device buffers until kspec_zerospan_fetch copies them out.
 * kspec_timer_* bracket work on the plan's stream with CUDA events. */
int kspec_dev_alloc(kspec_plan* plan, int64_t bytes, void** dptr);
int kspec_dev_free(kspec_plan* plan, void* dptr);
int kspec_dev_upload(kspec_plan* plan, void* dptr, const void* host, int64_t bytes);
int kspec_dev_download(kspec_plan* plan, void* host, const void* dptr, int64_t bytes);
int kspec_dev_fill_l2(kspec_plan* plan);                 /* writes a > L2-sized scratch buffer (bench hygiene) */
int kspec_host_alloc(int64_t bytes, void** hptr);        /* pinned host memory */
int kspec_host_free(void* hptr);
int kspec_zerospan_batch_dev(kspec_plan* plan, const void* d_samples, int64_t nScans, double gain, const double* adj,
                             int hmMode, int xRes, int rowsKind, int wantHm,
                             const double* max, const double* min, const double* avg, int carry,
                             int64_t scanIndexBase, int64_t nScansTotal);
int kspec_zerospan_fetch(kspec_plan* plan, double* rows, double* hm_rows, double* max, double* min, double* avg);
int kspec_sync(kspec_plan* plan);
int kspec_timer_start(kspec_plan* plan);
int kspec_timer_stop(kspec_plan* plan, float* ms);
/* durations (ms, CUDA events on the plan's stream) of the most recent fused scan-kernel launches, oldest first;
 * at most 64 are kept.  Synchronises the stream. */
int kspec_kernel_times(kspec_plan* plan, float* ms, int cap, int* n);
/* leave nSMs streaming multiprocessors out of the fused kernel's grid so that kernels of other streams -- the NCCL exchange
 * of the previous batch -- can run beside it (default 0) */
int kspec_plan_reserve_sms(kspec_plan* plan, int nSMs);
/* number of kernels this plan has launched since creation (bench "gpu_launches") */
int kspec_launch_count(const kspec_plan* plan, int64_t* n);

/* ---- multi-GPU: one process per GPU, only the per-bin vectors are exchanged (NCCL over NVLink) ----------------- */
int kspec_comm_unique_id(char id[128]);                                  /* rank 0 creates, ships it to the others */
int kspec_comm_init(kspec_comm** out, int nRanks, int rank, const char id[128], int device);
/* MAX on max, MIN on min, SUM on avg (pre-weighted partials of kspec_zerospan_batch); n float64 each, in place */
int kspec_comm_allreduce_stats(kspec_comm* comm, double* max, double* min, double* avg, int64_t n);
/* SUM over ranks of a float64 host vector, in place (the sharded stitch of kspec_scan_shard) */
int kspec_comm_allreduce_sum(kspec_comm* comm, double* v, int64_t n);
/* same reduction on the statistics the plan's last kspec_zerospan_batch_dev left on the device, without a host round
 * trip and ASYNCHRONOUSLY: the vectors are snapshotted onto the communicator's own stream and reduced there, so the
 * plan may start its next batch at once.  kspec_comm_join makes the plan's stream wait for the reduction and copies the
 * reduced vectors back into the plan (kspec_zerospan_fetch then returns them). */
int kspec_comm_allreduce_plan(kspec_comm* comm, kspec_plan* plan);
/* kspec_comm_join is only valid while the plan still holds the batch that was snapshotted: if the plan has run another
 * batch since (the overlapped order: batch k, allreduce_plan, batch k+1, ...), its statistics belong to batch k+1 and
 * join returns KSPEC_ERR_STATE without touching them; the reduced vectors of batch k are then read with
 * kspec_comm_fetch_reduced (host float64, n = fftSize each), which works in either order.  A reduction is pending until
 * one of the two has consumed it or another all-reduce has reused the communicator's buffer. */
int kspec_comm_join(kspec_comm* comm, kspec_plan* plan);
int kspec_comm_fetch_reduced(kspec_comm* comm, double* max, double* min, double* avg, int64_t n);
/* The exchange as the TAIL OF THE COMPUTE KERNEL instead of a collective call: kspec_comm_peer_setup (collective: every rank calls
 * it once, with its plan) gives every rank a symmetric device buffer, maps the peers' buffers through CUDA IPC (NVLink peer access)
 * and attaches the exchange to the plan.  From then on every kspec_zerospan_batch_dev call of that plan that is a shard of a larger
 * capture (nScansTotal != nScans, carry == 0) ends with its statistics kernel writing this rank's Max / Min / pre-weighted Avg
 * straight into every rank's buffer and raising a flag there, and a small second kernel waits for the flags of all ranks and
 * reduces: kspec_zerospan_fetch returns Fft.Max / Min / Avg of the WHOLE capture, no kspec_comm_allreduce_* call needed (and none may
 * be mixed in for that plan).  Every rank must run the same sequence of such batches.  A rank that never arrives trips a ~2 s
 * time-out in the waiting kernel: kspec_comm_peer_status reports it (the statistics are then not reduced).  Up to 8 ranks. */
int kspec_comm_peer_setup(kspec_comm* comm, kspec_plan* plan);
int kspec_comm_peer_status(kspec_comm* comm, int* timedOut);
int kspec_comm_finalize(kspec_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* KSPEC_H */
