"""Host-side mirror of kspecanal.py's spectrum-path functions: same names, same dict ``d`` keys, same error
behaviour, but the arithmetic runs in libkspec.so on the GPU.  ``K:`` = /root/reference/python/kspecanal.py.

What stays on the host (untouched by design, BASELINE.json north_star): the rtlsdr tuning/read loop
(``sdr_setup``/``sdr_read``, K:287-347), pickle file formats (K:509-564) and all plotting.  What moves to the
GPU: everything between "samples are in RAM" and "arrays handed to matplotlib".

A maintainer wires these in by replacing the bodies of the reference functions; the
functions also work stand-alone for headless captures (``kspec.synth.ArrayRtlSdr``).
"""
import pickle
import sys
import time

import numpy as np

from . import _ffi
from .engine import Plan, heatmap_width

CUMUMODE_MAX, CUMUMODE_MIN, CUMUMODE_AVG, CUMUMODE_RAW = "MAX", "MIN", "AVG", "RAW"
PLTCOMPRESS_MAX, PLTCOMPRESS_MIN, PLTCOMPRESS_AVG, PLTCOMPRESS_RAW = "MAX", "MIN", "AVG", "RAW"
gSdrReadUnit = 2 ** 18                      # K:310
gFft2FullMult4Less, gFft2FullMult4More = 8, 2   # K:49-50
HEATMAP_ROWS = 128                          # K:448, K:611


class KspecQuit(SystemExit):
    """Raised where the reference calls prg_quit(d, msg) (K:967-972): message printed, cmd.stop set, exit."""


def prg_quit(d, msg=None, tryExit=True):
    if msg is not None:
        print(msg)
    d["cmd.stop"] = True
    if tryExit:
        raise KspecQuit()


# ------------------------------------------------------------------------------------------------------
# configuration (handle_args tail, K:771-949) without the argv parsing, which stays in kspecanal.py
# ------------------------------------------------------------------------------------------------------
DEFAULTS = dict(
    samplingRate=2.4e6, gain=19.1, centerFreq=92e6, fftSize=2 ** 14, curScanNonOverlap=0.1, curScanCumuMode=CUMUMODE_AVG,
    window="WIN.ONES", minAmp4Clip=(1 / 256) * 0.00001, scanRangeNonOverlap=0.5, prgLoopCnt=8192, xRes=512,
    pltCompress=PLTCOMPRESS_AVG, pltCompressHM=PLTCOMPRESS_MAX, AdjSigLvls="", SaveSigLvls="", bDataMin=True, bDataMax=True,
    bDataAvg=True, bDataCur=True, bScanRangeBaseDataIsRaw=False, bPltHeatMap=True, bPltLevels=False, bUsePSD=False,
    zeroSpanSaveFile="/tmp/zerospan.save", zeroSpanPlayFile="/tmp/zerospan.save",
)


def derive_config(d):
    """Fill ``d`` the way handle_args does after parsing (K:926-949): fullSize rule, the four window tables,
    theWin, xRes clamp / auto sub-multiple (without the blocking input(), K:947)."""
    for k, v in DEFAULTS.items():
        d.setdefault(k, v)
    d.setdefault("cmd.stop", False)
    F = int(d["fftSize"])
    d["fullSize"] = F * gFft2FullMult4Less if F < (d["samplingRate"] // 8) else F * gFft2FullMult4More
    d["WIN.HAMMING"] = np.hamming(F)
    d["WIN.HANNING"] = np.hanning(F)
    d["WIN.KAISER"] = np.kaiser(F, 64)
    d["WIN.ONES"] = np.ones(F)
    if not str(d["window"]).upper().startswith("WIN."):
        d["window"] = "WIN.%s" % str(d["window"]).upper()
    d["theWin"] = d[d["window"]]            # KeyError for an unknown window, as in the reference (K:936)
    if d["xRes"] > F:
        d["xRes"] = F
    elif F % d["xRes"] != 0:
        for i in range(int(F / 300), 0, -1):
            if F % i == 0:
                d["xRes"] = F // i
                break
    if "startFreq" not in d or "endFreq" not in d:
        d["startFreq"], d["endFreq"] = _calc_startendfreq(d["centerFreq"], d["samplingRate"])
    return d


def _calc_startendfreq(centerFreq, samplingRate):
    return centerFreq - samplingRate / 2, centerFreq + samplingRate / 2


def _fixupfreqs_scanrange(d):
    """K:701-709."""
    freqBands = (d["endFreq"] - d["startFreq"]) / d["samplingRate"]
    if (freqBands % 1) != 0:
        d["orig.EndFreq"] = d["endFreq"]
        d["endFreq"] = d["startFreq"] + np.ceil(freqBands) * d["samplingRate"]
    d["centerFreq"] = d["startFreq"] + ((d["endFreq"] - d["startFreq"]) / 2)


# ------------------------------------------------------------------------------------------------------
# plan cache: one GPU plan per (shape, window, mode, ingest format)
# ------------------------------------------------------------------------------------------------------
def _plan(d, in_fmt=_ffi.IN_C128):
    # bUsePSD (K:374-384) replaces the cumulate loop by a Welch PSD of the same scan: its own cumulate mode in the plan
    cumu = "PSD" if d.get("bUsePSD") else str(d["curScanCumuMode"]).upper()
    key = (int(d["fftSize"]), int(d["fullSize"]), float(d["curScanNonOverlap"]), cumu,
           d["window"], in_fmt, d.get("kspec.precision", "auto"), int(d.get("kspec.device", 0)))
    cache = d.setdefault("kspec.plans", {})
    if key not in cache:
        if str(d["curScanCumuMode"]).upper() not in _ffi.CUMU:
            prg_quit(d, "ERROR: Unknown cumuMode [{}], Quiting...".format(d["curScanCumuMode"]))      # K:144-146
        cache[key] = Plan(key[0], key[1], key[2], d["theWin"], key[3], in_fmt, precision=key[6], device=key[7],
                          u8_offset=d.get("kspec.u8Offset", 127.5), u8_scale=d.get("kspec.u8Scale", 1 / 127.5))
    return cache[key]


def close_plans(d):
    for p in d.pop("kspec.plans", {}).values():
        p.close()


# ------------------------------------------------------------------------------------------------------
# device read, host side, unchanged semantics (K:287-347)
# ------------------------------------------------------------------------------------------------------
def sdr_setup(sdr, fC, fS, gain, reopen=None):
    """K:287-308: tune, discard 16Ki settling samples; on failure close, reopen and report bOk=False."""
    try:
        sdr.sample_rate = fS
        sdr.center_freq = fC
        sdr.gain = gain
        bOk = True
        sdr.read_samples(16 * 1024)
    except Exception:
        print("WARN:SetupSDR:FAILED: fC[{}] fS[{}] gain[{}]".format(fC, fS, gain))
        sdr.close()
        if reopen is not None:
            sdr = reopen()
        bOk = False
    return sdr, bOk


def sdr_read(sdr, length):
    """K:311-347: reads above 2^18 are split; a non power-of-two tail is read rounded UP and the excess dropped."""
    length = int(length)
    if length > gSdrReadUnit:
        loopCnt, remaining, readLength = length // gSdrReadUnit, length % gSdrReadUnit, gSdrReadUnit
    else:
        loopCnt, remaining, readLength = 0, length, 0
    samples = np.zeros(length, dtype=complex)
    for i in range(loopCnt):
        samples[i * readLength:(i + 1) * readLength] = sdr.read_samples(readLength)
    if remaining > 0:
        iStart = gSdrReadUnit * loopCnt
        adjustedRead = int(2 ** np.ceil(np.log2(remaining)))
        samples[iStart:iStart + remaining] = sdr.read_samples(adjustedRead)[0:remaining]
    return samples


# ------------------------------------------------------------------------------------------------------
# the operator seam: sdr_curscan (K:351-397)
# ------------------------------------------------------------------------------------------------------
def sdr_curscan(d):
    """Read fullSize samples from d['sdr'] and return float64[fftSize]: windowed, normalised, cumulated,
    fftshift-ed magnitude spectrum -- computed by one fused GPU kernel (kspec_curscan)."""
    samples = sdr_read(d["sdr"], d["fullSize"])
    return _plan(d).curscan(samples)


def curscan_samples(d, samples):
    """sdr_curscan on samples already in memory (uint8 interleaved IQ, complex64 or complex128)."""
    samples = np.ascontiguousarray(samples)
    return _plan(d, _ffi.in_format(samples)).curscan(samples)


# ------------------------------------------------------------------------------------------------------
# zero_span (K:426-505): compute lines 464-484 for a block of scans in one GPU batch
# ------------------------------------------------------------------------------------------------------
def zero_span_init(d):
    """K:437-458: fresh Max/Min/Avg/Cur and the 128-row waterfall ring."""
    d["Fft.Max"] = d["Fft.Min"] = d["Fft.Avg"] = d["Fft.Cur"] = None
    d["kspec.state"] = None
    d["PltHeatMapWidth"] = heatmap_width(d["fftSize"], d["xRes"], d["pltCompressHM"])
    d["fftHM"] = np.zeros((HEATMAP_ROWS, d["PltHeatMapWidth"]))
    d["fftHMIndex"] = 0


def _carried_state(d):
    """(max, min, avg) carried from the earlier scans of this run, or None before the first one.  The GPU batch always
    computes all three; bDataMax / bDataMin / bDataAvg (K:471-476) decide which of them reach d['Fft.*']."""
    return d.get("kspec.state")


def _publish_stats(d, out):
    d["kspec.state"] = (out["max"], out["min"], out["avg"])
    if d.get("bDataMax", True):
        d["Fft.Max"] = out["max"]
    if d.get("bDataMin", True):
        d["Fft.Min"] = out["min"]
    if d.get("bDataAvg", True):
        d["Fft.Avg"] = out["avg"]


def zero_span_block(d, samples, n_scans, rows="db"):
    """The zero_span loop body for ``n_scans`` consecutive scans whose IQ is already in memory:
    sdr_curscan -> LogNoGain (K:469) -> Max/Min/Avg (K:471-476) -> waterfall rows (K:478-484).
    Updates d['Fft.*'] and the waterfall ring exactly as n_scans iterations of the reference loop would;
    returns the batch dict (rows = Fft.Cur of every scan when rows == 'db')."""
    samples = np.ascontiguousarray(samples)
    plan = _plan(d, _ffi.in_format(samples))
    state = _carried_state(d)
    adj = d["Fft.Adj"] if d.get("AdjSigLvls", "") != "" else None
    out = plan.zerospan_batch(samples, n_scans, d["gain"], d["xRes"], d["pltCompressHM"], adj=adj, rows=rows,
                              want_hm=True, state=state)
    _publish_stats(d, out)
    if rows == "db":
        d["Fft.Cur"] = out["rows"][-1]
    for r in out["hm_rows"]:                                   # ring of 128 rows (K:480-484)
        d["fftHM"][d["fftHMIndex"], :] = r
        d["fftHMIndex"] = (d["fftHMIndex"] + 1) % HEATMAP_ROWS
    return out


def zero_span(d, block=64):
    """Headless zero_span: reads prgLoopCnt scans from d['sdr'] in blocks and runs them on the GPU."""
    zero_span_init(d)
    d["sdr"], _ = sdr_setup(d["sdr"], d["centerFreq"], d["samplingRate"], d["gain"])
    done = 0
    while done < d["prgLoopCnt"] and not d["cmd.stop"]:
        n = min(block, d["prgLoopCnt"] - done)
        bufs = []
        try:
            for _ in range(n):
                bufs.append(sdr_read(d["sdr"], d["fullSize"]))
        except EOFError:
            d["cmd.stop"] = True
        if bufs:
            zero_span_block(d, np.concatenate(bufs), len(bufs))
        done += len(bufs)
    return done


def zero_span_u8_file(d, path, block=256, save=None, clock=time.time):
    """zeroSpan / zeroSpanSave over an rtl_sdr raw capture (interleaved uint8 I,Q, octave/load_rtlsdr.m:8-12) without the
    detour through complex128: the bytes are memory-mapped and handed to the GPU as they are (2 B per sample over PCIe, the
    uint8 -> float conversion is fused into the first FFT stage).  Scan k = samples [k*fullSize, (k+1)*fullSize), at most
    prgLoopCnt scans.  ``save``: optional open file; the zeroSpanSave records (K:523-525) are appended to it."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    S = int(d["fullSize"])
    n_total = min(int(d["prgLoopCnt"]), len(raw) // (2 * S))
    zero_span_init(d)
    plan = _plan(d, _ffi.IN_U8_IQ)
    done = 0
    while done < n_total and not d["cmd.stop"]:
        n = min(block, n_total - done)
        chunk = np.ascontiguousarray(raw[2 * S * done:2 * S * (done + n)])
        state = _carried_state(d)
        adj = d["Fft.Adj"] if d.get("AdjSigLvls", "") != "" else None
        t = clock()
        out = plan.zerospan_batch(chunk, n, d["gain"], d["xRes"], d["pltCompressHM"], adj=adj,
                                  rows="linear" if save is not None else "db", want_hm=True, state=state)
        _publish_stats(d, out)
        if save is not None:
            for row in out["rows"]:
                pickle.dump(t, save)
                pickle.dump(np.array(row), save)
        else:
            d["Fft.Cur"] = out["rows"][-1]
        for r in out["hm_rows"]:
            d["fftHM"][d["fftHMIndex"], :] = r
            d["fftHMIndex"] = (d["fftHMIndex"] + 1) % HEATMAP_ROWS
        done += n
    return done


# ------------------------------------------------------------------------------------------------------
# zeroSpanSave / zeroSpanPlay streams (K:509-564): format unchanged, FFT work batched on the GPU
# ------------------------------------------------------------------------------------------------------
def zero_span_save(d, block=64, clock=time.time):
    """K:510-526: header (centerFreq, samplingRate, gain) then per scan (time, float64[fftSize] linear)."""
    with open(d["zeroSpanSaveFile"], "wb+") as f:
        pickle.dump(d["centerFreq"], f)
        pickle.dump(d["samplingRate"], f)
        pickle.dump(d["gain"], f)
        d["sdr"], _ = sdr_setup(d["sdr"], d["centerFreq"], d["samplingRate"], d["gain"])
        clock()                                                # prevTime (K:515)
        done = 0
        while done < d["prgLoopCnt"] and not d["cmd.stop"]:
            n = min(block, d["prgLoopCnt"] - done)
            bufs, times = [], []
            try:
                for _ in range(n):
                    times.append(clock())
                    bufs.append(sdr_read(d["sdr"], d["fullSize"]))
            except EOFError:
                d["cmd.stop"] = True
                times = times[:len(bufs)]
            if bufs:
                samples = np.concatenate(bufs)
                plan = _plan(d, _ffi.in_format(samples))
                out = plan.zerospan_batch(samples, len(bufs), d["gain"], d["xRes"], d["pltCompressHM"], rows="linear",
                                          want_hm=False)
                for t, row in zip(times, out["rows"]):
                    pickle.dump(t, f)
                    pickle.dump(np.array(row), f)
            done += len(bufs)
    return done


def zero_span_play_setup(d):
    """K:530-543: open the stream, take centerFreq/samplingRate/gain from its header."""
    d["zeroSpanFile"] = f = open(d["zeroSpanPlayFile"], "rb")
    d["centerFreq"] = pickle.load(f)
    d["samplingRate"] = pickle.load(f)
    d["gain"] = pickle.load(f)
    d["startFreq"], d["endFreq"] = _calc_startendfreq(d["centerFreq"], d["samplingRate"])


def zero_span_play(d):
    """K:547-564: next (time, spectrum) record; EOF -> cmd.stop and None."""
    try:
        d["timeWas"] = pickle.load(d["zeroSpanFile"])
        timeWasMilli = int((d["timeWas"] - int(d["timeWas"])) * 1000)
        timeWas = time.strftime("%Y%m%d%Z%H%M%S", time.gmtime(d["timeWas"]))
        d["timeWasStr"] = "{}.{:03}".format(timeWas, timeWasMilli)
        data = pickle.load(d["zeroSpanFile"])
    except Exception:
        prg_quit(d, "WARN:zero_span_play:loading failed, stoping...", False)
        d["timeWas"] = 194700000
        data = None
    return data


def zero_span_play_all(d, block=64):
    """zeroSpanPlay end to end (do_run, K:1131-1134): zero_span with sdr_curscan re-bound to zero_span_play.  Records are
    read in blocks and the loop body (dB, Max/Min/Avg, waterfall rows, K:469-484) runs on the GPU.  Returns #scans."""
    zero_span_play_setup(d)
    zero_span_init(d)
    done = 0
    while done < d["prgLoopCnt"] and not d["cmd.stop"]:
        recs = []
        while len(recs) < min(block, d["prgLoopCnt"] - done):
            r = zero_span_play(d)
            if r is None:
                break
            recs.append(r)
        if not recs:
            break
        state = _carried_state(d)
        adj = d["Fft.Adj"] if d.get("AdjSigLvls", "") != "" else None
        out = _plan(d).zerospan_rows_batch(np.array(recs), d["gain"], d["xRes"], d["pltCompressHM"], adj=adj, state=state)
        _publish_stats(d, out)
        d["Fft.Cur"] = out["rows"][-1]
        for r in out["hm_rows"]:
            d["fftHM"][d["fftHMIndex"], :] = r
            d["fftHMIndex"] = (d["fftHMIndex"] + 1) % HEATMAP_ROWS
        done += len(recs)
    d["zeroSpanFile"].close()
    return done


# ------------------------------------------------------------------------------------------------------
# stepped scan (K:569-732)
# ------------------------------------------------------------------------------------------------------
def scan_geometry(d):
    """Index arithmetic of _scan_range (K:588-600, K:621-629, K:688-689) with the reference's float64 expressions."""
    freqSpan = d["samplingRate"]
    R = d["scanRangeNonOverlap"]
    if ((freqSpan * R) % 1) != 0:
        prg_quit(d, "ERROR: freqSpan [{}] x scanRangeNonOverlap [{}] is not int".format(freqSpan, R))
    if ((d["fftSize"] * R) % 1) != 0:
        prg_quit(d, "ERROR: fftSize[{}] x scanRangeNonOverlap [{}] is not int".format(d["fftSize"], R))
    if not (0 < R <= 1):
        prg_quit(d, "ERROR: scanRangeNonOverlap [{}] must be in (0,1]".format(R))
    numGroups = int((d["endFreq"] - d["startFreq"]) / freqSpan)
    totalEntries = numGroups * d["fftSize"]
    curFreq = d["startFreq"] + freqSpan / 2
    startFreq = curFreq - freqSpan / 2
    steps = []
    i = 0
    while startFreq < d["endFreq"]:
        iStart = int(i * d["fftSize"] * R)
        steps.append((curFreq, iStart, int((i + 1) * d["fftSize"] * R)))
        curFreq += freqSpan * R
        startFreq = curFreq - freqSpan / 2
        i += 1
    return numGroups, totalEntries, steps


def _scan_range(d, freqsAll, fftAll, runCount=-1, reopen=None):
    """K:569-698.  Host: retune + read every step (K:630-639).  GPU (one batch): per-step sdr_curscan, clip, dB,
    overlap stitch, Max/Min/Avg (K:640-668).  Host: frequency axis (K:631-634) and the waterfall ring (K:696-697)."""
    numGroups, totalEntries, steps = scan_geometry(d)
    F, S = d["fftSize"], d["fullSize"]
    if freqsAll is None:                                           # K:601-614
        floor = 10 * np.log10(np.ones(totalEntries) * d["minAmp4Clip"]) - d["gain"]
        floor[np.isinf(floor)] = 0
        d["Fft.Cur"], d["Fft.Max"], d["Fft.Avg"] = floor.copy(), floor.copy(), floor.copy()
        d["Fft.Min"] = 10 * np.log10(np.ones(totalEntries)) - d["gain"]
        span = numGroups * d["samplingRate"]
        freqsAll = np.fft.fftshift(np.fft.fftfreq(totalEntries, 1 / span) + d["startFreq"] + span / 2)
        fftAll = np.ones(totalEntries)
        W = totalEntries if d["pltCompressHM"] == PLTCOMPRESS_RAW or totalEntries // d["xRes"] == 0 else d["xRes"]
        d["fftHMMax"], d["fftHMIndex"] = HEATMAP_ROWS, 0
        d["fftHM"] = np.full((HEATMAP_ROWS, W), _hm_init_value(d, totalEntries))
    samples = np.zeros(len(steps) * S, dtype=complex)
    ok = np.ones(len(steps), dtype=np.uint8)
    for i, (curFreq, iStart, iDone) in enumerate(steps):
        d["sdr"], bOk = sdr_setup(d["sdr"], curFreq, d["samplingRate"], d["gain"], reopen)
        iEnd = iStart + F
        sEnd = F - max(0, iEnd - totalEntries)
        freqs = np.fft.fftshift(np.fft.fftfreq(F, 1 / d["samplingRate"]) + curFreq)
        freqsAll[iStart:iEnd] = freqs[0:sEnd]
        if bOk:
            samples[i * S:(i + 1) * S] = sdr_read(d["sdr"], S)
        else:
            print("WARN:_scanRange: Dummy data for step {}".format(i))
            ok[i] = 0
    # bDataMax / bDataMin gate the Max / Min update (K:663-666; Avg is always updated, K:667): a disabled curve keeps its
    # values because the batch works on a scratch copy of it
    state = dict(cur=d["Fft.Cur"], max=d["Fft.Max"] if d.get("bDataMax", True) else d["Fft.Max"].copy(),
                 min=d["Fft.Min"] if d.get("bDataMin", True) else d["Fft.Min"].copy(), avg=d["Fft.Avg"])
    _plan(d).scan_batch(samples, len(steps), [s[1] for s in steps], [s[2] for s in steps], totalEntries, d["minAmp4Clip"],
                        d["gain"], state, 0 if runCount == 0 else 1, step_ok=ok, base_is_raw=d["bScanRangeBaseDataIsRaw"])
    fftAvg = d["Fft.Avg"] - d["Fft.Adj"] if d.get("AdjSigLvls", "") != "" else d["Fft.Avg"]
    d["fftHM"][d["fftHMIndex"], :] = _data_plotcompress(d, fftAvg, d["pltCompressHM"])      # K:696-697
    return freqsAll, fftAll


def _hm_init_value(d, totalEntries):
    """K:612-614: the ring starts as compress(ones*minAmp4Clip): a constant row."""
    return d["minAmp4Clip"]


def scan_range(d, reopen=None):
    """K:712-732 without plotting."""
    freqs = ffts = None
    _fixupfreqs_scanrange(d)
    for i in range(d["prgLoopCnt"]):
        if d["cmd.stop"]:
            break
        freqs, ffts = _scan_range(d, freqs, ffts, i, reopen)
        d["fftHMIndex"] = (d["fftHMIndex"] + 1) % d["fftHMMax"]
    return freqs, ffts


# ------------------------------------------------------------------------------------------------------
# array helpers with the reference's names (K:124-237); GPU for the reductions
# ------------------------------------------------------------------------------------------------------
def _data_plotcompress(d, data, mode):
    """K:168-202: RAW identity, MAX/AVG over xRes groups (GPU); MIN/unknown quit like the reference (K:188, K:202)."""
    mode = str(mode).upper()
    if mode == PLTCOMPRESS_RAW:
        return data
    if mode == "CONV":                                   # PLTCOMPRESS_CONV -> data_proc 'Conv' (K:113-120, K:183-184)
        return _plan(d).conv_smooth(data, DataProcConv)
    if mode in (PLTCOMPRESS_MAX, PLTCOMPRESS_AVG):
        if len(data) // d["xRes"] == 0:
            return data
        return _plan(d).plotcompress(data, d["xRes"], mode)
    prg_quit(d, "ERROR:_data_plotcompress: Unknown mode [{}]".format(mode))


DataProcConv = np.kaiser(128, 64)                        # K:87


def plot_highs(d, freqs, levels):
    """K:243-272 without the matplotlib calls: the strongest points, at least pltHighsDelta4Marking of the span apart.
    Returns [(freq, level), ...] strongest first and prints them like the reference does."""
    idx = _plan(d).plot_highs(freqs, levels, d.get("pltHighsNumMarkers", 5), d.get("pltHighsDelta4Marking", 0.025))
    marks = [(float(freqs[i]), float(levels[i])) for i in idx]
    for f, lv in marks:
        print("plotHighs:Marked: {}, {}".format(f, lv))
    return marks


def data_plotcompress(d, xData, yData, mode=None):
    """K:205-221: x axis always AVG-compressed, y per ``mode``."""
    if mode is None:
        mode = d["pltCompress"]
    if str(mode).upper() == PLTCOMPRESS_RAW:
        return xData, yData
    if str(mode).upper() == "CONV":
        return xData, _data_plotcompress(d, yData, mode)
    return _data_plotcompress(d, xData, PLTCOMPRESS_AVG), _data_plotcompress(d, yData, mode)


def data_2d_plotcompress(d, data, mode=None):
    """K:224-237."""
    if mode is None:
        mode = d["pltCompressHM"]
    if str(mode).upper() == PLTCOMPRESS_RAW:
        return data
    return np.array([_data_plotcompress(d, row, mode) for row in data])


def main(argv=None):  # pragma: no cover - convenience entry, see INTEGRATION.md
    print("kspec.hotpath is a library; see INTEGRATION.md for wiring it into kspecanal.py", file=sys.stderr)
    return 2
