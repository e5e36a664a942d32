"""Headless command line with kspecanal.py's option set (handle_args, K:771-949), for machines without a dongle or a
display: the IQ comes from an rtl_sdr raw file or from the synthetic generator, the spectra from libkspec.so, and the
results go to an .npz file / the zeroSpanSave stream instead of a matplotlib window.

    python -m kspec.cli zeroSpanSave fftSize 2048 window hanning curScanNonOverlap 0.5 iqFile capture.bin \
                        zeroSpanSaveFile /tmp/cap.save prgLoopCnt 146
    python -m kspec.cli zeroSpanPlay fftSize 2048 zeroSpanPlayFile /tmp/cap.save outFile /tmp/play.npz
    python -m kspec.cli quickFullScan scanRangeNonOverlap 1.0 iqSynth 7 prgLoopCnt 1 outFile /tmp/scan.npz

Keys are case-insensitive ``KEY value`` pairs exactly as in the reference (unknown key -> quit, K:907-909); the extra keys
are iqFile (uint8 interleaved I,Q, octave/load_rtlsdr.m:8-12), iqSynth <seed>, outFile <npz>, precision auto|f32|f64.
Plot-only keys (bPltLevels, pltHighs*, bGrid ...) are accepted and ignored.
"""
import sys

import numpy as np

from . import hotpath as H
from . import synth

MODES = ("ZEROSPAN", "ZEROSPANSAVE", "ZEROSPANPLAY", "SCAN", "FMSCAN", "QUICKFULLSCAN")
FLOAT_KEYS = {"CENTERFREQ": "centerFreq", "STARTFREQ": "startFreq", "ENDFREQ": "endFreq", "SAMPLINGRATE": "samplingRate",
              "GAIN": "gain", "MINAMP4CLIP": "minAmp4Clip", "CURSCANNONOVERLAP": "curScanNonOverlap",
              "SCANRANGENONOVERLAP": "scanRangeNonOverlap", "PLTHIGHSDELTA4MARKING": "pltHighsDelta4Marking"}
INT_KEYS = {"FFTSIZE": "fftSize", "XRES": "xRes", "PRGLOOPCNT": "prgLoopCnt", "PLTHIGHSNUMMARKERS": "pltHighsNumMarkers",
            "IQSYNTH": "iqSynth"}
BOOL_KEYS = {"BDATAMIN": "bDataMin", "BDATAMAX": "bDataMax", "BDATAAVG": "bDataAvg", "BDATACUR": "bDataCur",
             "BPLTHEATMAP": "bPltHeatMap", "BPLTLEVELS": "bPltLevels", "PLTHIGHSPAUSE": "pltHighsPause", "BGRID": "bGrid",
             "BUSEPSD": "bUsePSD", "BSCANRANGEBASEDATAISRAW": "bScanRangeBaseDataIsRaw"}
STR_KEYS = {"SAVESIGLVLS": "SaveSigLvls", "ADJSIGLVLS": "AdjSigLvls", "ZEROSPANSAVEFILE": "zeroSpanSaveFile",
            "ZEROSPANPLAYFILE": "zeroSpanPlayFile", "IQFILE": "iqFile", "OUTFILE": "outFile"}
UPPER_KEYS = {"CURSCANCUMUMODE": "curScanCumuMode", "PLTCOMPRESS": "pltCompress"}


def handle_args(d, argv):
    """K:778-949: defaults, KEY value pairs, aliases, derived values."""
    d["prgMode"] = "FMSCAN"                                   # gPrgModeDefault, K:41
    i = 0
    while i < len(argv):
        cur = argv[i].upper()
        if cur in MODES:
            d["prgMode"] = cur
        elif cur in FLOAT_KEYS:
            i += 1
            d[FLOAT_KEYS[cur]] = float(argv[i])
        elif cur in INT_KEYS:
            i += 1
            d[INT_KEYS[cur]] = int(argv[i])
        elif cur in BOOL_KEYS:
            i += 1
            d[BOOL_KEYS[cur]] = argv[i].upper() == "TRUE"     # _arg_boolean, K:771-775
        elif cur in STR_KEYS:
            i += 1
            d[STR_KEYS[cur]] = argv[i]
        elif cur in UPPER_KEYS:
            i += 1
            d[UPPER_KEYS[cur]] = argv[i].upper()
        elif cur == "WINDOW":
            i += 1
            d["window"] = "WIN.{}".format(argv[i].upper())    # K:866-868
        elif cur == "PRECISION":
            i += 1
            d["kspec.precision"] = argv[i].lower()
        else:
            H.prg_quit(d, "ERROR:handle_args: Unknown argument [{}]".format(cur))
        i += 1
    d.setdefault("samplingRate", H.DEFAULTS["samplingRate"])
    if d["prgMode"] == "FMSCAN":                              # K:912-915
        d["prgMode"], d["startFreq"], d["endFreq"] = "SCAN", 88e6, 108e6
    elif d["prgMode"] == "QUICKFULLSCAN":                     # K:916-921
        d["prgMode"], d["startFreq"], d["endFreq"], d["fftSize"], d["pltCompress"] = "SCAN", 30e6, 1.5e9, 64, "RAW"
    if d["prgMode"] == "SCAN":
        d.setdefault("startFreq", 88e6)
        d.setdefault("endFreq", 108e6)
        H._fixupfreqs_scanrange(d)
    else:
        d.setdefault("centerFreq", H.DEFAULTS["centerFreq"])
        d["startFreq"], d["endFreq"] = H._calc_startendfreq(d["centerFreq"], d["samplingRate"])
    return H.derive_config(d)


def make_source(d):
    """The object that stands in for rtlsdr.RtlSdr() (K:1146)."""
    if d.get("iqFile"):
        return synth.ArrayRtlSdr.from_u8_file(d["iqFile"], d.get("kspec.u8Offset", 127.5), d.get("kspec.u8Scale", 1 / 127.5))
    seed = d.get("iqSynth", 1)
    if d["prgMode"] == "SCAN":
        return synth.ArrayRtlSdr(per_tune=lambda t, fc, n: synth.step_tones(t + seed, n, fs=d["samplingRate"]))
    n = d["prgLoopCnt"] * d["fullSize"]
    return synth.ArrayRtlSdr(synth.tones_noise(n, seed, fs=d["samplingRate"]))


def do_run(d):
    """K:1126-1136."""
    if d["prgMode"] == "SCAN":
        freqs, _ = H.scan_range(d)
        return dict(freqs=freqs)
    if d["prgMode"] == "ZEROSPANPLAY":
        return dict(nScans=H.zero_span_play_all(d))
    if d.get("iqFile"):                                        # raw uint8 capture: bytes go to the GPU unconverted
        if d["prgMode"] == "ZEROSPANSAVE":
            with open(d["zeroSpanSaveFile"], "wb+") as f:
                for k in ("centerFreq", "samplingRate", "gain"):   # K:512-514
                    H.pickle.dump(d[k], f)
                return dict(nScans=H.zero_span_u8_file(d, d["iqFile"], save=f))
        return dict(nScans=H.zero_span_u8_file(d, d["iqFile"]))
    if d["prgMode"] == "ZEROSPANSAVE":
        return dict(nScans=H.zero_span_save(d))
    return dict(nScans=H.zero_span(d))


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    d = {"cmd.stop": False}
    handle_args(d, argv)
    if not d.get("iqFile"):                                    # synthetic source: bounded run unless told otherwise
        d["prgLoopCnt"] = min(d["prgLoopCnt"], 2 if d["prgMode"] == "SCAN" else 256)
    if d["prgMode"] == "SCAN" or (d["prgMode"] != "ZEROSPANPLAY" and not d.get("iqFile")):
        d["sdr"] = make_source(d)
    print("INFO: prgMode[{}] fftSize[{}] fullSize[{}] window[{}] curScanNonOverlap[{}] xRes[{}]".format(
        d["prgMode"], d["fftSize"], d["fullSize"], d["window"], d["curScanNonOverlap"], d["xRes"]))
    res = do_run(d)
    out = {k: np.asarray(v) for k, v in res.items() if v is not None}
    for k in ("Fft.Cur", "Fft.Max", "Fft.Min", "Fft.Avg", "fftHM"):
        if d.get(k) is not None:
            out[k.replace(".", "_")] = np.asarray(d[k])
    if d.get("Fft.Max") is not None:
        j = int(np.argmax(d["Fft.Max"]))
        print("INFO: strongest bin {} of {}: max {:.2f} dB".format(j, len(d["Fft.Max"]), float(d["Fft.Max"][j])))
    if d.get("outFile"):
        np.savez_compressed(d["outFile"], **out)
        print("INFO: results written to", d["outFile"])
    H.close_plans(d)
    return 0


if __name__ == "__main__":
    sys.exit(main())
