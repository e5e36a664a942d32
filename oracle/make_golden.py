#!/usr/bin/env python3
"""ORACLE support — generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (container only).

TEST INFRASTRUCTURE.  Usage (from the repo root, in the build container where /root/reference
exists):   python oracle/make_golden.py

Every fixture holds the exact inputs (so nothing depends on RNG stream stability) and the outputs of
the reference's own functions: sdr_curscan (K:351-397), fftvals_dispproc (K:150-165), zero_span
(K:426-505), zero_span_save (K:510-526), _scan_range (K:569-698), handle_args (K:778-949).
Sizes are reduced so that the whole golden directory stays a few MB; the full BASELINE sizes are
covered on the GPU by seeded oracle-vs-CUDA tests.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))
sys.path.insert(0, ROOT)

from kspec import synth  # noqa: E402  (input generators only; no GPU code)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def run_zerospan(argv, capture, n_scans, adj=None):
    """Drive the reference's zero_span() headless; return per-scan and final products."""
    holder = {}
    ns = ref_loader.load(sdr_factory=lambda: holder["sdr"])
    d = ref_loader.base_dict(ns, list(argv) + ["prgLoopCnt", n_scans, "bPltLevels", "false"])
    full = d["fullSize"]
    # (a) per-scan products by calling the reference functions directly
    lin_rows, db_rows = [], []
    for k in range(n_scans):
        d["sdr"] = synth.ArrayRtlSdr(capture[k * full:(k + 1) * full])
        with quiet():
            lin = ns["sdr_curscan"](d)
            db = ns["fftvals_dispproc"](d, lin, "LogNoGain")
        lin_rows.append(lin)
        db_rows.append(db)
    # (b) the outer loop itself, for max/min/avg and the waterfall ring
    d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(capture)
    if adj is not None:
        d["AdjSigLvls"] = "x"
        d["Fft.Adj"] = adj
    with quiet():
        ns["zero_span"](d)
    hm = np.array(d["AxHeatMap"].imshow.call_args[0][0])
    offs = [int(i * d["fftSize"] * d["curScanNonOverlap"]) for i in range(int(full / (d["fftSize"] * d["curScanNonOverlap"])))]
    offs = [o for o in offs if o + d["fftSize"] <= full]
    return dict(
        params=json.dumps(dict(fftSize=d["fftSize"], fullSize=full, curScanNonOverlap=d["curScanNonOverlap"],
                               curScanCumuMode=d["curScanCumuMode"], window=d["window"], gain=d["gain"],
                               xRes=d["xRes"], pltCompressHM=d["pltCompressHM"], samplingRate=d["samplingRate"],
                               nScans=n_scans)),
        window=np.array(d["theWin"]), offsets=np.array(offs, dtype=np.int64),
        lin_rows=np.array(lin_rows), db_rows=np.array(db_rows), hm=hm,
        fft_max=d["Fft.Max"], fft_min=d["Fft.Min"], fft_avg=d["Fft.Avg"],
    )


def run_scan(argv, step_bufs, n_pass, fail_steps=()):
    """Drive the reference's _scan_range() headless for n_pass passes."""
    holder = {}
    ns = ref_loader.load(sdr_factory=lambda: holder["sdr"])
    d = ref_loader.base_dict(ns, list(argv) + ["bPltLevels", "false"])
    n_steps = len(step_bufs)
    states, freqs, ffts = [], None, None
    for p in range(n_pass):
        fails = {s + 0 for s in fail_steps} if p == 0 else set()
        d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(per_tune=lambda t, fc, n: step_bufs[t], fail_tunes=fails)
        with quiet():
            freqs, ffts = ns["_scan_range"](d, freqs, ffts, p)
        states.append([np.array(d[k]) for k in ("Fft.Cur", "Fft.Max", "Fft.Min", "Fft.Avg")])
        d["fftHMIndex"] = (d["fftHMIndex"] + 1) % d["fftHMMax"]     # what scan_range does, K:732
    # per-step sdr_curscan outputs, for the stitch-only oracle entry point
    lin_rows = []
    for s in range(n_steps):
        d["sdr"] = synth.ArrayRtlSdr(step_bufs[s])
        with quiet():
            lin_rows.append(ns["sdr_curscan"](d))
    out = dict(
        params=json.dumps(dict(fftSize=d["fftSize"], fullSize=d["fullSize"], curScanNonOverlap=d["curScanNonOverlap"],
                               curScanCumuMode=d["curScanCumuMode"], window=d["window"], gain=d["gain"],
                               xRes=d["xRes"], pltCompressHM=d["pltCompressHM"], samplingRate=d["samplingRate"],
                               scanRangeNonOverlap=d["scanRangeNonOverlap"], startFreq=d["startFreq"],
                               endFreq=d["endFreq"], minAmp4Clip=d["minAmp4Clip"], nSteps=n_steps, nPass=n_pass,
                               failSteps=list(fail_steps),
                               bScanRangeBaseDataIsRaw=d["bScanRangeBaseDataIsRaw"])),
        window=np.array(d["theWin"]), freqs_all=np.array(freqs), lin_rows=np.array(lin_rows),
        hm=np.array(d["fftHM"][:n_pass]),
    )
    for p, st in enumerate(states):
        for name, arr in zip(("cur", "max", "min", "avg"), st):
            out["p%d_%s" % (p, name)] = arr
    return out


def n_scan_steps(start, end, fs, r):
    n, s = 0, start
    cur = start + fs / 2
    while s < end:
        cur += fs * r
        s = cur - fs / 2
        n += 1
    return n


def main():
    assert ref_loader.available(), "run in the build container: /root/reference is required"
    os.makedirs(OUT, exist_ok=True)

    # g1: cfg-1 shape (F=2048, hanning, 50% overlap, AVG), 6 scans, complex64 capture
    F, n = 2048, 4
    cap = synth.tones_noise(n * F * 8, seed=1)
    g = run_zerospan(["zeroSpan", "fftSize", F, "window", "hanning", "curScanNonOverlap", 0.5], cap, n)
    np.savez_compressed(os.path.join(OUT, "g1_zerospan_2048_hanning.npz"), capture=cap, **g)

    # g1b: same with an adjSigLvls baseline, MAX cumulate, default overlap 0.1 (non-uniform hops), 3 scans, hamming
    n = 3
    cap = synth.tones_noise(n * 1024 * 8, seed=11)
    adj = np.linspace(-3.0, 3.0, 1024)
    g = run_zerospan(["zeroSpan", "fftSize", 1024, "window", "hamming", "curScanCumuMode", "max", "xRes", 256], cap, n, adj=adj)
    np.savez_compressed(os.path.join(OUT, "g1b_zerospan_1024_hamming_adj.npz"), capture=cap, adj=adj, **g)

    # g3: cfg-3 shape (F=8192, kaiser(64), 75% overlap), 2 scans
    F, n = 8192, 2
    cap = synth.tones_noise(n * F * 8, seed=3, gate=(40000, 0.5))
    g = run_zerospan(["zeroSpan", "fftSize", F, "window", "kaiser", "curScanNonOverlap", 0.25], cap, n)
    np.savez_compressed(os.path.join(OUT, "g3_zerospan_8192_kaiser.npz"), capture=cap, **g)

    # g4: cfg-4 family (big pow2 frame through the multi-pass path, ones window, MAX cumulate), uint8 ingest
    F, n = 32768, 1
    cap_u8 = synth.to_u8_iq(synth.tones_noise(n * F * 8, seed=4, dtype=np.complex128))
    cap = synth.from_u8_iq(cap_u8)
    g = run_zerospan(["zeroSpan", "fftSize", F, "window", "ones", "curScanCumuMode", "max"], cap, n)
    np.savez_compressed(os.path.join(OUT, "g4_zerospan_32768_ones_max_u8.npz"), capture_u8=cap_u8, **g)

    # g2: cfg-2 shape (quickFullScan: F=64, ones, r=0.1, RAW plot) on a short range, R=1.0 and R=0.5,
    #     two passes, one tune failure in pass 0
    fs = 2.4e6
    for tag, R in (("r100", 1.0), ("r050", 0.5)):
        start, end = 30e6, 30e6 + 10.3 * fs            # end gets rounded up to 11 bands (K:701-709)
        nst = n_scan_steps(start, start + 11 * fs, fs, R)
        bufs = [synth.step_tones(s, 512) for s in range(nst)]
        g = run_scan(["scan", "startFreq", start, "endFreq", end, "fftSize", 64, "pltCompress", "raw",
                      "scanRangeNonOverlap", R], bufs, n_pass=2, fail_steps=(3,))
        np.savez_compressed(os.path.join(OUT, "g2_scan_64_%s.npz" % tag), step_bufs=np.array(bufs), **g)

    # g5a: cfg-5a shape (fmScan geometry 88..108 MHz -> 9 groups, 18 steps, F=4096, r=0.1, R=0.5), uint8 ingest
    nst = n_scan_steps(88e6, 88e6 + 9 * fs, fs, 0.5)
    bufs_u8 = [synth.to_u8_iq(synth.step_tones(s, 4096 * 8, dtype=np.complex128)) for s in range(nst)]
    bufs = [synth.from_u8_iq(b) for b in bufs_u8]
    g = run_scan(["fmScan", "fftSize", 4096], bufs, n_pass=1)
    np.savez_compressed(os.path.join(OUT, "g5a_fmscan_4096_u8.npz"), step_bufs_u8=np.array(bufs_u8), **g)

    # g5b: cfg-5b family (non power of two frame -> Bluestein), F=1200, R=0.25, base-data-is-raw variant too
    for tag, extra in (("cur", []), ("raw", ["bScanRangeBaseDataIsRaw", "true"])):
        start, end, R = 100e6, 100e6 + 3 * fs, 0.25
        nst = n_scan_steps(start, end, fs, R)
        bufs = [synth.step_tones(s + 100, 1200 * 8) for s in range(nst)]
        g = run_scan(["scan", "startFreq", start, "endFreq", end, "fftSize", 1200, "xRes", 300, "window", "hanning",
                      "scanRangeNonOverlap", R] + extra, bufs, n_pass=2)
        if tag == "raw":        # same inputs as the "cur" fixture: keep only what differs
            g = {k: v for k, v in g.items() if k not in ("lin_rows", "freqs_all", "window")}
            np.savez_compressed(os.path.join(OUT, "g5b_scan_1200_raw.npz"), **g)
        else:
            np.savez_compressed(os.path.join(OUT, "g5b_scan_1200_cur.npz"), step_bufs=np.array(bufs), **g)

    # g6: zeroSpanSave stream written by the reference with a frozen clock (format parity, K:510-526)
    F, n = 64, 3
    cap = synth.tones_noise(n * F * 8, seed=6)
    holder = {}
    ns = ref_loader.load(sdr_factory=lambda: holder["sdr"])
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "z.save")
        d = ref_loader.base_dict(ns, ["zeroSpanSave", "fftSize", F, "centerFreq", 881e6, "samplingRate", 2.4e6,
                                      "prgLoopCnt", n, "zeroSpanSaveFile", path])
        d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(cap)
        clock = iter([1000.0 + 0.25 * i for i in range(100)])
        with quiet(), mock.patch.object(ns["time"], "time", lambda: next(clock)):
            ns["zero_span_save"](d)
        blob = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "g6_zerospansave_64.npz"), capture=cap, blob=blob,
                        params=json.dumps(dict(fftSize=F, nScans=n, centerFreq=881e6, samplingRate=2.4e6,
                                               gain=d["gain"], curScanNonOverlap=d["curScanNonOverlap"],
                                               times=[1000.25 + 0.25 * i for i in range(n)])))

    # g7: derived-configuration table from handle_args (fullSize rule, xRes fix-up, scan end fix-up)
    rows = []
    ns = ref_loader.load()
    for argv in (["zeroSpan", "fftSize", 2048], ["zeroSpan", "fftSize", 64], ["zeroSpan", "fftSize", 2 ** 21],
                 ["zeroSpan", "fftSize", 2400000], ["zeroSpan", "fftSize", 300000], ["zeroSpan", "fftSize", 299999],
                 ["zeroSpan", "fftSize", 1200, "xRes", 512], ["quickFullScan"], ["fmScan"],
                 ["scan", "startFreq", 80e6, "endFreq", 120e6], ["zeroSpan", "fftSize", 16384, "xRes", 500]):
        with quiet():
            d = ref_loader.base_dict(ns, argv)
        rows.append(dict(argv=[str(a) for a in argv], fftSize=d["fftSize"], fullSize=d["fullSize"], xRes=d["xRes"],
                         startFreq=d["startFreq"], endFreq=d["endFreq"], centerFreq=d["centerFreq"],
                         pltCompress=d["pltCompress"], prgMode=d["prgMode"]))
    with open(os.path.join(OUT, "g7_handle_args.json"), "w") as f:
        json.dump(rows, f, indent=1)

    for fn in sorted(os.listdir(OUT)):
        print("%9d  %s" % (os.path.getsize(os.path.join(OUT, fn)), fn))


if __name__ == "__main__":
    main()
