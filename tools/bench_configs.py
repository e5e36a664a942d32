#!/usr/bin/env python3
"""Secondary benchmark: throughput of the five BASELINE.json config shapes (not the headline; bench.py is).
Prints one JSON line per config: IQ Msamples/s, frames/s, path, precision and ``roofline_frac`` = algorithmic bytes / time /
measured HBM peak with the SURVEY 8(d) formula  nScans*S*b_in + nScans*W_row*rb + 4*F*rb  (b_in = bytes per IQ sample of the
ingest format, W_row = waterfall row width, rb = 4 or 8 = working precision; stated per line).  zeroSpan shapes are timed
device-resident through kspec_zerospan_batch_dev; scan shapes through kspec_scan_pass_dev (engine + scan_stitch_kernel on
the device-resident state) and, as ``e2e``, through kspec_scan_pass from pinned host memory (chunked H2D inside the timed
region).  Run on a GPU box:  python tools/bench_configs.py > profiles/r2_configs.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))
from kspec import _ffi, synth                      # noqa: E402
from kspec.engine import Plan                      # noqa: E402
from kspec.hotpath import derive_config            # noqa: E402

FS = 2.4e6


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)).get("hbm_gbs", 6650.0) if os.path.isfile(p) else 6650.0


# stepped scans: name, fftSize, window, nonOverlap (curScan), scanRangeNonOverlap, groups, precision, ingest, env
SCANS = [
    ("cfg2 quickFullScan 64 ones r=0.1 R=1.0 (613 steps) f64", 64, "ones", 0.1, 1.0, 613, "f64", "c64"),
    ("cfg2 quickFullScan 64 ones r=0.1 R=0.5 (1226 steps) f64", 64, "ones", 0.1, 0.5, 613, "f64", "c64"),
    ("cfg2 quickFullScan 64 ones r=0.1 R=0.5 (1226 steps) f32", 64, "ones", 0.1, 0.5, 613, "f32", "c64"),
    ("cfg5a fmScan 4096 ones r=0.1 R=0.5 (18 steps) f64", 4096, "ones", 0.1, 0.5, 9, "f64", "c64"),
    ("cfg5a fmScan 4096 ones r=0.1 R=0.5 (18 steps) f64 uint8 ingest", 4096, "ones", 0.1, 0.5, 9, "f64", "u8"),
    ("cfg5b fmScan 2400000 ones r=0.1 R=0.5 (18 steps, mixed radix) f64", 2400000, "ones", 0.1, 0.5, 9, "f64", "c64"),
]

CONFIGS = [
    # name, fftSize, window, nonOverlap, cumu, precision, scans per batch, ingest
    ("cfg1 zeroSpan 2048 hanning 50% AVG f32", 2048, "hanning", 0.5, "AVG", "f32", 16384, "c64"),
    ("cfg1 zeroSpan 2048 hanning 50% AVG f64", 2048, "hanning", 0.5, "AVG", "f64", 16384, "c64"),
    ("cfg1 non-overlapped (r=1.0) f32", 2048, "hanning", 1.0, "AVG", "f32", 16384, "c64"),
    ("cfg1 reference default overlap (r=0.1) f32", 2048, "hanning", 0.1, "AVG", "f32", 16384, "c64"),
    ("cfg1 uint8 ingest f32", 2048, "hanning", 0.5, "AVG", "f32", 16384, "u8"),
    ("cfg2 scan 64 ones r=0.1 (1226 steps) f32", 64, "ones", 0.1, "AVG", "f32", 1226, "c64"),
    ("cfg2 scan 64 ones r=0.1 (1226 steps) f64", 64, "ones", 0.1, "AVG", "f64", 1226, "c64"),
    ("cfg3 zeroSpan 8192 kaiser 75% f64", 8192, "kaiser", 0.25, "AVG", "f64", 2197, "c64"),
    ("cfg3 zeroSpan 8192 kaiser 75% f32", 8192, "kaiser", 0.25, "AVG", "f32", 2197, "c64"),
    ("cfg5a scan 4096 ones r=0.1 f64", 4096, "ones", 0.1, "AVG", "f64", 1024, "c64"),
    ("reference default 16384 ones r=0.1 f32", 16384, "ones", 0.1, "AVG", "f32", 512, "c64"),
    ("cfg4 2^21 ones MAX r=0.1 (four-step) f64", 1 << 21, "ones", 0.1, "MAX", "f64", 4, "c64"),
    ("cfg4 2^21 ones MAX r=0.1 (four-step, element-wise column pass) f64", 1 << 21, "ones", 0.1, "MAX", "f64", 4, "c64", {"KSPEC_FOURSTEP_TILED": "0"}),
    ("cfg4 2^21 ones MAX r=0.1 (mixed radix engine, forced) f64", 1 << 21, "ones", 0.1, "MAX", "f64", 4, "c64", {"KSPEC_FORCE_MIXED": "1"}),
    ("reference default 16384 ones r=0.1 (four-step) f64", 16384, "ones", 0.1, "AVG", "f64", 512, "c64"),
    ("reference default 16384 ones r=0.1 (mixed radix engine, forced) f64", 16384, "ones", 0.1, "AVG", "f64", 512, "c64", {"KSPEC_FORCE_MIXED": "1"}),
    ("cfg5b 2400000 ones r=0.1 (mixed radix 1500x1600) f64", 2400000, "ones", 0.1, "AVG", "f64", 4, "c64"),
    ("cfg5b 2400000 ones r=0.1 (Bluestein 2^23, forced) f64", 2400000, "ones", 0.1, "AVG", "f64", 2, "c64", {"KSPEC_FORCE_BLUESTEIN": "1"}),
    ("48000 hanning r=0.5 (mixed radix 200x240) f64", 48000, "hanning", 0.5, "AVG", "f64", 256, "c64"),
    ("48000 hanning r=0.5 (Bluestein 2^17, forced) f64", 48000, "hanning", 0.5, "AVG", "f64", 256, "c64", {"KSPEC_FORCE_BLUESTEIN": "1"}),
]


def run_scan(name, F, win, r, R, groups, prec, ingest):
    from kspec.hotpath import scan_geometry
    d = derive_config(dict(fftSize=F, window=win, samplingRate=FS, curScanNonOverlap=r, scanRangeNonOverlap=R,
                           startFreq=88e6, endFreq=88e6 + groups * FS))
    S = d["fullSize"]
    _, total, steps = scan_geometry(d)
    ns = len(steps)
    i_start, i_done = [s[1] for s in steps], [s[2] for s in steps]
    eb = 2 if ingest == "u8" else 8
    pinned = _ffi.PinnedBuffer(ns * S * eb)
    host = pinned.view(np.uint8 if ingest == "u8" else np.complex64)
    blk = synth.tones_noise(min(ns, 4) * S, seed=5)
    if ingest == "u8":
        blk = synth.to_u8_iq(blk.astype(np.complex128))
    per = len(blk) // min(ns, 4)
    for i in range(ns):
        host[i * per:(i + 1) * per] = blk[(i % min(ns, 4)) * per:((i % min(ns, 4)) + 1) * per]
    plan = Plan(F, S, r, d["theWin"], "AVG", _ffi.IN_U8_IQ if ingest == "u8" else _ffi.IN_C64, precision=prec)
    floor = 10 * np.log10(np.ones(total) * d["minAmp4Clip"]) - d["gain"]
    plan.scan_state_init(dict(cur=floor, max=floor.copy(), min=np.zeros(total) - d["gain"], avg=floor.copy()))
    dptr = plan.dev_alloc(ns * S * eb)
    plan.dev_upload(dptr, host)
    reps = 3 if F >= (1 << 20) else 10

    def timed(fn):
        for _ in range(2):
            fn()
        plan.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        plan.sync()
        return (time.perf_counter() - t0) / reps * 1e3

    ms_dev = timed(lambda: plan.scan_pass(dptr, ns, i_start, i_done, d["minAmp4Clip"], d["gain"], 1, on_device=True))
    ms_e2e = timed(lambda: plan.scan_pass(host, ns, i_start, i_done, d["minAmp4Clip"], d["gain"], 1))
    t0 = time.perf_counter()
    plan.scan_state_fetch()
    ms_fetch = (time.perf_counter() - t0) * 1e3
    rb = 4 if plan.precision == "f32" else 8
    alg = ns * S * eb + 4 * total * 8 * 2 + 4 * F * rb       # samples once, the four state vectors read and written once, tables
    print(json.dumps({"config": name, "mode": "scan", "path": plan.path, "precision": plan.precision, "steps": ns, "frames_per_step": plan.n_frames,
                      "total_entries": total, "ms_per_pass": round(ms_dev, 4), "Msamples_per_s": round(ns * S / ms_dev / 1e3, 1),
                      "frames_per_s": round(ns * plan.n_frames / ms_dev * 1e3, 1), "b_in": eb, "algorithmic_bytes": alg,
                      "roofline_frac": round(alg / (ms_dev * 1e-3) / 1e9 / hbm_peak(), 5),
                      "e2e": {"ms_per_pass": round(ms_e2e, 4), "Msamples_per_s": round(ns * S / ms_e2e / 1e3, 1), "h2d_bytes_per_pass": ns * S * eb,
                              "h2d_GBps": round(ns * S * eb / ms_e2e / 1e6, 2), "state_fetch_ms_(not_in_e2e)": round(ms_fetch, 3)}}), flush=True)
    plan.dev_free(dptr)
    plan.close()
    pinned.free()


def main():
    only = os.environ.get("KSPEC_CFG_FILTER", "")
    for cfg in SCANS:
        if only and only not in cfg[0]:
            continue
        run_scan(*cfg)
    for cfg in CONFIGS:
        name, F, win, r, cumu, prec, n_scans, ingest = cfg[:8]
        if only and only not in name:
            continue
        env = cfg[8] if len(cfg) > 8 else {}
        os.environ.update(env)
        try:
            run_one(name, F, win, r, cumu, prec, n_scans, ingest)
        finally:
            for k in env:
                os.environ.pop(k, None)


def run_one(name, F, win, r, cumu, prec, n_scans, ingest):
    if True:
        d = derive_config(dict(fftSize=F, window=win, samplingRate=FS))
        S = d["fullSize"]
        fmt = _ffi.IN_U8_IQ if ingest == "u8" else _ffi.IN_C64
        base_scans = min(n_scans, 64)
        base = synth.tones_noise(base_scans * S, seed=1)
        if ingest == "u8":
            base = synth.to_u8_iq(base.astype(np.complex128))
        plan = Plan(F, S, r, d["theWin"], cumu, fmt, precision=prec)
        eb = 2 if ingest == "u8" else 8
        dptr = plan.dev_alloc(n_scans * S * eb)
        for i in range(n_scans // base_scans):
            plan.dev_upload(dptr, base, offset=i * base.nbytes)
        rem = n_scans % base_scans
        if rem:
            plan.dev_upload(dptr, base[:rem * S * (2 if ingest == "u8" else 1)], offset=(n_scans // base_scans) * base.nbytes)
        xres = d["xRes"]
        reps = 3 if F >= (1 << 20) else 10
        for _ in range(2):
            plan.zerospan_batch_dev(dptr, n_scans, 19.1, xres, "MAX", want_hm=True)
        plan.sync()
        plan.timer_start()
        for _ in range(reps):
            plan.zerospan_batch_dev(dptr, n_scans, 19.1, xres, "MAX", want_hm=True)
        ms = plan.timer_stop() / reps
        rb = 4 if plan.precision == "f32" else 8
        W = xres if F > xres else F
        alg = n_scans * S * eb + n_scans * W * rb + 4 * F * rb
        print(json.dumps({"config": name, "mode": "zeroSpan", "path": plan.path, "precision": plan.precision, "scans": n_scans, "frames_per_scan": plan.n_frames,
                          "ms_per_batch": round(ms, 4), "Msamples_per_s": round(n_scans * S / ms / 1e3, 1),
                          "frames_per_s": round(n_scans * plan.n_frames / ms * 1e3, 1), "b_in": eb, "algorithmic_bytes": alg,
                          "roofline_frac": round(alg / (ms * 1e-3) / 1e9 / hbm_peak(), 5)}), flush=True)
        plan.dev_free(dptr)
        plan.close()


if __name__ == "__main__":
    main()
