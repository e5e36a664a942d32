#!/usr/bin/env python3
"""Secondary benchmark: device-resident throughput of the five BASELINE.json config shapes (not the headline; bench.py is).
Prints one JSON line per config: IQ Msamples/s, frames/s, path, precision.  Run on a GPU box:  python tools/bench_configs.py"""
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
    ("cfg4 2^21 ones MAX r=0.1 (mixed radix engine, forced) f64", 1 << 21, "ones", 0.1, "MAX", "f64", 4, "c64", {"KSPEC_FORCE_MIXED": "1"}),
    ("reference default 16384 ones r=0.1 (four-step) f64", 16384, "ones", 0.1, "AVG", "f64", 512, "c64"),
    ("reference default 16384 ones r=0.1 (mixed radix engine, forced) f64", 16384, "ones", 0.1, "AVG", "f64", 512, "c64", {"KSPEC_FORCE_MIXED": "1"}),
    ("cfg5b 2400000 ones r=0.1 (mixed radix 1500x1600) f64", 2400000, "ones", 0.1, "AVG", "f64", 4, "c64"),
    ("cfg5b 2400000 ones r=0.1 (Bluestein 2^23, forced) f64", 2400000, "ones", 0.1, "AVG", "f64", 2, "c64", {"KSPEC_FORCE_BLUESTEIN": "1"}),
    ("48000 hanning r=0.5 (mixed radix 200x240) f64", 48000, "hanning", 0.5, "AVG", "f64", 256, "c64"),
    ("48000 hanning r=0.5 (Bluestein 2^17, forced) f64", 48000, "hanning", 0.5, "AVG", "f64", 256, "c64", {"KSPEC_FORCE_BLUESTEIN": "1"}),
]


def main():
    only = os.environ.get("KSPEC_CFG_FILTER", "")
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
        print(json.dumps({"config": name, "path": plan.path, "precision": plan.precision, "scans": n_scans, "frames_per_scan": plan.n_frames,
                          "ms_per_batch": round(ms, 4), "Msamples_per_s": round(n_scans * S / ms / 1e3, 1),
                          "frames_per_s": round(n_scans * plan.n_frames / ms * 1e3, 1)}), flush=True)
        plan.dev_free(dptr)
        plan.close()


if __name__ == "__main__":
    main()
