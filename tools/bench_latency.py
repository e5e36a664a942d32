#!/usr/bin/env python3
"""Latency of the minimal drop-in: ONE scan per call through kspec_curscan (the re-bound sdr_curscan of INTEGRATION.md section 3),
complex128 samples in host memory as the reference's sdr_read returns them, float64 spectrum back.  Next to it the oracle port of the
reference's own loop (K:385-397) on one host core.  Prints one JSON line per shape.  Run on a GPU box: python tools/bench_latency.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))
sys.path.insert(0, ROOT)
from kspec import _ffi, synth                      # noqa: E402
from kspec.engine import Plan                      # noqa: E402
from kspec.hotpath import derive_config            # noqa: E402
from oracle import kspec_oracle as O               # noqa: E402  (CPU baseline leg only)

SHAPES = [(2048, "hanning", 0.5, "auto"), (2048, "hanning", 0.5, "f32"), (16384, "hanning", 0.1, "auto"), (16384, "hanning", 0.1, "f32"),
          (8192, "kaiser", 0.25, "auto"), (2400000, "ones", 0.1, "auto")]


def main():
    for F, win, r, prec in SHAPES:
        d = derive_config(dict(fftSize=F, window=win, samplingRate=2.4e6))
        S = d["fullSize"]
        x = synth.tones_noise(S, seed=2, dtype=np.complex128)
        with Plan(F, S, r, d["theWin"], "AVG", _ffi.IN_C128, precision=prec) as plan:
            for _ in range(5):
                plan.curscan(x)
            n = 200 if F <= 16384 else 10
            t0 = time.perf_counter()
            for _ in range(n):
                out = plan.curscan(x)
            gpu_us = (time.perf_counter() - t0) / n * 1e6
            path, precision = plan.path, plan.precision
        reps = 5 if F <= 16384 else 1
        t0 = time.perf_counter()
        for _ in range(reps):
            ref = O.curscan(x, F, r, d["theWin"], "AVG")
        cpu_us = (time.perf_counter() - t0) / reps * 1e6
        err = float(np.max(np.abs(10 * np.log10(out[ref > 1e-7]) - 10 * np.log10(ref[ref > 1e-7]))))
        print(json.dumps({"fftSize": F, "fullSize": S, "window": win, "nonOverlap": r, "path": path, "precision": precision,
                          "kspec_curscan_us": round(gpu_us, 1), "numpy_one_core_us": round(cpu_us, 1), "speedup": round(cpu_us / gpu_us, 1),
                          "max_db_err_above_1e-7": err}), flush=True)


if __name__ == "__main__":
    main()
