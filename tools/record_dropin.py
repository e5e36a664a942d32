#!/usr/bin/env python3
"""Record what libkspec.so returns at the reference's operator seam, on a B200, for the captures of the golden fixtures.

    gpurun -- 'python tools/record_dropin.py gpurun_out/dropin_gpu_rows.npz'   ->  tests/golden/r2_dropin_gpu_rows.npz

The reference re-binds its module-global ``sdr_curscan`` at run time (kspecanal.py:531,543).  tests/test_dropin_reference.py
(CPU, where /root/reference exists but no GPU does) loads the UNMODIFIED kspecanal.py, re-binds ``sdr_curscan`` the same way to
replay the rows recorded here -- one kspec_curscan call per scan / per step, float64 engines, exactly what the three-line
patch of INTEGRATION.md section 3 returns -- and runs the reference's own zero_span / _scan_range on top of them.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))

from kspec import _ffi  # noqa: E402
from kspec.engine import Plan  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    return g


def main():
    out = {}
    # zeroSpan fixtures: one kspec_curscan call per scan (K:464)
    for name in ("g1_zerospan_2048_hanning.npz", "g1b_zerospan_1024_hamming_adj.npz"):
        g = load(name)
        p = g["params"]
        cap = g["capture"]
        S = p["fullSize"]
        with Plan(p["fftSize"], S, p["curScanNonOverlap"], g["window"], p["curScanCumuMode"], _ffi.in_format(cap), precision="auto") as plan:
            out[name[:-4] + "_rows"] = np.array([plan.curscan(cap[k * S:(k + 1) * S]) for k in range(p["nScans"])])
    # stepped scans: one kspec_curscan call per step (K:636)
    for name in ("g2_scan_64_r050.npz", "g2_scan_64_r100.npz"):
        g = load(name)
        p = g["params"]
        bufs = g["step_bufs"]
        with Plan(p["fftSize"], p["fullSize"], p["curScanNonOverlap"], g["window"], p["curScanCumuMode"], _ffi.in_format(np.ascontiguousarray(bufs[0])), precision="auto") as plan:
            out[name[:-4] + "_rows"] = np.array([plan.curscan(np.ascontiguousarray(b)) for b in bufs])
    out["provenance"] = np.array("kspec_curscan (KSPEC_PREC_AUTO = float64) through ctypes on an NVIDIA B200, tools/record_dropin.py")
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "dropin_gpu_rows.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
