"""kspecanal.py itself as a user of the library (BASELINE.json north_star: "kspecanal.py is a drop-in user of it").

The reference re-binds its module-global ``sdr_curscan`` at run time (kspecanal.py:531,543: zeroSpanPlay swaps in a reader
of pickled spectra).  INTEGRATION.md section 3 uses the same seam: ``sdr_curscan`` is re-bound to a function that returns what
``kspec_curscan`` computes.  The reference exists only in the build container and a GPU only on the B200 box, so the two
halves meet through a fixture: ``tests/golden/r2_dropin_gpu_rows.npz`` holds the float64 rows libkspec.so returned on a
B200 for the captures of the golden fixtures (tools/record_dropin.py), and this test -- CPU, skipped where the reference is
absent -- loads the UNMODIFIED kspecanal.py, applies exactly that re-binding, and runs the reference's own ``zero_span`` and
``_scan_range`` headless on top of the library's output.  Everything downstream of the seam (fftvals_dispproc, data_cumu,
_adj_siglvls, _data_plotcompress, the stitch) is the reference's own code; the results must equal the vectors the
unmodified reference produced with its own numpy ``sdr_curscan`` (tests/golden/g1*, g2*).
"""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from kspec import synth
from oracle import ref_loader

ROWS = os.path.join(GOLDEN, "r2_dropin_gpu_rows.npz")
pytestmark = pytest.mark.skipif(not (ref_loader.available() and os.path.isfile(ROWS)),
                                reason="needs /root/reference (build container) and the recorded library rows")
TOL = 1e-8      # dB: the float64 engines against the reference's float64 (include/kspec.h, KSPEC_PREC_AUTO)


def _rebind(ns, rows):
    """the three-line patch: sdr_curscan(d) keeps the device read (K:370) and returns the library's row for that scan"""
    it = iter(rows)

    def sdr_curscan(d):
        ns["sdr_read"](d["sdr"], d["fullSize"])          # K:370: the device read stays where it is
        return np.array(next(it))

    ns["sdr_curscan"] = sdr_curscan


@pytest.mark.parametrize("name", ["g1_zerospan_2048_hanning", "g1b_zerospan_1024_hamming_adj"])
def test_reference_zero_span_on_library_rows(name):
    g = load_golden(name + ".npz")
    p = g["params"]
    rows = np.load(ROWS)[name + "_rows"]
    holder = {}
    ns = ref_loader.load(sdr_factory=lambda: holder["sdr"])
    argv = ["zeroSpan", "fftSize", p["fftSize"], "window", p["window"].split(".")[-1].lower(), "curScanNonOverlap", p["curScanNonOverlap"],
            "curScanCumuMode", p["curScanCumuMode"].lower(), "xRes", p["xRes"], "prgLoopCnt", p["nScans"], "bPltLevels", "false"]
    d = ref_loader.base_dict(ns, argv)
    assert d["fullSize"] == p["fullSize"]
    d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(g["capture"])
    if "adj" in g:
        d["AdjSigLvls"] = "x"
        d["Fft.Adj"] = g["adj"]
    _rebind(ns, rows)
    with contextlib.redirect_stdout(io.StringIO()):
        ns["zero_span"](d)
    assert np.max(np.abs(d["Fft.Cur"] - g["db_rows"][-1])) < TOL
    assert np.max(np.abs(d["Fft.Max"] - g["fft_max"])) < TOL
    assert np.max(np.abs(d["Fft.Min"] - g["fft_min"])) < TOL
    assert np.max(np.abs(d["Fft.Avg"] - g["fft_avg"])) < TOL
    hm = np.array(d["AxHeatMap"].imshow.call_args[0][0])
    assert np.max(np.abs(hm - g["hm"])) < TOL
    assert np.array_equal(np.argmax(rows, axis=1), np.argmax(g["lin_rows"], axis=1))      # peak bins: bit-exact


@pytest.mark.parametrize("name", ["g2_scan_64_r050", "g2_scan_64_r100"])
def test_reference_scan_range_on_library_rows(name):
    g = load_golden(name + ".npz")
    p = g["params"]
    rows = np.load(ROWS)[name + "_rows"]
    bufs = g["step_bufs"]
    holder = {}
    ns = ref_loader.load(sdr_factory=lambda: holder["sdr"])
    argv = ["scan", "startFreq", p["startFreq"], "endFreq", p["endFreq"], "fftSize", p["fftSize"], "window", p["window"].split(".")[-1].lower(),
            "curScanNonOverlap", p["curScanNonOverlap"], "curScanCumuMode", p["curScanCumuMode"].lower(), "xRes", p["xRes"],
            "scanRangeNonOverlap", p["scanRangeNonOverlap"], "pltCompress", "raw", "bPltLevels", "false"]
    d = ref_loader.base_dict(ns, argv)
    d["pltCompressHM"] = p["pltCompressHM"]
    freqs = ffts = None
    for ps in range(p["nPass"]):
        fails = set(p["failSteps"]) if ps == 0 else set()
        d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(per_tune=lambda t, fc, n: bufs[t], fail_tunes=fails)
        # a failed tune never reaches sdr_curscan (the reference substitutes ones(fftSize), K:637-639)
        _rebind(ns, [r for s, r in enumerate(rows) if s not in fails])
        with contextlib.redirect_stdout(io.StringIO()):
            freqs, ffts = ns["_scan_range"](d, freqs, ffts, ps)
        for k, key in (("cur", "Fft.Cur"), ("max", "Fft.Max"), ("min", "Fft.Min"), ("avg", "Fft.Avg")):
            assert np.max(np.abs(d[key] - g["p%d_%s" % (ps, k)])) < TOL, (ps, k)
        d["fftHMIndex"] = (d["fftHMIndex"] + 1) % d["fftHMMax"]
    assert np.array_equal(freqs, g["freqs_all"])
