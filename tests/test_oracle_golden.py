"""Pin the oracle (oracle/kspec_oracle.py) to vectors produced by the UNMODIFIED reference
(tests/golden/*, written by oracle/make_golden.py in the build container).  CPU only.

Bar: bit-exact (the restatement performs the same float64 numpy operations in the same order).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from kspec import synth
from oracle import kspec_oracle as O


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def _zerospan_inputs(g):
    p = g["params"]
    if "capture" in g:
        cap = g["capture"].astype(np.complex128)
    else:
        cap = synth.from_u8_iq(g["capture_u8"])
    return p, cap


@pytest.mark.parametrize("name", [
    "g1_zerospan_2048_hanning.npz", "g1b_zerospan_1024_hamming_adj.npz",
    "g3_zerospan_8192_kaiser.npz", "g4_zerospan_32768_ones_max_u8.npz"])
def test_zerospan_golden(name):
    g = load_golden(name)
    p, cap = _zerospan_inputs(g)
    F, S, r = p["fftSize"], p["fullSize"], p["curScanNonOverlap"]
    assert S == O.full_size(F, p["samplingRate"])
    win = O.window_table(p["window"], F)
    assert same(win, g["window"])
    assert same(O.frame_offsets(F, S, r), g["offsets"])
    lin = [O.curscan(cap[k * S:(k + 1) * S], F, r, win, p["curScanCumuMode"]) for k in range(p["nScans"])]
    assert same(lin, g["lin_rows"])
    adj = g.get("adj")
    z = O.zerospan(lin, p["gain"], p["xRes"], p["pltCompressHM"], adj=adj)
    assert same(z["cur_rows"], g["db_rows"])
    assert same(z["max"], g["fft_max"]) and same(z["min"], g["fft_min"]) and same(z["avg"], g["fft_avg"])
    W = O.heatmap_width(F, p["xRes"], p["pltCompressHM"])
    assert g["hm"].shape == (O.HEATMAP_ROWS, W)
    assert same(z["hm_rows"], g["hm"][:p["nScans"]])
    # peak bin of the strongest synthetic tone: +300 kHz at 2.4 MS/s -> shifted bin F/2 + f*F/fs
    if "hanning" in name:
        assert int(np.argmax(lin[0])) == F // 2 + int(round(300e3 * F / 2.4e6))


def _scan_inputs(name, g):
    p = g["params"]
    if "step_bufs" in g:
        bufs = [b.astype(np.complex128) for b in g["step_bufs"]]
    elif "step_bufs_u8" in g:
        bufs = [synth.from_u8_iq(b) for b in g["step_bufs_u8"]]
    else:
        bufs = None
    return p, bufs


@pytest.mark.parametrize("name", ["g2_scan_64_r100.npz", "g2_scan_64_r050.npz", "g5a_fmscan_4096_u8.npz",
                                  "g5b_scan_1200_cur.npz", "g5b_scan_1200_raw.npz"])
def test_scan_golden(name):
    g = load_golden(name)
    p, bufs = _scan_inputs(name, g)
    if bufs is None:                       # "raw" variant shares the inputs of the "cur" fixture
        g0 = load_golden("g5b_scan_1200_cur.npz")
        _, bufs = _scan_inputs(name, g0)
        g["lin_rows"], g["freqs_all"], g["window"] = g0["lin_rows"], g0["freqs_all"], g0["window"]
    F, S, r, R = p["fftSize"], p["fullSize"], p["curScanNonOverlap"], p["scanRangeNonOverlap"]
    win = O.window_table(p["window"], F)
    assert same(win, g["window"])
    geo = O.scan_geometry(p["startFreq"], p["endFreq"], p["samplingRate"], F, R)
    num_groups, total, steps = geo
    assert len(steps) == p["nSteps"] == len(bufs)
    lin = [O.curscan(b[:S], F, r, win, p["curScanCumuMode"]) for b in bufs]
    assert same(lin, g["lin_rows"])
    assert same(O.scan_freq_axis(p["startFreq"], p["samplingRate"], F, num_groups)[:0], g["freqs_all"][:0])
    st = O.scan_init_state(total, p["gain"], p["minAmp4Clip"])
    for ps in range(p["nPass"]):
        ok = [not (ps == 0 and s in p["failSteps"]) for s in range(p["nSteps"])]
        O.scan_pass(lin, ok, geo, p["gain"], st, ps, p["minAmp4Clip"], p["bScanRangeBaseDataIsRaw"])
        for k in ("cur", "max", "min", "avg"):
            assert same(st[k], g["p%d_%s" % (ps, k)]), (ps, k)
        hm_row = O.plotcompress(st["avg"], p["xRes"], p["pltCompressHM"])
        assert same(hm_row, g["hm"][ps])


def test_scan_freq_axis_matches_reference_overwrite():
    """freqsAll is initialised by K:609 and then overwritten slice by slice (K:631-634); both agree to
    float64 rounding, the reference's final array is the per-step one."""
    g = load_golden("g2_scan_64_r100.npz")
    p = g["params"]
    geo = O.scan_geometry(p["startFreq"], p["endFreq"], p["samplingRate"], p["fftSize"], p["scanRangeNonOverlap"])
    ax = O.scan_freq_axis(p["startFreq"], p["samplingRate"], p["fftSize"], geo[0])
    np.testing.assert_allclose(ax, g["freqs_all"], rtol=0, atol=1e-3)


def test_handle_args_table():
    rows = json.load(open(os.path.join(GOLDEN, "g7_handle_args.json")))
    for row in rows:
        a = row["argv"]
        fs = 2.4e6
        assert O.full_size(row["fftSize"], fs) == row["fullSize"], a
        x_in = int(a[a.index("xRes") + 1]) if "xRes" in a else 512
        assert O.adjust_xres(row["fftSize"], x_in) == row["xRes"], a
        if row["prgMode"] == "SCAN":
            if a[0] == "quickFullScan":
                s, e = 30e6, 1.5e9
            elif a[0] == "fmScan":
                s, e = 88e6, 108e6
            else:
                s, e = float(a[a.index("startFreq") + 1]), float(a[a.index("endFreq") + 1])
            s2, e2, c2 = O.fixup_scan_range(s, e, fs)
            assert (s2, e2, c2) == (row["startFreq"], row["endFreq"], row["centerFreq"]), a


def test_known_answers():
    """SURVEY section 4 KATs: bin-centred tone of amplitude A -> 2A linear for every window; tone at
    offset f lands at shifted bin F/2 + f*F/fs; scan init floor = -93.1824 dB; tune failure = -gain."""
    F, fs, A = 2048, 2.4e6, 0.5
    k = 256
    n = np.arange(F * 8)
    x = A * np.exp(2j * np.pi * k * n / F)
    for w in ("ones", "hanning", "hamming", "kaiser"):
        lin = O.curscan(x, F, 0.5, O.window_table(w, F))
        assert int(np.argmax(lin)) == F // 2 + k
        # symmetric (non periodic) numpy windows leak a little off a bin centre: 2A to ~1e-3
        assert abs(lin[F // 2 + k] - 2 * A) < 2e-3 * 2 * A, w
    st = O.scan_init_state(16, 19.1)
    assert abs(st["cur"][0] - (-93.1824)) < 1e-3 and st["min"][0] == -19.1
    assert O.frame_offsets(2048, 16384, 0.1)[:4].tolist() == [0, 204, 409, 614]
    assert len(O.frame_offsets(2048, 16384, 0.5)) == 15
    assert len(O.frame_offsets(64, 512, 0.1)) == 71
    assert len(O.frame_offsets(8192, 65536, 0.25)) == 29
    assert len(O.frame_offsets(2 ** 21, 2 ** 22, 0.1)) == 11
    assert O.sdr_read_plan(4800000)[-1] == (131072, 81408)


@pytest.mark.parametrize("F,S,r,wname", [(2048, 16384, 0.5, "hanning"), (2048, 16384, 0.1, "kaiser"), (64, 512, 0.1, "ones"),
                                           (1000, 8000, 0.25, "hamming")])
def test_psd_restatement_against_scipy_welch(F, S, r, wname):
    """bUsePSD (K:374-384) calls matplotlib, which is neither in the reference tree nor installed: the oracle's
    restatement of mlab.psd is pinned against scipy.signal.welch, an independent implementation of the same estimate
    (density scaling, Fs = 2, no detrend, two-sided, mean over segments)."""
    signal = pytest.importorskip("scipy.signal")
    win = O.window_table(wname, F)
    x = synth.tones_noise(S, seed=11).astype(np.complex128)
    noverlap = int(F * (1 - r))
    _, pxx = signal.welch(x, fs=2.0, window=win, nperseg=F, noverlap=noverlap, nfft=F, detrend=False,
                          return_onesided=False, scaling="density", average="mean")
    got = O.curscan_psd(x, F, r, win)
    assert len(O.psd_segments(F, S, r)) == (S - noverlap) // (F - noverlap)
    assert np.allclose(got, np.fft.fftshift(pxx), rtol=1e-12, atol=0)
    # Parseval: integrating the density over the two-sided band gives the windowed mean power
    k = int(np.argmax(got))
    assert k == int(np.argmax(O.curscan(x, F, r, win)))           # same strongest bin as the magnitude path
