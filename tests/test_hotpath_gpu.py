"""The host-side mirror (kspec/hotpath.py: same function names and dict keys as kspecanal.py) driven the way the
reference drives itself -- a fake RtlSdr, zero_span / zero_span_save / zero_span_play / _scan_range -- and compared
with what the UNMODIFIED reference produced for the same inputs (tests/golden)."""
import io
import pickle

import numpy as np
import pytest

from conftest import load_golden
from kspec import hotpath as H
from kspec import synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _d(**kw):
    d = dict(kw)
    d.setdefault("bPltLevels", False)
    return H.derive_config(d)


def test_zero_span_loop_matches_reference():
    g = load_golden("g1_zerospan_2048_hanning.npz")
    p = g["params"]
    d = _d(fftSize=p["fftSize"], window="hanning", curScanNonOverlap=p["curScanNonOverlap"], prgLoopCnt=p["nScans"],
           gain=p["gain"], centerFreq=92e6)
    assert d["fullSize"] == p["fullSize"] and d["xRes"] == p["xRes"]
    d["sdr"] = synth.ArrayRtlSdr(g["capture"])
    assert H.zero_span(d, block=3) == p["nScans"]           # two GPU batches: state carries across them
    assert np.max(np.abs(d["Fft.Max"] - g["fft_max"])) < TOL
    assert np.max(np.abs(d["Fft.Min"] - g["fft_min"])) < TOL
    assert np.max(np.abs(d["Fft.Avg"] - g["fft_avg"])) < TOL
    assert np.max(np.abs(d["Fft.Cur"] - g["db_rows"][-1])) < TOL
    assert d["fftHM"].shape == g["hm"].shape
    assert np.max(np.abs(d["fftHM"] - g["hm"])) < TOL       # the whole 128-row ring, untouched rows stay zero
    assert d["fftHMIndex"] == p["nScans"] % 128
    H.close_plans(d)


def test_sdr_curscan_seam():
    g = load_golden("g1b_zerospan_1024_hamming_adj.npz")
    p = g["params"]
    d = _d(fftSize=1024, window="hamming", curScanCumuMode="MAX", xRes=256)
    d["sdr"] = synth.ArrayRtlSdr(g["capture"])
    for k in range(p["nScans"]):
        out = H.sdr_curscan(d)
        assert out.dtype == np.float64 and out.shape == (1024,)
        assert np.max(np.abs(10 * np.log10(out) - 10 * np.log10(g["lin_rows"][k]))) < TOL
    with pytest.raises(EOFError):
        H.sdr_curscan(d)
    H.close_plans(d)


def test_zero_span_save_and_play_stream_format(tmp_path):
    g = load_golden("g6_zerospansave_64.npz")
    p = g["params"]
    path = str(tmp_path / "z.save")
    d = _d(fftSize=p["fftSize"], centerFreq=p["centerFreq"], samplingRate=p["samplingRate"], prgLoopCnt=p["nScans"],
           zeroSpanSaveFile=path, zeroSpanPlayFile=path)
    d["sdr"] = synth.ArrayRtlSdr(g["capture"])
    clock = iter([1000.0 + 0.25 * i for i in range(100)])
    assert H.zero_span_save(d, block=2, clock=lambda: next(clock)) == p["nScans"]
    mine = open(path, "rb").read()
    ref = g["blob"].tobytes()
    assert len(mine) == len(ref)                            # same framing: header + (time, float64[F]) records
    fa, fb = io.BytesIO(mine), io.BytesIO(ref)
    for _ in range(3):
        assert pickle.load(fa) == pickle.load(fb)           # centerFreq, samplingRate, gain
    for k in range(p["nScans"]):
        assert pickle.load(fa) == pickle.load(fb) == p["times"][k]
        a, b = pickle.load(fa), pickle.load(fb)
        assert a.dtype == b.dtype == np.float64 and a.shape == b.shape
        assert np.max(np.abs(10 * np.log10(a) - 10 * np.log10(b))) < TOL
    # play it back through the mirror of zero_span_play (K:547-564)
    d2 = _d(fftSize=p["fftSize"], zeroSpanPlayFile=path)
    H.zero_span_play_setup(d2)
    assert (d2["centerFreq"], d2["samplingRate"], d2["gain"]) == (p["centerFreq"], p["samplingRate"], p["gain"])
    n = 0
    while True:
        rec = H.zero_span_play(d2)
        if rec is None:
            break
        assert d2["timeWas"] == p["times"][n]
        n += 1
    assert n == p["nScans"] and d2["cmd.stop"] is True
    d2["zeroSpanFile"].close()
    H.close_plans(d)


def test_zero_span_play_matches_reference_loop(tmp_path):
    """zeroSpanSave on the GPU, then zeroSpanPlay on the GPU == the reference's zero_span on the same capture"""
    g = load_golden("g1_zerospan_2048_hanning.npz")
    p = g["params"]
    path = str(tmp_path / "cap.save")
    d = _d(fftSize=p["fftSize"], window="hanning", curScanNonOverlap=p["curScanNonOverlap"], prgLoopCnt=p["nScans"], gain=p["gain"],
           zeroSpanSaveFile=path)
    d["sdr"] = synth.ArrayRtlSdr(g["capture"])
    assert H.zero_span_save(d) == p["nScans"]
    d2 = _d(fftSize=p["fftSize"], zeroSpanPlayFile=path, prgLoopCnt=100)       # window/overlap do not matter for play
    assert H.zero_span_play_all(d2, block=3) == p["nScans"]
    for k, ref in (("Fft.Max", "fft_max"), ("Fft.Min", "fft_min"), ("Fft.Avg", "fft_avg")):
        assert np.max(np.abs(d2[k] - g[ref])) < TOL, k
    assert np.max(np.abs(d2["Fft.Cur"] - g["db_rows"][-1])) < TOL
    assert np.max(np.abs(d2["fftHM"] - g["hm"])) < TOL
    H.close_plans(d)
    H.close_plans(d2)


@pytest.mark.parametrize("name", ["g2_scan_64_r100.npz", "g2_scan_64_r050.npz"])
def test_scan_range_with_tune_failure(name):
    g = load_golden(name)
    p = g["params"]
    bufs = g["step_bufs"]
    d = _d(fftSize=64, startFreq=p["startFreq"], endFreq=p["endFreq"], scanRangeNonOverlap=p["scanRangeNonOverlap"],
           pltCompress="RAW", gain=p["gain"])
    assert d["xRes"] == p["xRes"] == 64
    holder = {}
    freqs = ffts = None
    for ps in range(p["nPass"]):
        fails = set(p["failSteps"]) if ps == 0 else set()
        d["sdr"] = holder["sdr"] = synth.ArrayRtlSdr(per_tune=lambda t, fc, n: bufs[t], fail_tunes=fails)
        freqs, ffts = H._scan_range(d, freqs, ffts, ps, reopen=lambda: holder["sdr"])
        for k, kk in (("Fft.Cur", "cur"), ("Fft.Max", "max"), ("Fft.Min", "min"), ("Fft.Avg", "avg")):
            assert np.max(np.abs(d[k] - g["p%d_%s" % (ps, kk)])) < TOL, (ps, k)
        assert np.max(np.abs(d["fftHM"][d["fftHMIndex"]] - g["hm"][ps])) < TOL
        d["fftHMIndex"] = (d["fftHMIndex"] + 1) % d["fftHMMax"]
    assert np.max(np.abs(freqs - g["freqs_all"])) == 0       # frequency axis: same float64 expressions
    H.close_plans(d)


def test_cli_save_then_play(tmp_path):
    from kspec import cli
    raw = str(tmp_path / "cap.bin")
    synth.to_u8_iq(synth.tones_noise(5 * 16384, seed=3, dtype=np.complex128)).tofile(raw)
    save, out = str(tmp_path / "z.save"), str(tmp_path / "play.npz")
    assert cli.main(["zeroSpanSave", "fftSize", "2048", "window", "hanning", "curScanNonOverlap", "0.5", "iqFile", raw,
                     "zeroSpanSaveFile", save, "prgLoopCnt", "5"]) == 0
    assert cli.main(["zeroSpanPlay", "fftSize", "2048", "zeroSpanPlayFile", save, "outFile", out]) == 0
    z = np.load(out)
    assert z["Fft_Max"].shape == (2048,) and int(z["nScans"]) == 5
    assert int(np.argmax(z["Fft_Max"])) == 1024 + 256            # +300 kHz tone at 2.4 MS/s
    assert z["fftHM"].shape == (128, 512) and np.all(z["fftHM"][5:] == 0)
    assert cli.main(["quickFullScan", "endFreq", "40e6", "scanRangeNonOverlap", "1.0", "prgLoopCnt", "2", "outFile", out]) == 0


def test_unknown_modes_quit_like_the_reference():
    d = _d(fftSize=64, curScanCumuMode="MEDIAN")
    with pytest.raises(SystemExit):
        H.curscan_samples(d, np.zeros(512, dtype=np.complex64))
    assert d["cmd.stop"] is True
    d = _d(fftSize=64)
    with pytest.raises(SystemExit):
        H._data_plotcompress(d, np.zeros(64), "MIN")          # documented but unreachable in the reference (K:188, K:202)
    with pytest.raises(KeyError):
        _d(fftSize=64, window="rectangle")                    # K:936: only ones/hanning/hamming/kaiser exist


def test_sdr_curscan_with_use_psd():
    """bUsePSD true (K:374-384): the seam returns the Welch PSD of the scan instead of the cumulated magnitudes"""
    from oracle import kspec_oracle as O
    x = synth.tones_noise(16384 * 2, seed=21, dtype=np.complex128)
    d = _d(fftSize=2048, window="hanning", curScanNonOverlap=0.1, bUsePSD=True)
    d["sdr"] = synth.ArrayRtlSdr(x)
    for k in range(2):
        out = H.sdr_curscan(d)
        ref = O.curscan_psd(x[k * 16384:(k + 1) * 16384], 2048, 0.1, d["theWin"])
        assert out.shape == (2048,) and np.max(np.abs(10 * np.log10(out) - 10 * np.log10(ref))) < TOL
    H.close_plans(d)
