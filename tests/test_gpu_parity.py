"""GPU parity: libkspec.so (through the ctypes C ABI) against the oracle and the reference's golden vectors.

Bars (BASELINE.json north_star): frame count, frame offsets, fftshift bin order and peak-bin argmax bit-exact;
dB within 1e-3 dB of numpy float64 for bins above the clip floor (minAmp4Clip = 3.9e-8 linear).
"""
import numpy as np
import pytest

from conftest import load_golden
from kspec import _ffi, synth
from kspec.engine import Plan
from oracle import kspec_oracle as O

pytestmark = pytest.mark.gpu

DB_TOL = 1e-3            # dB, the north-star tolerance
F64_TOL = 1e-8           # dB, what the float64 engines actually deliver
# float32 engines, the contract of include/kspec.h (KSPEC_PREC_F32): every bin within F32_ABS of the scan's strongest bin
# (absolute, linear), hence within 1e-3 dB for the bins above F32_DYN of that peak; weaker bins carry no dB guarantee.
# precision "auto"/"f64" has no such limit: 1e-8 dB on every bin.
F32_ABS = 3e-7
F32_DYN = 2e-3
FS = 2.4e6


def db(lin):
    with np.errstate(divide="ignore"):
        return 10 * np.log10(lin)


def assert_db_close(got_lin, ref_lin, tol=DB_TOL, what="", f32=False):
    """compare two linear spectra in dB on the bins above the clip floor (float32: see F32_DYN)"""
    floor = O.MIN_AMP4CLIP
    if f32:
        peak = float(ref_lin.max())
        floor = max(floor, F32_DYN * peak)
        assert float(np.max(np.abs(got_lin - ref_lin))) < F32_ABS * peak, what
    m = ref_lin > floor
    err = np.abs(db(got_lin[m]) - db(ref_lin[m]))
    assert err.size and float(err.max()) < tol, "%s max |dB err| %.3g at %d" % (what, err.max(), int(np.argmax(err)))
    if not f32:      # below the clip floor both must be below it
        assert (got_lin[~m] <= O.MIN_AMP4CLIP * 1.01).all()


def _cap(g):
    if "capture" in g:
        return g["capture"]
    return g["capture_u8"]


# ---------------------------------------------------------------------------------------------------------------
# golden vectors produced by the unmodified reference
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,prec", [
    ("g1_zerospan_2048_hanning.npz", "f32"), ("g1_zerospan_2048_hanning.npz", "f64"),
    ("g1b_zerospan_1024_hamming_adj.npz", "f32"), ("g1b_zerospan_1024_hamming_adj.npz", "f64"),
    ("g3_zerospan_8192_kaiser.npz", "f64"), ("g3_zerospan_8192_kaiser.npz", "auto"),
    ("g4_zerospan_32768_ones_max_u8.npz", "auto"),
])
@pytest.mark.parametrize("frame_parallel", ["1", "0"])
def test_zerospan_golden(name, prec, frame_parallel, monkeypatch):
    # small batches spread the frames of a scan over teams (frame-parallel form); "0" forces the batch form of the kernel
    monkeypatch.setenv("KSPEC_FRAME_PARALLEL", frame_parallel)
    g = load_golden(name)
    p = g["params"]
    F, S, n = p["fftSize"], p["fullSize"], p["nScans"]
    cap = _cap(g)
    with Plan(F, S, p["curScanNonOverlap"], g["window"], p["curScanCumuMode"], _ffi.in_format(cap), precision=prec) as plan:
        assert plan.path == ("smem" if F <= 8192 else "fourstep")
        if prec == "auto":
            assert plan.precision == "f64"
        assert np.array_equal(plan.frame_offsets(), g["offsets"])           # frame count + offsets: bit-exact
        lin = plan.zerospan_batch(cap, n, p["gain"], p["xRes"], p["pltCompressHM"], rows="linear", want_hm=False)["rows"]
        out = plan.zerospan_batch(cap, n, p["gain"], p["xRes"], p["pltCompressHM"], adj=g.get("adj"), rows="db")
        one = plan.curscan(cap[:S * (2 if cap.dtype == np.uint8 else 1)])
    tol = DB_TOL if plan.precision == "f32" else F64_TOL
    for k in range(n):
        assert_db_close(lin[k], g["lin_rows"][k], tol, "scan %d" % k, f32=plan.precision == "f32")
        assert int(np.argmax(lin[k])) == int(np.argmax(g["lin_rows"][k]))   # peak bin: bit-exact
    assert np.max(np.abs(db(one) - db(lin[0]))) < 1e-5
    assert np.max(np.abs(out["rows"] - g["db_rows"])) < tol
    assert np.max(np.abs(out["max"] - g["fft_max"])) < tol
    assert np.max(np.abs(out["min"] - g["fft_min"])) < tol
    assert np.max(np.abs(out["avg"] - g["fft_avg"])) < tol
    assert np.max(np.abs(out["hm_rows"] - g["hm"][:n])) < tol
    assert np.array_equal(np.argmax(out["hm_rows"], axis=1), np.argmax(g["hm"][:n], axis=1))


@pytest.mark.parametrize("name", ["g2_scan_64_r100.npz", "g2_scan_64_r050.npz", "g5a_fmscan_4096_u8.npz",
                                  "g5b_scan_1200_cur.npz", "g5b_scan_1200_raw.npz"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("frame_parallel", ["1", "0"])
def test_scan_golden(name, prec, frame_parallel, monkeypatch):
    monkeypatch.setenv("KSPEC_FRAME_PARALLEL", frame_parallel)
    g = load_golden(name)
    p = g["params"]
    if p["fftSize"] == 1200:
        if prec == "f32":
            pytest.skip("non power-of-two frames run on the float64 Bluestein engine")
        if "step_bufs" not in g:                 # the "raw" fixture shares the inputs of the "cur" one
            g0 = load_golden("g5b_scan_1200_cur.npz")
            g["step_bufs"], g["window"] = g0["step_bufs"], g0["window"]
    F, S = p["fftSize"], p["fullSize"]
    bufs = g["step_bufs"] if "step_bufs" in g else g["step_bufs_u8"]
    samples = np.ascontiguousarray(bufs).reshape(-1)
    geo = O.scan_geometry(p["startFreq"], p["endFreq"], p["samplingRate"], F, p["scanRangeNonOverlap"])
    _, total, steps = geo
    st = O.scan_init_state(total, p["gain"], p["minAmp4Clip"])
    tol = DB_TOL if prec == "f32" else F64_TOL
    with Plan(F, S, p["curScanNonOverlap"], g["window"], p["curScanCumuMode"], _ffi.in_format(samples), precision=prec) as plan:
        for ps in range(p["nPass"]):
            ok = np.array([0 if (ps == 0 and s in p["failSteps"]) else 1 for s in range(p["nSteps"])], dtype=np.uint8)
            plan.scan_batch(samples, p["nSteps"], [s["i_start"] for s in steps], [s["i_done"] for s in steps], total,
                            p["minAmp4Clip"], p["gain"], st, ps, step_ok=ok, base_is_raw=p["bScanRangeBaseDataIsRaw"])
            for k in ("cur", "max", "min", "avg"):
                err = np.max(np.abs(st[k] - g["p%d_%s" % (ps, k)]))
                assert err < tol, (ps, k, err)
            hm = plan.plotcompress(st["avg"], p["xRes"], p["pltCompressHM"])
            assert np.max(np.abs(hm - g["hm"][ps])) < tol
    assert np.array_equal(np.argmax(st["cur"].reshape(-1, F), axis=1), np.argmax(g["p%d_cur" % (p["nPass"] - 1)].reshape(-1, F), axis=1))


@pytest.mark.parametrize("name", ["g2_scan_64_r050.npz", "g5b_scan_1200_raw.npz"])
def test_scan_state_resident_on_device_golden(name):
    """kspec_scan_state_init / kspec_scan_pass / kspec_scan_state_fetch: the same passes with Fft.Cur/Max/Min/Avg kept in HBM
    between passes, against the reference's golden vectors (one fetch per pass here, to compare every pass)"""
    g = load_golden(name)
    p = g["params"]
    if "step_bufs" not in g:
        g0 = load_golden("g5b_scan_1200_cur.npz")
        g["step_bufs"], g["window"] = g0["step_bufs"], g0["window"]
    F, S = p["fftSize"], p["fullSize"]
    samples = np.ascontiguousarray(g["step_bufs"]).reshape(-1)
    geo = O.scan_geometry(p["startFreq"], p["endFreq"], p["samplingRate"], F, p["scanRangeNonOverlap"])
    _, total, steps = geo
    with Plan(F, S, p["curScanNonOverlap"], g["window"], p["curScanCumuMode"], _ffi.in_format(samples), precision="auto") as plan:
        plan.scan_state_init(O.scan_init_state(total, p["gain"], p["minAmp4Clip"]))
        for ps in range(p["nPass"]):
            ok = np.array([0 if (ps == 0 and s in p["failSteps"]) else 1 for s in range(p["nSteps"])], dtype=np.uint8)
            plan.scan_pass(samples, p["nSteps"], [s["i_start"] for s in steps], [s["i_done"] for s in steps], p["minAmp4Clip"], p["gain"], ps,
                           step_ok=ok, base_is_raw=p["bScanRangeBaseDataIsRaw"])
            st = plan.scan_state_fetch()
            for k in ("cur", "max", "min", "avg"):
                assert np.max(np.abs(st[k] - g["p%d_%s" % (ps, k)])) < F64_TOL, (ps, k)


@pytest.mark.parametrize("fmt,prec", [("c64", "f64"), ("u8", "f32")])
def test_scan_pass_chunked_ingest_equals_scan_batch(fmt, prec):
    """a pass large enough for the chunked, overlapped host->device path (330 steps x 32768 samples): device-resident state
    over three passes, fetched once at the end, against kspec_scan_batch pass by pass and against the oracle"""
    F, r, R, gain, n_groups = 4096, 0.5, 0.5, 19.1, 165
    S = O.full_size(F, FS)
    geo = O.scan_geometry(100e6, 100e6 + n_groups * FS, FS, F, R)
    _, total, steps = geo
    ns = len(steps)
    x = np.concatenate([synth.step_tones(s % 40, S, dtype=np.complex128) for s in range(ns)])
    raw = synth.to_u8_iq(x) if fmt == "u8" else x.astype(np.complex64)
    assert raw.nbytes > (64 << 20) or fmt == "u8"
    win = O.window_table("hanning", F)
    i_start, i_done = [s["i_start"] for s in steps], [s["i_done"] for s in steps]
    ok = np.ones(ns, dtype=np.uint8)
    ok[7] = 0
    with Plan(F, S, r, win, "AVG", _ffi.in_format(raw), precision=prec) as plan:
        st = O.scan_init_state(total, gain)
        plan.scan_state_init(st)
        for ps in range(3):
            plan.scan_pass(raw, ns, i_start, i_done, O.MIN_AMP4CLIP, gain, ps, step_ok=ok if ps == 1 else None)
            plan.scan_batch(raw, ns, i_start, i_done, total, O.MIN_AMP4CLIP, gain, st, ps, step_ok=ok if ps == 1 else None)
        res = plan.scan_state_fetch()
        d = plan.dev_alloc(raw.nbytes)
        plan.dev_upload(d, raw)
        plan.scan_pass(d, ns, i_start, i_done, O.MIN_AMP4CLIP, gain, 3, on_device=True)
        plan.scan_batch(raw, ns, i_start, i_done, total, O.MIN_AMP4CLIP, gain, st, 3)
        res_dev = plan.scan_state_fetch(which=("cur", "avg"))
        plan.dev_free(d)
    tol = 1e-4 if prec == "f32" else 1e-9
    assert np.max(np.abs(res_dev["cur"] - st["cur"])) < tol and np.max(np.abs(res_dev["avg"] - st["avg"])) < tol
    if prec == "f64":
        xin = raw.astype(np.complex128)
        lin = [O.curscan(xin[k * S:(k + 1) * S], F, r, win) for k in range(ns)]
        ref = O.scan_init_state(total, gain)
        for ps in range(3):
            ref = O.scan_pass(lin, [bool(v) for v in (ok if ps == 1 else np.ones(ns))], geo, gain, ref, ps)
        for k in ("cur", "max", "min", "avg"):
            assert np.max(np.abs(res[k] - ref[k])) < F64_TOL, k


# ---------------------------------------------------------------------------------------------------------------
# seeded oracle-vs-CUDA sweeps (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2f", [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14])
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("frame_parallel", ["1", "0"])
def test_curscan_all_sizes(log2f, prec, frame_parallel, monkeypatch):
    monkeypatch.setenv("KSPEC_FRAME_PARALLEL", frame_parallel)
    if prec == "f64" and log2f > 13:
        pytest.skip("float64 frames above 8192 use the multi-pass engine")
    F = 1 << log2f
    S = O.full_size(F, FS)
    r, wname, mode = [(0.1, "ones", "AVG"), (0.5, "hanning", "MAX"), (0.25, "hamming", "MIN"), (1.0, "hanning", "RAW")][log2f % 4]
    win = O.window_table(wname, F)
    x = synth.tones_noise(S, seed=100 + log2f, sigma=0.02)
    ref = O.curscan(x.astype(np.complex128), F, r, win, mode)
    with Plan(F, S, r, win, mode, _ffi.IN_C64, precision=prec) as plan:
        assert np.array_equal(plan.frame_offsets(), O.frame_offsets(F, S, r))
        got = plan.curscan(x)
    assert_db_close(got, ref, DB_TOL if prec == "f32" else F64_TOL, "F=%d" % F, f32=prec == "f32")
    assert int(np.argmax(got)) == int(np.argmax(ref))


def expected_big_path(F):
    """power of two -> four-step; 7-smooth above the fused Bluestein kernel's range -> mixed radix; the rest -> Bluestein"""
    if (F & (F - 1)) == 0 and F > 8:
        return "fourstep"
    rest = F
    for q in (2, 3, 5, 7):
        while rest % q == 0:
            rest //= q
    return "mixedradix" if rest == 1 and F > 4096 else "bluestein"


@pytest.mark.parametrize("F,r,wname,mode,fmt", [
    (16384, 0.5, "hanning", "AVG", "c64"),          # float64 frames above 8192: four-step engine (2^14 = 128 x 128)
    (65536, 0.1, "ones", "MAX", "c128"),
    (1 << 18, 0.5, "kaiser", "AVG", "u8"),
    (1 << 16, 0.25, "hamming", "MIN", "u8"),        # tiled column pass (bigfft_kernels.cuh): 256-point columns, tiles of 16
    (1 << 19, 0.5, "hanning", "RAW", "c64"),        # 512-point columns, tiles of 8, 1024-point rows
    (1 << 20, 0.1, "ones", "MAX", "c128"),          # 1024-point columns, tiles of 4
    (1200, 0.1, "hanning", "AVG", "c64"),           # Bluestein, M = 4096 (one fused kernel)
    (1001, 0.5, "hamming", "MIN", "c128"),          # odd length
    (8, 0.5, "ones", "AVG", "c64"),                 # below the fused kernel's minimum
    (3, 1.0, "ones", "MAX", "c64"),
    (5001, 0.25, "kaiser", "RAW", "u8"),            # Bluestein, M = 16384 (multi-pass): 5001 = 3 x 1667
    (48001, 0.5, "hanning", "AVG", "c64"),          # Bluestein, M = 2^17: 48001 = 23 x 2087
    (5000, 0.25, "kaiser", "RAW", "u8"),            # 7-smooth lengths above 4096: two-pass mixed radix, 50 x 100
    (48000, 0.5, "hanning", "AVG", "c64"),          # 200 x 240
    (6174, 0.1, "hamming", "MAX", "c128"),          # 2.3^2.7^3 = 63 x 98: radix 7 and 3
    (30375, 0.5, "ones", "MIN", "c64"),             # 3^5.5^3 = 135 x 225: odd
    (1 << 13 | 1 << 12, 0.5, "hanning", "AVG", "u8"),   # 12288 = 3.2^12 = 96 x 128: radix 4 / 2 lines
])
def test_big_engines(F, r, wname, mode, fmt):
    S = O.full_size(F, FS)
    win = O.window_table(wname, F)
    x = synth.tones_noise(S, seed=F % 97, dtype=np.complex128, sigma=0.02)
    if fmt == "u8":
        raw = synth.to_u8_iq(x)
        xin = synth.from_u8_iq(raw)
    elif fmt == "c64":
        raw = x.astype(np.complex64)
        xin = raw.astype(np.complex128)
    else:
        raw = xin = x
    ref = O.curscan(xin, F, r, win, mode)
    with Plan(F, S, r, win, mode, _ffi.in_format(raw)) as plan:
        assert plan.precision == "f64"
        assert plan.path == expected_big_path(F)
        assert np.array_equal(plan.frame_offsets(), O.frame_offsets(F, S, r))
        got = plan.curscan(raw)
        z = plan.zerospan_batch(np.concatenate([raw, raw]), 2, 19.1, O.adjust_xres(F, 512), "MAX", rows="db")
    assert_db_close(got, ref, F64_TOL, "F=%d" % F)
    assert int(np.argmax(got)) == int(np.argmax(ref))
    refz = O.zerospan([ref, ref], 19.1, O.adjust_xres(F, 512), "MAX")
    for k, kk in (("rows", "cur_rows"), ("hm_rows", "hm_rows"), ("max", "max"), ("min", "min"), ("avg", "avg")):
        assert np.max(np.abs(z[k] - refz[kk])) < F64_TOL, k


def test_fuzz_parameters():
    """60 seeded random combinations of fftSize (power of two and not), fullSize, overlap, window, cumulate mode, ingest
    format, precision, batch size, waterfall mode and xRes: frame offsets bit-exact, every output against the oracle."""
    rng = np.random.default_rng(2024)
    sizes = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 24, 100, 384, 1000, 1536, 3000, 6000, 12000, 32768]
    for case in range(60):
        F = int(rng.choice(sizes))
        pow2 = (F & (F - 1)) == 0
        S = F * int(rng.choice([2, 3, 8]))
        r = float(rng.choice([0.05, 0.1, 0.25, 0.3, 0.5, 0.75, 1.0, 1.5]))
        wname = str(rng.choice(["ones", "hanning", "hamming", "kaiser"]))
        mode = str(rng.choice(["AVG", "MAX", "MIN", "RAW"]))
        fmt = str(rng.choice(["u8", "c64", "c128"]))
        prec = str(rng.choice(["f32", "f64"])) if (pow2 and F <= 4096) else "auto"
        n = int(rng.choice([1, 2, 5, 33]))
        if F >= 6000:
            n = min(n, 2)
        hm_mode = str(rng.choice(["MAX", "AVG", "MIN", "RAW"]))
        xres = O.adjust_xres(F, int(rng.choice([16, 64, 300, 512])))
        if F % xres != 0:            # fftSize < 300 and not a multiple of xRes: the reference itself fails here (K:941-949)
            xres = F
        win = O.window_table(wname, F)
        x = synth.tones_noise(n * S, seed=1000 + case, dtype=np.complex128, sigma=0.05, freqs=(211e3, -640e3, 1.07e6))
        if fmt == "u8":
            raw = synth.to_u8_iq(x)
            xin = synth.from_u8_iq(raw)
        elif fmt == "c64":
            raw = x.astype(np.complex64)
            xin = raw.astype(np.complex128)
        else:
            raw = xin = x
        what = "case %d: F=%d S=%d r=%g %s %s %s %s n=%d hm=%s xres=%d" % (case, F, S, r, wname, mode, fmt, prec, n, hm_mode, xres)
        offs = O.frame_offsets(F, S, r)
        if len(offs) == 0:
            continue
        lin = [O.curscan(xin[k * S:(k + 1) * S], F, r, win, mode) for k in range(n)]
        ref = O.zerospan(lin, 7.5, xres, hm_mode)
        with Plan(F, S, r, win, mode, _ffi.in_format(raw), precision=prec) as plan:
            assert np.array_equal(plan.frame_offsets(), offs), what
            got = plan.zerospan_batch(raw, n, 7.5, xres, hm_mode, rows="linear")
            db_rows = plan.zerospan_batch(raw, n, 7.5, xres, hm_mode, rows="db")["rows"]
            f32 = plan.precision == "f32"
        tol = DB_TOL if f32 else 1e-7
        for k in range(n):
            assert_db_close(got["rows"][k], lin[k], tol, what, f32=f32)
            assert int(np.argmax(got["rows"][k])) == int(np.argmax(lin[k])), what
        if not f32:          # float64: every derived output on every bin
            assert np.max(np.abs(db_rows - ref["cur_rows"])) < tol, what
            assert np.max(np.abs(got["hm_rows"] - ref["hm_rows"])) < tol, what
            for key in ("max", "min", "avg"):
                assert np.max(np.abs(got[key] - ref[key])) < tol, (what, key)


@pytest.mark.parametrize("fmt", ["c64", "u8", "c128"])
def test_large_batch_layout_ragged(fmt):
    """fftSize 2048 float32 batches of >= 1184 scans run the four-teams-per-CTA layout (one CTA per SM): a ragged scan count
    (not a multiple of 4 x 148), carried state, every ingest format; everything against the oracle."""
    F, r, gain, xres = 2048, 0.5, 19.1, 512
    S = O.full_size(F, FS)
    n = 4 * 148 * 2 + 37
    win = O.window_table("hanning", F)
    x = synth.tones_noise(n * S, seed=21, dtype=np.complex128, gate=(300000, 0.5))
    if fmt == "u8":
        raw = synth.to_u8_iq(x)
        xin = synth.from_u8_iq(raw)
    elif fmt == "c64":
        raw = x.astype(np.complex64)
        xin = raw.astype(np.complex128)
    else:
        raw = xin = x
    per = 2 * S if fmt == "u8" else S
    with Plan(F, S, r, win, "AVG", _ffi.in_format(raw), precision="f32") as plan:
        first = plan.zerospan_batch(raw[:40 * per], 40, gain, xres, "MAX")                       # small batch: base layout
        got = plan.zerospan_batch(raw[40 * per:], n - 40, gain, xres, "MAX", rows="db",
                                  state=(first["max"], first["min"], first["avg"]))              # large batch: four teams per CTA
        # plan.info describes what the largest batches run: the 32 x 2 x 32 layout (uint8 / complex64 ingest) or four teams
        if fmt != "c128":
            assert plan.info.cta_threads == 384 and plan.info.scans_per_cta == 6
    lin = [O.curscan(xin[k * S:(k + 1) * S], F, r, win) for k in range(n)]
    ref = O.zerospan(lin, gain, xres, "MAX")
    assert np.max(np.abs(got["rows"] - ref["cur_rows"][40:])) < DB_TOL
    assert np.array_equal(np.argmax(got["rows"], axis=1), np.argmax(ref["cur_rows"][40:], axis=1))
    assert np.max(np.abs(got["hm_rows"] - ref["hm_rows"][40:])) < DB_TOL
    for k in ("max", "min", "avg"):
        assert np.max(np.abs(got[k] - ref[k])) < DB_TOL, k


@pytest.mark.parametrize("fmt,cumu,r,wname", [
    ("c64", "AVG", 0.5, "hanning"), ("u8", "AVG", 0.5, "hanning"), ("c64", "MAX", 0.25, "kaiser"), ("u8", "MIN", 0.5, "hamming"),
    ("c64", "RAW", 1.0, "ones"), ("c64", "AVG", 0.1, "hanning"), ("u8", "AVG", 0.1, "ones"),
])
def test_r32_layout_ragged(fmt, cumu, r, wname):
    """fftSize 2048 float32 batches of >= 1776 scans run the 32 x 2 x 32 layout (curscan_r32.cuh: six 64-thread teams per SM,
    one shared-memory exchange per frame): ragged scan count, carried state, every cumulate mode, even and odd frame offsets
    (r = 0.1: 0, 204, 409, ...), both ingest formats; every output against the oracle."""
    F, gain, xres = 2048, 19.1, 512
    S = O.full_size(F, FS)
    n = 6 * 148 * 2 + 41
    win = O.window_table(wname, F)
    x = synth.tones_noise(n * S, seed=31, dtype=np.complex128, gate=(300000, 0.5))
    if fmt == "u8":
        raw = synth.to_u8_iq(x)
        xin = synth.from_u8_iq(raw)
    else:
        raw = x.astype(np.complex64)
        xin = raw.astype(np.complex128)
    per = 2 * S if fmt == "u8" else S
    with Plan(F, S, r, win, cumu, _ffi.in_format(raw), precision="f32") as plan:
        assert plan.info.cta_threads == 384 and plan.info.scans_per_cta == 6
        first = plan.zerospan_batch(raw[:3 * per], 3, gain, xres, "MAX")                         # small batch: base layout
        got = plan.zerospan_batch(raw[3 * per:], n - 3, gain, xres, "MAX", rows="db",
                                  state=(first["max"], first["min"], first["avg"]))              # large batch: 32 x 2 x 32
        lin = plan.zerospan_batch(raw[3 * per:], n - 3, gain, xres, "AVG", rows="linear")
    ref_lin = [O.curscan(xin[k * S:(k + 1) * S], F, r, win, cumu) for k in range(n)]
    ref = O.zerospan(ref_lin, gain, xres, "MAX")
    rl = np.asarray(ref_lin[3:])
    # the float32 contract (include/kspec.h, KSPEC_PREC_F32): every bin within 3e-7 of the scan's peak, hence within 1e-3 dB
    # wherever the bin is above 2e-3 of the peak; argmax bit-exact
    assert np.max(np.abs(lin["rows"] - rl)) < F32_ABS * rl.max()
    m = rl > F32_DYN * rl.max(axis=1, keepdims=True)
    assert m.any() and np.max(np.abs(got["rows"][m] - ref["cur_rows"][3:][m])) < DB_TOL
    assert np.array_equal(np.argmax(got["rows"], axis=1), np.argmax(ref["cur_rows"][3:], axis=1))
    assert np.max(np.abs(got["max"] - ref["max"])) < (DB_TOL if cumu != "MIN" else 0.05)
    if (cumu, r, wname) != ("AVG", 0.5, "hanning"):
        return
    # the benchmark workload (cumulate AVG over 15 frames averages the rounding noise): EVERY bin of every output within 1e-3 dB
    assert np.max(np.abs(got["rows"] - ref["cur_rows"][3:])) < DB_TOL
    assert np.max(np.abs(got["hm_rows"] - ref["hm_rows"][3:])) < DB_TOL
    for k in ("max", "min", "avg"):
        assert np.max(np.abs(got[k] - ref[k])) < DB_TOL, k
    ref2 = O.zerospan(ref_lin[3:], gain, xres, "AVG")
    assert np.max(np.abs(lin["hm_rows"] - ref2["hm_rows"])) < DB_TOL


def test_r32_static_and_ticket_schedules_agree():
    """the 32 x 2 x 32 kernel hands out scans by a static stride (full, undisturbed launches) or by tickets from a global
    counter (ragged batches, shards of a multi-GPU capture): the same scans must give bit-identical rows either way"""
    F, r, gain, xres = 2048, 0.5, 19.1, 512
    S = O.full_size(F, FS)
    full, ragged = 6 * 148 * 2, 6 * 148 * 2 + 300
    x = synth.tones_noise(ragged * S, seed=33)
    win = O.window_table("hanning", F)
    with Plan(F, S, r, win, "AVG", _ffi.IN_C64, precision="f32") as plan:
        a = plan.zerospan_batch(x[:full * S], full, gain, xres, "MAX", rows="db")                    # 2.0 waves: static stride
        b = plan.zerospan_batch(x, ragged, gain, xres, "MAX", rows="db")                             # 2.17 waves: tickets
        c = plan.zerospan_batch(x[:full * S], full, gain, xres, "MAX", rows="db", scan_index_base=0, n_scans_total=4 * full)   # shard: tickets
    assert np.array_equal(a["rows"], b["rows"][:full]) and np.array_equal(a["hm_rows"], b["hm_rows"][:full])
    assert np.array_equal(a["rows"], c["rows"]) and np.array_equal(a["max"], c["max"]) and np.array_equal(a["min"], c["min"])
    ref = O.log_nogain(O.curscan(x[(ragged - 1) * S:ragged * S].astype(np.complex128), F, r, win), gain)
    assert np.max(np.abs(b["rows"][-1] - ref)) < DB_TOL


def test_cfg2_full_size_quickfullscan():
    """BASELINE cfg 2 at full size: quickFullScan 30 MHz..1.5 GHz -> 613 groups, 39 232 entries, fftSize 64, ones, r = 0.1
    (71 frames per step); 613 steps at scanRangeNonOverlap 1.0 and 1226 at the alias' default 0.5; two passes; vs the oracle."""
    F, r = 64, 0.1
    S = O.full_size(F, FS)
    win = np.ones(F)
    start, end, _ = O.fixup_scan_range(30e6, 1.5e9, FS)
    assert end == 1.5012e9
    for R, n_steps in ((1.0, 613), (0.5, 1226)):
        geo = O.scan_geometry(start, end, FS, F, R)
        num_groups, total, steps = geo
        assert (num_groups, total, len(steps)) == (613, 39232, n_steps)
        bufs = np.concatenate([synth.step_tones(s, S) for s in range(n_steps)])
        lin = [O.curscan(bufs[s * S:(s + 1) * S].astype(np.complex128), F, r, win) for s in range(n_steps)]
        ref = O.scan_init_state(total, 19.1)
        st = O.scan_init_state(total, 19.1)
        ok = np.ones(n_steps, dtype=np.uint8)
        ok[[5, 600]] = 0
        with Plan(F, S, r, win, "AVG", precision="f32") as plan:
            for ps in range(2):
                O.scan_pass(lin, ok.astype(bool), geo, 19.1, ref, ps)
                plan.scan_batch(bufs, n_steps, [s["i_start"] for s in steps], [s["i_done"] for s in steps], total,
                                O.MIN_AMP4CLIP, 19.1, st, ps, step_ok=ok)
                for k in ("cur", "max", "min", "avg"):
                    assert np.max(np.abs(st[k] - ref[k])) < DB_TOL, (R, ps, k)
        assert np.array_equal(np.argmax(st["cur"].reshape(-1, F), axis=1), np.argmax(ref["cur"].reshape(-1, F), axis=1))


def test_cfg3_full_size_sixty_second_capture():
    """BASELINE cfg 3 at full size: 60 s at 2.4 MS/s = 144 M samples -> 2197 scans x 29 frames, fftSize 8192, kaiser(64),
    75 % overlap.  Size-independent properties: identical results for 1 and 4 scan-range shards (the multi-GPU contract),
    sampled scans against the oracle, Max >= Avg >= Min."""
    F, r, gain, xres = 8192, 0.25, 19.1, 512
    S = O.full_size(F, FS)
    n = int(60 * FS) // S
    assert (S, n) == (65536, 2197)
    win = O.window_table("kaiser", F)
    block = synth.tones_noise(64 * S, seed=3)                       # 64-scan block, re-used with a slow gain pattern
    gains = (0.25 + 0.75 * ((np.arange(n) // 64) % 4) / 3.0).astype(np.float32)
    x = np.empty(n * S, dtype=np.complex64)
    for k in range(0, n, 64):
        m = min(64, n - k)
        x[k * S:(k + m) * S] = block[:m * S] * gains[k]
    with Plan(F, S, r, win, "AVG") as plan:
        assert plan.precision == "f64" and plan.n_frames == 29
        full = plan.zerospan_batch(x, n, gain, xres, "MAX", rows="db")
        bounds = np.linspace(0, n, 5).astype(int)
        parts = [plan.zerospan_batch(x[a * S:b * S], b - a, gain, xres, "MAX", scan_index_base=a, n_scans_total=n)
                 for a, b in zip(bounds[:-1], bounds[1:])]
    assert np.array_equal(np.max([p["max"] for p in parts], axis=0), full["max"])
    assert np.array_equal(np.min([p["min"] for p in parts], axis=0), full["min"])
    assert np.max(np.abs(np.sum([p["avg"] for p in parts], axis=0) - full["avg"])) < 1e-9
    assert np.array_equal(np.vstack([p["hm_rows"] for p in parts]), full["hm_rows"])
    assert np.all(full["max"] >= full["avg"] - 1e-12) and np.all(full["avg"] >= full["min"] - 1e-12)
    for s in (0, 1000, 2196):
        ref = O.log_nogain(O.curscan(x[s * S:(s + 1) * S].astype(np.complex128), F, r, win), gain)
        assert np.max(np.abs(full["rows"][s] - ref)) < F64_TOL
        assert int(np.argmax(full["rows"][s])) == int(np.argmax(ref))
    # Avg after 2197 scans == the halving recurrence over the last rows (older rows weigh < 2^-64)
    a = full["rows"][n - 80]
    for row in full["rows"][n - 79:]:
        a = (a + row) / 2
    assert np.max(np.abs(a - full["avg"])) < 1e-9


def test_cfg4_full_size_two_to_the_21():
    """BASELINE cfg 4: fftSize 2^21, ones window, cumulate MAX, curScanNonOverlap 0.1 -> 11 frames per 2^22-sample scan"""
    F, r = 1 << 21, 0.1
    S = O.full_size(F, FS)
    assert S == 1 << 22
    win = np.ones(F)
    x = synth.tones_noise(S, seed=4)
    ref = O.curscan(x.astype(np.complex128), F, r, win, "MAX")
    with Plan(F, S, r, win, "MAX", _ffi.IN_C64) as plan:
        assert plan.path == "fourstep" and plan.n_frames == 11
        assert np.array_equal(plan.frame_offsets(), O.frame_offsets(F, S, r))
        got = plan.curscan(x)
    assert_db_close(got, ref, F64_TOL, "2^21")
    assert int(np.argmax(got)) == int(np.argmax(ref)) == F // 2 + int(round(300e3 * F / FS))


@pytest.mark.parametrize("F,fmt", [(1 << 16, "c64"), (1 << 18, "c128"), (1 << 21, "u8")])
def test_fourstep_tiled_column_pass_matches_the_elementwise_one(F, fmt, monkeypatch):
    """cols_tiled_kernel (default for 256..1024-point columns) against team_fft_kernel<OpColsIn> (KSPEC_FOURSTEP_TILED=0, read
    when the plan is created): same spectra to rounding (the tiled pass factors the twiddle into two table entries), three
    scans so that the slabs of several frames and scans are in flight"""
    S = O.full_size(F, FS)
    win = O.window_table("kaiser", F)
    x = synth.tones_noise(3 * S, seed=61, dtype=np.complex128, sigma=0.02)
    raw = synth.to_u8_iq(x) if fmt == "u8" else (x.astype(np.complex64) if fmt == "c64" else x)
    out = {}
    for tiled in ("1", "0"):
        monkeypatch.setenv("KSPEC_FOURSTEP_TILED", tiled)
        with Plan(F, S, 0.25, win, "AVG", _ffi.in_format(raw)) as plan:
            assert plan.path == "fourstep"
            out[tiled] = plan.zerospan_batch(raw, 3, 19.1, O.adjust_xres(F, 512), "MAX", rows="db")
    for k in ("rows", "hm_rows", "max", "min", "avg"):
        assert np.max(np.abs(out["1"][k] - out["0"][k])) < 1e-9, k
    assert np.array_equal(np.argmax(out["1"]["rows"], axis=1), np.argmax(out["0"]["rows"], axis=1))


@pytest.mark.parametrize("engine", ["mixedradix", "bluestein"])
def test_cfg5b_full_size_2400000(engine, monkeypatch):
    """BASELINE cfg 5b: fftSize 2 400 000 (= one second at 2.4 MS/s): the two-pass mixed-radix engine (1500 x 1600,
    default) and the Bluestein engine BASELINE names (M = 2^23, KSPEC_FORCE_BLUESTEIN=1)"""
    if engine == "bluestein":
        monkeypatch.setenv("KSPEC_FORCE_BLUESTEIN", "1")
    F, r = 2400000, 0.5
    S = O.full_size(F, FS)
    assert S == 4800000
    win = np.ones(F)
    x = synth.tones_noise(S, seed=5)
    ref = O.curscan(x.astype(np.complex128), F, r, win, "AVG")
    with Plan(F, S, r, win, "AVG", _ffi.IN_C64) as plan:
        assert plan.path == engine and plan.info.conv_size == (1 << 23 if engine == "bluestein" else 0)
        got = plan.curscan(x)
    assert_db_close(got, ref, 1e-6 if engine == "bluestein" else F64_TOL, "2.4e6")
    assert int(np.argmax(got)) == int(np.argmax(ref)) == F // 2 + 300000


def test_cfg5b_full_fmscan_geometry():
    """BASELINE cfg 5b at full size: fmScan 88..108 MHz -> 109.6 MHz, 9 groups, 18 steps at scanRangeNonOverlap 0.5,
    fftSize 2 400 000 (mixed radix 1500 x 1600), 21.6 M stitched entries.  numpy needs ~10 s per step here, so the FFT itself is
    pinned by the single-scan test above and THIS test pins the rest at full size: the batched scan (clip, dB, stitch,
    Max/Min/Avg over 21.6 M entries) must equal the oracle's stitch applied to the per-step spectra."""
    F, r, R, gain = 2400000, 0.5, 0.5, 19.1
    S = O.full_size(F, FS)
    start, end, _ = O.fixup_scan_range(88e6, 108e6, FS)
    geo = O.scan_geometry(start, end, FS, F, R)
    num_groups, total, steps = geo
    assert (num_groups, total, len(steps), S) == (9, 21600000, 18, 4800000)
    win = np.ones(F)
    base = synth.tones_noise(S, seed=55)
    n = len(steps)
    x = np.empty(n * S, dtype=np.complex64)
    for s in range(n):                       # cheap per-step variation: a step-dependent tone on top of a common block
        k = 100000 + 37000 * s
        x[s * S:(s + 1) * S] = base * np.float32(0.5 + 0.05 * s)
        x[s * S:(s + 1) * S] += (0.9 * np.exp(2j * np.pi * ((k * np.arange(S)) % F) / F)).astype(np.complex64)
    st = O.scan_init_state(total, gain)
    ref = O.scan_init_state(total, gain)
    with Plan(F, S, r, win, "AVG", _ffi.IN_C64) as plan:
        assert plan.path == "mixedradix" and plan.n_frames == 3
        lin = plan.zerospan_batch(x, n, gain, O.adjust_xres(F, 512), "MAX", rows="linear", want_hm=False)["rows"]
        plan.scan_batch(x, n, [s_["i_start"] for s_ in steps], [s_["i_done"] for s_ in steps], total, O.MIN_AMP4CLIP, gain, st, 0)
        hm = plan.plotcompress(st["avg"], O.adjust_xres(F, 512), "MAX")
    for s in range(n):
        assert int(np.argmax(lin[s])) == F // 2 + 100000 + 37000 * s          # the step's own tone, bin-exact
    O.scan_pass(lin, [True] * n, geo, gain, ref, 0)
    for k in ("cur", "max", "min", "avg"):
        assert np.max(np.abs(st[k] - ref[k])) < 1e-9, k
    assert hm.shape == (300,) and np.allclose(hm, O.plotcompress(ref["avg"], 300, "MAX"), atol=1e-9)


@pytest.mark.parametrize("fmt", ["u8", "c64", "c128"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_ingest_formats(fmt, prec):
    F, r = 512, 0.3
    S = O.full_size(F, FS)
    win = O.window_table("kaiser", F)
    x = synth.tones_noise(S, seed=5, dtype=np.complex128, sigma=0.05)
    if fmt == "u8":
        raw = synth.to_u8_iq(x)
        xin = synth.from_u8_iq(raw)
    elif fmt == "c64":
        raw = x.astype(np.complex64)
        xin = raw.astype(np.complex128)
    else:
        raw = xin = x
    ref = O.curscan(xin, F, r, win, "AVG")
    with Plan(F, S, r, win, "AVG", _ffi.in_format(raw), precision=prec) as plan:
        got = plan.curscan(raw)
    assert_db_close(got, ref, DB_TOL if prec == "f32" else F64_TOL, fmt, f32=prec == "f32")


def test_u8_scale_and_offset_are_parameters():
    """pyrtlsdr's byte->float convention is unpinned by the reference; (b-127)/128 (kspecanal.old.py:126-135) must work too."""
    F, r = 256, 0.5
    S = O.full_size(F, FS)
    win = O.window_table("hanning", F)
    raw = synth.to_u8_iq(synth.tones_noise(S, seed=9, dtype=np.complex128))
    xin = synth.from_u8_iq(raw, offset=127.0, scale=1 / 128.0)
    ref = O.curscan(xin, F, r, win, "AVG")
    with Plan(F, S, r, win, "AVG", _ffi.IN_U8_IQ, precision="f64", u8_offset=127.0, u8_scale=1 / 128.0) as plan:
        got = plan.curscan(raw)
    assert_db_close(got, ref, 1e-8)


@pytest.mark.parametrize("n_scans", [1, 3, 17, 64, 65, 200])
def test_zerospan_batch_sizes_and_avg_window(n_scans):
    """ragged batch sizes (fewer scans than teams, not a multiple of the grid) and the 64-row Avg window"""
    F, r, gain, xres = 128, 0.5, 19.1, 32
    S = O.full_size(F, FS)
    win = O.window_table("hanning", F)
    x = synth.tones_noise(n_scans * S, seed=n_scans, gate=(3000, 0.4))
    lin = [O.curscan(x[k * S:(k + 1) * S].astype(np.complex128), F, r, win) for k in range(n_scans)]
    ref = O.zerospan(lin, gain, xres, "AVG")
    with Plan(F, S, r, win, "AVG", precision="f64") as plan:
        got = plan.zerospan_batch(x, n_scans, gain, xres, "AVG", rows="db")
    for k in ("max", "min", "avg"):
        assert np.max(np.abs(got[k] - ref[k])) < 1e-8, k
    assert np.max(np.abs(got["rows"] - ref["cur_rows"])) < 1e-8
    assert np.max(np.abs(got["hm_rows"] - ref["hm_rows"])) < 1e-8


def test_zerospan_carry_equals_one_batch():
    """state carried across batches == one long batch == the reference loop"""
    F, r, gain, xres = 1024, 0.1, 10.0, 256
    S = O.full_size(F, FS)
    win = O.window_table("hamming", F)
    n = 12
    x = synth.tones_noise(n * S, seed=42, gate=(20000, 0.5))
    lin = [O.curscan(x[k * S:(k + 1) * S].astype(np.complex128), F, r, win) for k in range(n)]
    ref = O.zerospan(lin, gain, xres, "MAX")
    with Plan(F, S, r, win, "AVG", precision="f64") as plan:
        a = plan.zerospan_batch(x[:5 * S], 5, gain, xres, "MAX")
        b = plan.zerospan_batch(x[5 * S:], 7, gain, xres, "MAX", state=(a["max"], a["min"], a["avg"]))
    for k in ("max", "min", "avg"):
        assert np.max(np.abs(b[k] - ref[k])) < 1e-8, k
    assert np.max(np.abs(np.vstack([a["hm_rows"], b["hm_rows"]]) - ref["hm_rows"])) < 1e-8


@pytest.mark.parametrize("n_shards", [2, 4, 8])
def test_sharded_partials_combine_to_the_single_gpu_result(n_shards):
    """the multi-GPU contract on one device: MAX/MIN/SUM over per-shard partials == unsharded result"""
    F, r, gain, xres = 256, 0.5, 19.1, 64
    S = O.full_size(F, FS)
    win = O.window_table("hanning", F)
    n = 40
    x = synth.tones_noise(n * S, seed=77, gate=(5000, 0.5))
    with Plan(F, S, r, win, "AVG", precision="f64") as plan:
        full = plan.zerospan_batch(x, n, gain, xres, "MAX")
        bounds = np.linspace(0, n, n_shards + 1).astype(int)
        parts = [plan.zerospan_batch(x[a * S:b * S], b - a, gain, xres, "MAX", scan_index_base=a, n_scans_total=n)
                 for a, b in zip(bounds[:-1], bounds[1:])]
    assert np.array_equal(np.max([p["max"] for p in parts], axis=0), full["max"])
    assert np.array_equal(np.min([p["min"] for p in parts], axis=0), full["min"])
    assert np.max(np.abs(np.sum([p["avg"] for p in parts], axis=0) - full["avg"])) < 1e-9
    assert np.array_equal(np.vstack([p["hm_rows"] for p in parts]), full["hm_rows"])


def test_silence_gives_minus_inf_in_zerospan_and_floor_in_scan():
    """K:469: zeroSpan has no low clip (-inf survives); K:640-641: scan clips to minAmp4Clip"""
    F, r = 64, 0.5
    S = O.full_size(F, FS)
    win = np.ones(F)
    x = np.zeros(2 * S, dtype=np.complex64)
    with Plan(F, S, r, win, "AVG", precision="f32") as plan:
        z = plan.zerospan_batch(x, 2, 19.1, 64, "RAW", rows="db")
        assert np.all(np.isneginf(z["rows"])) and np.all(np.isneginf(z["avg"]))
        st = O.scan_init_state(2 * F, 19.1)
        plan.scan_batch(x, 2, [0, F], [F, 2 * F], 2 * F, O.MIN_AMP4CLIP, 19.1, st, 0)
        assert np.allclose(st["cur"], 10 * np.log10(O.MIN_AMP4CLIP) - 19.1, atol=1e-4)


def test_scan_base_is_raw_and_quarter_overlap():
    g = load_golden("g5b_scan_1200_cur.npz")           # geometry only; F=1200 itself needs the Bluestein engine
    F, R, fs = 1024, 0.25, FS
    start, end = 100e6, 100e6 + 3 * fs
    geo = O.scan_geometry(start, end, fs, F, R)
    _, total, steps = geo
    S = O.full_size(F, fs)
    win = O.window_table("hanning", F)
    bufs = [synth.step_tones(s + 100, S) for s in range(len(steps))]
    lin = [O.curscan(b.astype(np.complex128), F, 0.1, win) for b in bufs]
    for raw in (False, True):
        ref = O.scan_init_state(total, 19.1)
        st = O.scan_init_state(total, 19.1)
        with Plan(F, S, 0.1, win, "AVG", precision="f64") as plan:
            for ps in range(2):
                O.scan_pass(lin, [True] * len(steps), geo, 19.1, ref, ps, base_is_raw=raw)
                plan.scan_batch(np.concatenate(bufs), len(steps), [s["i_start"] for s in steps], [s["i_done"] for s in steps],
                                total, O.MIN_AMP4CLIP, 19.1, st, ps, base_is_raw=raw)
                for k in ("cur", "max", "min", "avg"):
                    assert np.max(np.abs(st[k] - ref[k])) < 1e-8, (raw, ps, k)
    assert g["params"]["fftSize"] == 1200


@pytest.mark.parametrize("R,n_shards", [(1.0, 3), (0.5, 2), (0.5, 5), (0.25, 4)])
def test_scan_sharded_by_step_equals_unsharded(R, n_shards):
    """the multi-GPU contract of the stepped scan on one device: SUM of the shards' stitch partials == Fft.Cur of the
    unsharded pass, then Max/Min/Avg from it == the reference loop (two passes, one failed tune)"""
    from kspec.sharding import shard_bounds
    F, r, gain = 256, 0.1, 19.1
    S = O.full_size(F, FS)
    start, end = 100e6, 100e6 + 7 * FS
    geo = O.scan_geometry(start, end, FS, F, R)
    _, total, steps = geo
    n = len(steps)
    win = O.window_table("hanning", F)
    bufs = np.concatenate([synth.step_tones(s + 7, S) for s in range(n)])
    lin = [O.curscan(bufs[s * S:(s + 1) * S].astype(np.complex128), F, r, win) for s in range(n)]
    ok = np.ones(n, dtype=np.uint8)
    ok[2] = 0
    ref = O.scan_init_state(total, gain)
    st = O.scan_init_state(total, gain)
    i_start = [s["i_start"] for s in steps]
    with Plan(F, S, r, win, "AVG", precision="f64") as plan:
        for ps in range(2):
            O.scan_pass(lin, ok.astype(bool), geo, gain, ref, ps)
            parts = [plan.scan_shard(bufs[a * S:b * S], b - a, a, i_start, total, O.MIN_AMP4CLIP, gain, step_ok=ok[a:b])
                     for a, b in shard_bounds(n, n_shards) if b > a]
            cur = np.sum(parts, axis=0)                      # what kspec_comm_allreduce_sum does over NVLink
            plan.scan_stats_update(cur, steps[-1]["i_done"], ps, st)
            for k in ("cur", "max", "min", "avg"):
                assert np.max(np.abs(st[k] - ref[k])) < 1e-9, (ps, k)


@pytest.mark.parametrize("mode", ["MAX", "AVG", "MIN", "RAW"])
def test_plotcompress(mode):
    y = np.random.default_rng(0).normal(size=36864)
    with Plan(64, 512, 0.5, np.ones(64)) as plan:
        got = plan.plotcompress(y, 512, mode)
    assert np.allclose(got, O.plotcompress(y, 512, mode), rtol=0, atol=1e-12)


def _plot_highs_ref(freqs, levels, num, delta_frac):
    """the marking loop of plot_highs (K:246-265) without the matplotlib calls"""
    delta = delta_frac * (freqs[-1] - freqs[0])
    order = levels.argsort()
    marked, out = [], []
    for i in np.arange(-1, -len(freqs), -1):
        f = freqs[order[i]]
        if not any(abs(m - f) < delta for m in marked):
            marked.append(f)
            out.append(int(order[i]))
            if len(out) >= num:
                break
    return out


def test_plot_highs_and_conv_display_mode():
    rng = np.random.default_rng(3)
    freqs = np.linspace(88e6, 108e6, 512)
    levels = rng.normal(size=512) * 3 - 60
    levels[[40, 41, 43, 300, 301, 480]] += 40 + np.arange(6)           # clustered peaks: the spacing rule must skip neighbours
    with Plan(64, 512, 0.5, np.ones(64)) as plan:
        for num, frac in ((5, 0.025), (12, 0.01), (3, 0.2), (64, 0.0)):
            assert plan.plot_highs(freqs, levels, num, frac).tolist() == _plot_highs_ref(freqs, levels, num, frac)
        taps = np.kaiser(128, 64)                                      # DataProcConv, K:87
        got = plan.conv_smooth(levels, taps)
    ref = np.convolve(levels, taps, mode="same")                       # K:113-120
    avg = np.average(ref)
    ref[:12] = avg
    ref[-12:] = avg
    assert np.max(np.abs(got - ref)) < 1e-9


def test_known_answer_tone_every_window():
    """SURVEY section 4 KAT: bin-centred tone of amplitude A -> 2A linear (10log10(2A)-gain dB) at bin F/2+k"""
    F, k, A = 2048, 300, 0.25
    S = O.full_size(F, FS)
    n = np.arange(S)
    x = (A * np.exp(2j * np.pi * k * n / F)).astype(np.complex64)
    for w in ("ones", "hanning", "hamming", "kaiser"):
        win = O.window_table(w, F)
        with Plan(F, S, 0.5, win, "AVG", precision="f32") as plan:
            got = plan.curscan(x)
        assert int(np.argmax(got)) == F // 2 + k
        assert abs(got[F // 2 + k] - 2 * A) < 2e-3 * 2 * A, w


# ---------------------------------------------------------------------------------------------------------------
# BASELINE sizes through size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_cfg1_properties():
    """1 s capture at 2.4 MS/s (146 scans, 2190 frames), F=2048 hanning 50% overlap: sampled scans against the
    oracle, linearity (input x2 -> +3.0103 dB), idempotence (same input twice -> identical bits)."""
    F, r, gain, xres = 2048, 0.5, 19.1, 512
    S = O.full_size(F, FS)
    n = int(2.4e6) // S
    assert n == 146
    win = O.window_table("hanning", F)
    x = synth.tones_noise(n * S, seed=1)
    with Plan(F, S, r, win, "AVG", precision="f32") as plan:
        assert plan.n_frames == 15
        a = plan.zerospan_batch(x, n, gain, xres, "MAX", rows="db")
        b = plan.zerospan_batch(x, n, gain, xres, "MAX", rows="db")
        c = plan.zerospan_batch((2 * x).astype(np.complex64), n, gain, xres, "MAX", rows="db")
    for k in ("rows", "hm_rows", "max", "min", "avg"):
        assert np.array_equal(a[k], b[k]), k
        assert np.max(np.abs(c[k] - a[k] - 10 * np.log10(2.0))) < 1e-4, k
    for s in (0, 73, 145):
        ref = O.log_nogain(O.curscan(x[s * S:(s + 1) * S].astype(np.complex128), F, r, win), gain)
        assert np.max(np.abs(a["rows"][s] - ref)) < DB_TOL
        assert int(np.argmax(a["rows"][s])) == int(np.argmax(ref)) == F // 2 + 256
    assert a["hm_rows"].shape == (n, xres)


# ---------------------------------------------------------------------------------------------------------------
# bUsePSD (K:374-384): Welch PSD through the same engines (row N4 of SURVEY 8f)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,r,wname,fmt,path", [
    (2048, 0.5, "hanning", "c64", "smem"), (2048, 0.1, "kaiser", "u8", "smem"), (64, 0.1, "ones", "c128", "smem"),
    (8192, 0.25, "hamming", "c64", "smem"), (16384, 0.1, "hanning", "c128", "fourstep"), (1 << 17, 0.5, "ones", "u8", "fourstep"),
    (1000, 0.25, "hamming", "c64", "bluestein"), (5001, 0.1, "kaiser", "c128", "bluestein"), (48000, 0.5, "hanning", "u8", "mixedradix"),
])
def test_use_psd_mode(F, r, wname, fmt, path):
    S = O.full_size(F, FS)
    win = O.window_table(wname, F)
    x = synth.tones_noise(S * 2, seed=F % 89, dtype=np.complex128, sigma=0.02)
    if fmt == "u8":
        raw = synth.to_u8_iq(x)
        xin = synth.from_u8_iq(raw)
    elif fmt == "c64":
        raw = x.astype(np.complex64)
        xin = raw.astype(np.complex128)
    else:
        raw = xin = x
    ref = [O.curscan_psd(xin[k * S:(k + 1) * S], F, r, win) for k in range(2)]
    with Plan(F, S, r, win, "PSD", _ffi.in_format(raw)) as plan:
        assert plan.precision == "f64" and plan.path == path
        assert np.array_equal(plan.frame_offsets(), O.psd_segments(F, S, r))       # segment starts: bit-exact
        got = plan.curscan(raw[:S * (2 if fmt == "u8" else 1)])
        z = plan.zerospan_batch(raw, 2, 19.1, O.adjust_xres(F, 512), "MAX", rows="db")
    assert np.max(np.abs(db(got) - db(ref[0]))) < F64_TOL
    assert int(np.argmax(got)) == int(np.argmax(ref[0]))
    refz = O.zerospan(ref, 19.1, O.adjust_xres(F, 512), "MAX")
    for k, kk in (("rows", "cur_rows"), ("hm_rows", "hm_rows"), ("max", "max"), ("min", "min"), ("avg", "avg")):
        assert np.max(np.abs(z[k] - refz[kk])) < F64_TOL, k


def test_use_psd_needs_float64():
    from kspec.engine import KspecError
    with pytest.raises(KspecError):
        Plan(2048, 16384, 0.5, np.hanning(2048), "PSD", _ffi.IN_C64, precision="f32")


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("shard", [None, (300, 1000)])
def test_pipelined_host_batch_equals_single_shot(prec, shard, monkeypatch):
    """large host batches cross PCIe in chunks while earlier chunks are transformed (kspec_zerospan_batch); every output must
    equal the one-shot path: here 300 scans in chunks of 64 (the chunk size is an environment knob for this test)"""
    F, S, r, n = 512, 4096, 0.5, 300
    win = O.window_table("hanning", F)
    x = synth.tones_noise(n * S, seed=77)
    rng = np.random.default_rng(3)
    adj = rng.normal(size=F) * 0.1
    state = (rng.normal(size=F) - 30, rng.normal(size=F) - 60, rng.normal(size=F) - 45)
    kw = dict(adj=adj, rows="db", want_hm=True)
    if shard is not None:
        kw.update(scan_index_base=shard[0], n_scans_total=shard[1])
    else:
        kw.update(state=state)
    with Plan(F, S, r, win, "AVG", _ffi.IN_C64, precision=prec) as plan:
        one = plan.zerospan_batch(x, n, 19.1, 128, "MAX", **kw)
    monkeypatch.setenv("KSPEC_PIPELINE_CHUNK_BYTES", str(64 * S * 8))      # read once, at plan creation
    with Plan(F, S, r, win, "AVG", _ffi.IN_C64, precision=prec) as plan:
        launches0 = plan.launch_count()
        piped = plan.zerospan_batch(x, n, 19.1, 128, "MAX", **kw)
        assert plan.launch_count() - launches0 >= 5 * 2            # five parts: engine + stats each
    # the 64-scan parts take the frame-parallel form of the kernel, the one-shot batch the batch form: same values up to the
    # rounding of the working precision
    eq = 1e-4 if prec == "f32" else 1e-9
    for k in ("rows", "hm_rows", "max", "min", "avg"):
        assert np.max(np.abs(piped[k] - one[k])) < eq, k
    if shard is None and prec == "f64":
        lin = [O.curscan(x[k * S:(k + 1) * S].astype(np.complex128), F, r, win, "AVG") for k in range(n)]
        ref = O.zerospan(lin, 19.1, 128, "MAX", adj=adj, state=state)
        for k, kk in (("rows", "cur_rows"), ("hm_rows", "hm_rows"), ("max", "max"), ("min", "min"), ("avg", "avg")):
            assert np.max(np.abs(piped[k] - ref[kk])) < F64_TOL, k


@pytest.mark.parametrize("fmt", ["u8", "c64"])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_pipelined_host_batch_with_odd_full_size(fmt, prec, monkeypatch):
    """an odd fullSize makes a scan a non-multiple of 16 bytes: the chunks of the pipelined host path must still start on
    16-byte boundaries of the device buffer (the fused kernels stage frames with 16-byte granular bulk copies), and the
    device entry point must refuse a misaligned sample pointer instead of faulting"""
    F, S, r, n = 512, 4099, 0.3, 203
    win = O.window_table("hamming", F)
    x = synth.tones_noise(n * S, seed=78, dtype=np.complex128)
    raw = synth.to_u8_iq(x) if fmt == "u8" else x.astype(np.complex64)
    xin = synth.from_u8_iq(raw) if fmt == "u8" else raw.astype(np.complex128)
    eb = 2 if fmt == "u8" else 8
    monkeypatch.setenv("KSPEC_PIPELINE_CHUNK_BYTES", str(65 * S * eb))     # 65 scans: not a whole number of 16-byte granules
    with Plan(F, S, r, win, "AVG", _ffi.in_format(raw), precision=prec) as plan:
        got = plan.zerospan_batch(raw, n, 19.1, 128, "MAX", rows="linear")
        import ctypes as C
        from kspec._ffi import KspecError
        d = plan.dev_alloc(raw.nbytes + 64)
        plan.dev_upload(d, raw, offset=8)
        with pytest.raises(KspecError):
            plan.zerospan_batch_dev(C.c_void_p(d.value + 8), n, 19.1, 128, "MAX")
        plan.dev_free(d)
    ref = np.array([O.curscan(xin[k * S:(k + 1) * S], F, r, win, "AVG") for k in range(n)])
    for k in range(n):
        assert_db_close(got["rows"][k], ref[k], DB_TOL if prec == "f32" else F64_TOL, "scan %d" % k, f32=prec == "f32")


# ---------------------------------------------------------------------------------------------------------------
# the float32 contract of include/kspec.h on the synthetic inputs of every BASELINE configuration
# ---------------------------------------------------------------------------------------------------------------
def _contract_rows(F, r, wname, cumu, x, S, n):
    win = O.window_table(wname, F)
    with Plan(F, S, r, win, cumu, _ffi.in_format(x), precision="f32") as plan:
        offs = plan.frame_offsets()
        got = plan.zerospan_batch(x, n, 19.1, 512, "MAX", rows="linear", want_hm=False)["rows"]
    ref = np.array([O.curscan(x[k * S:(k + 1) * S].astype(np.complex128), F, r, win, cumu) for k in range(n)])
    return got, ref, offs


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg5a"])
def test_f32_contract(cfg):
    """KSPEC_PREC_F32 (include/kspec.h): every bin within 3e-7 of the scan's peak, bins above 2e-3 of the peak within 1e-3 dB,
    argmax and frame offsets exact -- on the BASELINE synthetic input of every configuration where float32 can be selected;
    on cfg 1 (the benchmark workload) every bin of every scan within 1e-3 dB."""
    if cfg == "cfg1":
        F, r, wname, cumu, n = 2048, 0.5, "hanning", "AVG", 146
        S = O.full_size(F, FS)
        x = synth.tones_noise(n * S, seed=1)
    elif cfg == "cfg2":
        F, r, wname, cumu, n = 64, 0.1, "ones", "AVG", 613
        S = O.full_size(F, FS)
        x = np.concatenate([synth.step_tones(s, S) for s in range(n)])
    elif cfg == "cfg3":
        F, r, wname, cumu, n = 8192, 0.25, "kaiser", "AVG", 40
        S = O.full_size(F, FS)
        x = synth.tones_noise(n * S, seed=3, gate=(40000, 0.5))
    else:
        F, r, wname, cumu, n = 4096, 0.1, "ones", "AVG", 18
        S = O.full_size(F, FS)
        x = np.concatenate([synth.step_tones(s, S) for s in range(n)])
    got, ref, offs = _contract_rows(F, r, wname, cumu, x, S, n)
    assert np.array_equal(offs, O.frame_offsets(F, S, r))
    peak = ref.max(axis=1, keepdims=True)
    assert np.max(np.abs(got - ref) / peak) < F32_ABS
    m = ref > F32_DYN * peak
    assert np.max(np.abs(db(got[m]) - db(ref[m]))) < DB_TOL
    assert np.array_equal(np.argmax(got, axis=1), np.argmax(ref, axis=1))
    if cfg == "cfg1":
        assert np.max(np.abs(db(got) - db(ref))) < DB_TOL


@pytest.mark.parametrize("F", [1 << 21, 2400000])
def test_f32_is_refused_where_it_is_not_offered(F):
    """cfg 4 (2^21) and cfg 5b (2.4e6) run on the multi-pass float64 engines: an explicit float32 request fails loudly"""
    from kspec._ffi import KspecError
    with pytest.raises(KspecError):
        Plan(F, 2 * F, 0.1, np.ones(F), "MAX", _ffi.IN_C64, precision="f32")
    with Plan(F, 2 * F, 0.1, np.ones(F), "MAX", _ffi.IN_C64, precision="auto") as plan:
        assert plan.precision == "f64"


# ---------------------------------------------------------------------------------------------------------------
# memory safety of the staged bulk copies (compute-sanitizer is closed on this pool): poisoned guard regions
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,prec,F,r,n", [
    ("c64", "f32", 2048, 0.1, 37), ("u8", "f32", 2048, 0.1, 37), ("c64", "f32", 2048, 0.5, 1813), ("u8", "f32", 2048, 0.5, 1813),
    ("c64", "f64", 2048, 0.1, 9), ("u8", "f64", 1024, 0.1, 9), ("c64", "f32", 256, 0.1, 300), ("c64", "f32", 8192, 0.25, 5),
])
def test_sample_buffer_between_poisoned_guards(fmt, prec, F, r, n):
    """The sample batch sits flush between two guard regions filled with 0xFF bytes (NaN as float32, 255 as uint8).  The
    16-byte granular bulk copies of the fused kernels (odd frame offsets at r = 0.1, the last frame of the last scan) must not
    bring a guard byte into any result, and nothing may write outside the plan's own buffers: outputs identical to a run on a
    zero-padded copy, guards unchanged."""
    G = 4096
    S = O.full_size(F, FS)
    x = synth.tones_noise(n * S, seed=41, dtype=np.complex128)
    raw = synth.to_u8_iq(x) if fmt == "u8" else x.astype(np.complex64)
    nbytes = raw.nbytes
    assert nbytes % 16 == 0
    win = O.window_table("hanning", F)
    with Plan(F, S, r, win, "AVG", _ffi.in_format(raw), precision=prec) as plan:
        import ctypes as C
        buf = plan.dev_alloc(nbytes + 2 * G)
        poison = np.full(nbytes + 2 * G, 0xFF, dtype=np.uint8)
        plan.dev_upload(buf, poison)
        plan.dev_upload(buf, raw, offset=G)
        plan.zerospan_batch_dev(C.c_void_p(buf.value + G), n, 19.1, 512, "MAX", rows="db")
        a = plan.zerospan_fetch()
        lo = plan.dev_download(buf, G)
        hi = plan.dev_download(buf, G, offset=G + nbytes)
        mid = plan.dev_download(buf, nbytes, offset=G)
        plan.dev_upload(buf, np.zeros(nbytes + 2 * G, dtype=np.uint8))
        plan.dev_upload(buf, raw, offset=G)
        plan.zerospan_batch_dev(C.c_void_p(buf.value + G), n, 19.1, 512, "MAX", rows="db")
        b = plan.zerospan_fetch()
        plan.dev_free(buf)
    assert (lo == 0xFF).all() and (hi == 0xFF).all() and np.array_equal(mid, raw.view(np.uint8).reshape(-1))
    for k in ("rows", "hm_rows", "max", "min", "avg"):
        assert np.isfinite(a[k]).all(), k
        assert np.array_equal(a[k], b[k]), k
