"""CPU-only checks of the drop-in boundary: libkspec.so loads, exports every symbol include/kspec.h declares,
refuses to run without a GPU (no CPU fallback), and the host-side mirror reproduces the reference's
configuration arithmetic.  No compute calls."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from kspec import _ffi, hotpath
from kspec.engine import Plan, device_count


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kspec.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kspec_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(_ffi.SIGNATURES) == syms


def header_enums():
    txt = open(os.path.join(ROOT, "include", "kspec.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return {k: int(v) for k, v in re.findall(r"\b(KSPEC_[A-Z0-9_]+)\s*=\s*(-?\d+)", txt)}


def test_binding_constants_match_the_header_enums():
    e = header_enums()
    for name, v in _ffi.CUMU.items():
        assert e["KSPEC_CUMU_" + name] == v
    assert (e["KSPEC_IN_U8_IQ"], e["KSPEC_IN_C64"], e["KSPEC_IN_C128"]) == (_ffi.IN_U8_IQ, _ffi.IN_C64, _ffi.IN_C128)
    for name, v in _ffi.PREC.items():
        assert e["KSPEC_PREC_" + name.upper()] == v
    for name, v in _ffi.COMPRESS.items():
        assert e["KSPEC_COMPRESS_" + name] == v
    for v, name in _ffi.PATH_NAME.items():
        assert e["KSPEC_PATH_" + name.upper()] == v
    assert (e["KSPEC_ROWS_NONE"], e["KSPEC_ROWS_LINEAR"], e["KSPEC_ROWS_DB"]) == (_ffi.ROWS_NONE, _ffi.ROWS_LINEAR, _ffi.ROWS_DB)


def test_library_exports_every_symbol():
    assert os.path.isfile(_ffi.LIB_PATH), "build with __graft_entry__.build() / make -C prgs-sdr-kspecanal_b200"
    h = ctypes.CDLL(_ffi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(h, s), s
    assert _ffi.lib().kspec_version() == 100


def test_plan_info_struct_matches_header():
    # 4 (+4 pad) + 8 + 11*4 (+4 pad) + 8 + 8 with natural alignment
    assert ctypes.sizeof(_ffi.PlanInfo) == 80


@pytest.mark.skipif(device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(_ffi.KspecError) as e:
        Plan(2048, 16384, 0.5, np.hanning(2048))
    assert "no CPU fallback" in str(e.value)


def test_bad_arguments_are_errors_not_exits():
    lib = _ffi.lib()
    h = ctypes.c_void_p()
    w = np.ones(8)
    rc = lib.kspec_plan_create(ctypes.byref(h), 8, 4, 0.5, 1, _ffi.dptr(w), 1, 0.0, 1.0, 0, 0)
    assert rc == -1 and b"fullSize" in lib.kspec_last_error()
    rc = lib.kspec_plan_create(ctypes.byref(h), 8, 64, 0.5, 9, _ffi.dptr(w), 1, 0.0, 1.0, 0, 0)
    assert rc == -1 and b"cumuMode" in lib.kspec_last_error()
    rc = lib.kspec_zerospan_fetch(None, None, None, None, None, None)
    assert rc == -1


def test_derive_config_matches_reference_handle_args():
    rows = json.load(open(os.path.join(GOLDEN, "g7_handle_args.json")))
    for row in rows:
        a = row["argv"]
        if row["prgMode"] == "SCAN":
            d = dict(fftSize=row["fftSize"])
            if a[0] == "quickFullScan":
                d.update(startFreq=30e6, endFreq=1.5e9)
            elif a[0] == "fmScan":
                d.update(startFreq=88e6, endFreq=108e6)
            else:
                d.update(startFreq=float(a[a.index("startFreq") + 1]), endFreq=float(a[a.index("endFreq") + 1]))
            d["samplingRate"] = 2.4e6
            hotpath._fixupfreqs_scanrange(d)
            assert (d["startFreq"], d["endFreq"], d["centerFreq"]) == (row["startFreq"], row["endFreq"], row["centerFreq"])
        else:
            d = dict(fftSize=row["fftSize"])
        if "xRes" in a:
            d["xRes"] = int(a[a.index("xRes") + 1])
        if row["fftSize"] > 2 ** 20:
            continue      # window tables of 2M+ points: covered by the oracle test, skip the allocation here
        hotpath.derive_config(d)
        assert d["fullSize"] == row["fullSize"] and d["xRes"] == row["xRes"], a


def test_cli_handle_args_matches_reference_table():
    """kspec.cli.handle_args == the reference's handle_args (K:778-949) on the golden argv table"""
    from kspec import cli
    rows = json.load(open(os.path.join(GOLDEN, "g7_handle_args.json")))
    for row in rows:
        if row["fftSize"] > 2 ** 20:
            continue
        d = {"cmd.stop": False}
        cli.handle_args(d, row["argv"])
        for k in ("prgMode", "fftSize", "fullSize", "xRes", "startFreq", "endFreq", "centerFreq", "pltCompress"):
            assert d[k] == row[k], (row["argv"], k, d[k], row[k])
    with pytest.raises(SystemExit):
        cli.handle_args({"cmd.stop": False}, ["zeroSpan", "noSuchKey", "1"])


def test_scan_geometry_matches_oracle():
    from oracle import kspec_oracle as O
    for (s, e, F, R) in ((30e6, 30e6 + 11 * 2.4e6, 64, 1.0), (88e6, 109.6e6, 4096, 0.5), (100e6, 107.2e6, 1200, 0.25)):
        d = dict(startFreq=s, endFreq=e, samplingRate=2.4e6, fftSize=F, scanRangeNonOverlap=R)
        ng, tot, steps = hotpath.scan_geometry(d)
        g2 = O.scan_geometry(s, e, 2.4e6, F, R)
        assert (ng, tot) == g2[:2]
        assert [(st[1], st[2]) for st in steps] == [(x["i_start"], x["i_done"]) for x in g2[2]]
        assert [st[0] for st in steps] == [x["cur_freq"] for x in g2[2]]


def test_scan_geometry_rejects_non_integer_steps():
    d = dict(startFreq=0.0, endFreq=4.8e6, samplingRate=2.4e6, fftSize=64, scanRangeNonOverlap=0.3)
    with pytest.raises(SystemExit):
        hotpath.scan_geometry(d)
    assert d["cmd.stop"] is True


def test_sdr_read_pattern():
    """K:311-347: 2^18 chunks, power-of-two over-read of the tail."""
    calls = []

    class Dev:
        def read_samples(self, n):
            calls.append(int(n))
            return np.arange(int(n)) + 0j
    out = hotpath.sdr_read(Dev(), 4800000)
    assert calls == [2 ** 18] * 18 + [131072] and len(out) == 4800000
    calls.clear()
    hotpath.sdr_read(Dev(), 16384)
    assert calls == [16384]
