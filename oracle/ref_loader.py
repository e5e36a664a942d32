"""ORACLE support — load the UNMODIFIED reference functions from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``kspecanal.py`` cannot be imported: it imports ``matplotlib.pyplot`` and
``rtlsdr`` (neither installed) at the top (K:11,13) and runs the program at module level from
``gD = {}`` on (K:1139).  Recipe: register stub modules, read the file, compile
only the lines before ``gD = {}`` and exec them into a fresh namespace.  Nothing is copied into the
repository; the GPU box has no /root/reference and never calls this.
"""
import os
import sys
import types
from unittest import mock

REF_FILE = "/root/reference/python/kspecanal.py"


def available():
    return os.path.isfile(REF_FILE)


def load(sdr_factory=None):
    """Return the reference's module namespace (dict).  ``sdr_factory`` is what ``rtlsdr.RtlSdr()``
    returns when the reference re-opens the device after a tune failure (K:304-305)."""
    plt = mock.MagicMock(name="matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    rtl = types.ModuleType("rtlsdr")
    rtl.RtlSdr = sdr_factory if sdr_factory is not None else mock.MagicMock(name="RtlSdr")
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "rtlsdr")}
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["rtlsdr"] = rtl
    try:
        with open(REF_FILE) as f:
            lines = f.readlines()
        cut = next(i for i, ln in enumerate(lines) if ln.startswith("gD = {}"))
        ns = {"__name__": "kspecanal_ref"}
        exec(compile("".join(lines[:cut]), REF_FILE, "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ns["plt"] = plt
    ns["rtlsdr"] = rtl
    return ns


def base_dict(ns, argv):
    """Run the reference's own handle_args (K:778-949) on a CLI vector and return the dict ``d``."""
    d = {"cmd.stop": False}
    old = sys.argv
    sys.argv = ["kspecanal.py"] + [str(a) for a in argv]
    try:
        with mock.patch("builtins.input", lambda *a, **k: ""):
            ns["handle_args"](d)
    finally:
        sys.argv = old
    d["AxLevels"] = mock.MagicMock()
    d["AxHeatMap"] = mock.MagicMock()
    d["AxFreqs"] = mock.MagicMock()
    return d
