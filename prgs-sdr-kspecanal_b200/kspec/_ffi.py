"""ctypes binding of libkspec.so (include/kspec.h).  No torch, no cupy, no CPU fallback.

The library is built in-tree (``make -C prgs-sdr-kspecanal_b200`` or ``__graft_entry__.build()``) next to
this file.  Importing this module without it raises ``KspecError`` immediately: the product path never
degrades to numpy.
"""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libkspec.so")

OK = 0
CUMU = {"RAW": 0, "AVG": 1, "MAX": 2, "MIN": 3, "PSD": 4}
IN_U8_IQ, IN_C64, IN_C128 = 0, 1, 2
PREC = {"auto": 0, "f32": 1, "f64": 2}
PREC_NAME = {1: "f32", 2: "f64"}
COMPRESS = {"RAW": 0, "MAX": 1, "AVG": 2, "MIN": 3}
PATH_NAME = {0: "smem", 1: "fourstep", 2: "bluestein", 3: "mixedradix"}
ROWS_NONE, ROWS_LINEAR, ROWS_DB = 0, 1, 2


class KspecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libkspec error %d: %s" % (code, msg))
        self.code = code


class PlanInfo(C.Structure):
    _fields_ = [
        ("fft_size", C.c_int32), ("full_size", C.c_int64), ("n_frames", C.c_int32), ("precision", C.c_int32),
        ("path", C.c_int32), ("in_fmt", C.c_int32), ("device", C.c_int32), ("sm_count", C.c_int32),
        ("cta_threads", C.c_int32), ("ctas_per_sm", C.c_int32), ("smem_bytes", C.c_int32), ("scans_per_cta", C.c_int32),
        ("tma_stages", C.c_int32),
        ("conv_size", C.c_int64), ("win_adj", C.c_double),
    ]


# every symbol include/kspec.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.c_int64
SIGNATURES = {
    "kspec_version": (C.c_int, []),
    "kspec_last_error": (C.c_char_p, []),
    "kspec_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kspec_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, _I64, C.c_double, C.c_int, _D, C.c_int, C.c_double, C.c_double,
                                    C.c_int, C.c_int]),
    "kspec_plan_destroy": (C.c_int, [_P]),
    "kspec_plan_frames": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(C.c_int)]),
    "kspec_plan_info": (C.c_int, [_P, C.POINTER(PlanInfo)]),
    "kspec_curscan": (C.c_int, [_P, _P, _D]),
    "kspec_zerospan_batch": (C.c_int, [_P, _P, _I64, C.c_double, _D, C.c_int, C.c_int, C.c_int, _D, _D, _D, _D, _D, C.c_int,
                                       _I64, _I64]),
    "kspec_zerospan_rows_batch": (C.c_int, [_P, _D, _I64, C.c_double, _D, C.c_int, C.c_int, _D, _D, _D, _D, _D, C.c_int]),
    "kspec_scan_batch": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_I64), C.POINTER(_I64), _I64, C.c_double,
                                   C.c_double, C.c_int, C.c_int, _D, _D, _D, _D]),
    "kspec_scan_state_init": (C.c_int, [_P, _I64, _D, _D, _D, _D]),
    "kspec_scan_pass": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_I64), C.POINTER(_I64), C.c_double, C.c_double, C.c_int, C.c_int]),
    "kspec_scan_pass_dev": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_I64), C.POINTER(_I64), C.c_double, C.c_double, C.c_int, C.c_int]),
    "kspec_scan_state_fetch": (C.c_int, [_P, _D, _D, _D, _D]),
    "kspec_scan_shard": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_I64), _I64, C.c_double, C.c_double, _D]),
    "kspec_scan_stats_update": (C.c_int, [_P, _D, _I64, _I64, C.c_int, _D, _D, _D]),
    "kspec_plotcompress": (C.c_int, [_P, _D, _I64, C.c_int, C.c_int, _D]),
    "kspec_plot_highs": (C.c_int, [_P, _D, _D, _I64, C.c_int, C.c_double, C.POINTER(_I64), C.POINTER(C.c_int)]),
    "kspec_conv_smooth": (C.c_int, [_P, _D, _I64, _D, C.c_int, C.c_int, _D]),
    "kspec_dev_alloc": (C.c_int, [_P, _I64, C.POINTER(_P)]),
    "kspec_dev_free": (C.c_int, [_P, _P]),
    "kspec_dev_upload": (C.c_int, [_P, _P, _P, _I64]),
    "kspec_dev_download": (C.c_int, [_P, _P, _P, _I64]),
    "kspec_dev_fill_l2": (C.c_int, [_P]),
    "kspec_host_alloc": (C.c_int, [_I64, C.POINTER(_P)]),
    "kspec_host_free": (C.c_int, [_P]),
    "kspec_zerospan_batch_dev": (C.c_int, [_P, _P, _I64, C.c_double, _D, C.c_int, C.c_int, C.c_int, C.c_int, _D, _D, _D, C.c_int,
                                           _I64, _I64]),
    "kspec_zerospan_fetch": (C.c_int, [_P, _D, _D, _D, _D, _D]),
    "kspec_sync": (C.c_int, [_P]),
    "kspec_timer_start": (C.c_int, [_P]),
    "kspec_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "kspec_kernel_times": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "kspec_plan_reserve_sms": (C.c_int, [_P, C.c_int]),
    "kspec_launch_count": (C.c_int, [_P, C.POINTER(_I64)]),
    "kspec_comm_unique_id": (C.c_int, [C.c_char_p]),
    "kspec_comm_init": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "kspec_comm_allreduce_stats": (C.c_int, [_P, _D, _D, _D, _I64]),
    "kspec_comm_allreduce_sum": (C.c_int, [_P, _D, _I64]),
    "kspec_comm_allreduce_plan": (C.c_int, [_P, _P]),
    "kspec_comm_join": (C.c_int, [_P, _P]),
    "kspec_comm_fetch_reduced": (C.c_int, [_P, _D, _D, _D, _I64]),
    "kspec_comm_peer_setup": (C.c_int, [_P, _P]),
    "kspec_comm_peer_status": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "kspec_comm_finalize": (C.c_int, [_P]),
}

_lib = None


def lib():
    """The loaded library; raises KspecError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise KspecError(-2, "%s not found: build it with `make -C prgs-sdr-kspecanal_b200` "
                                 "(there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != OK:
        raise KspecError(rc, lib().kspec_last_error().decode("utf-8", "replace"))


def dptr(a):
    """float64 ndarray (C-contiguous, writable where needed) -> double*; None -> NULL."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_D)


def vptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def in_format(samples):
    """ingest format code of a host sample array (uint8 interleaved IQ, complex64 or complex128)."""
    dt = np.asarray(samples).dtype
    if dt == np.uint8:
        return IN_U8_IQ
    if dt == np.complex64:
        return IN_C64
    if dt == np.complex128:
        return IN_C128
    raise TypeError("IQ samples must be uint8 (interleaved I,Q), complex64 or complex128, not %s" % dt)


IN_ELEM_BYTES = {IN_U8_IQ: 2, IN_C64: 8, IN_C128: 16}


class PinnedBuffer:
    """cudaHostAlloc-ed (pinned) host memory exposed as a numpy array (``.array``, uint8; ``.view(dtype)``).
    H2D/D2H copies from pinned memory run at full PCIe rate and asynchronously."""

    def __init__(self, n_bytes):
        self._p = _P()
        self.nbytes = int(n_bytes)
        check(lib().kspec_host_alloc(self.nbytes, C.byref(self._p)))
        self._raw = (C.c_uint8 * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(self._raw, dtype=np.uint8)

    def view(self, dtype):
        return self.array.view(dtype)

    def free(self):
        if self._p is not None and self._p.value:
            self.array = None
            self._raw = None
            lib().kspec_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
