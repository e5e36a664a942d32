"""Thin object layer over the C ABI: one ``Plan`` = one (fftSize, fullSize, window, overlap, cumulate mode,
ingest format, device).  All arithmetic happens inside libkspec.so on the GPU; this file only marshals
numpy arrays.  Reference lines replaced are cited per method (``K:`` = python/kspecanal.py).
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import KspecError, check, dptr, vptr  # noqa: F401


def device_count():
    n = C.c_int(0)
    rc = _ffi.lib().kspec_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def heatmap_width(fft_size, x_res, hm_mode):
    """K:449-457."""
    return x_res if (hm_mode.upper() != "RAW" and fft_size > x_res) else fft_size


class Plan:
    """Derived state of handle_args' tail (K:926-936) and sdr_curscan's setup (K:368-373) on one GPU."""

    def __init__(self, fft_size, full_size, non_overlap, window, cumu_mode="AVG", in_fmt=_ffi.IN_C64,
                 precision="auto", device=0, u8_offset=127.5, u8_scale=1.0 / 127.5):
        self._h = C.c_void_p()
        window = np.ascontiguousarray(window, dtype=np.float64)
        if window.shape != (fft_size,):
            raise ValueError("window must have fftSize=%d entries" % fft_size)
        mode = _ffi.CUMU.get(str(cumu_mode).upper())
        if mode is None:
            # the reference quits on an unknown cumuMode (K:144-146)
            raise KspecError(-1, "Unknown cumuMode [%s]" % cumu_mode)
        check(_ffi.lib().kspec_plan_create(C.byref(self._h), int(fft_size), int(full_size), float(non_overlap), mode,
                                           dptr(window), int(in_fmt), float(u8_offset), float(u8_scale),
                                           _ffi.PREC[precision], int(device)))
        self.fft_size, self.full_size, self.in_fmt = int(fft_size), int(full_size), int(in_fmt)
        info = _ffi.PlanInfo()
        check(_ffi.lib().kspec_plan_info(self._h, C.byref(info)))
        self.info = info
        self.precision = _ffi.PREC_NAME[info.precision]
        self.path = _ffi.PATH_NAME[info.path]
        self.n_frames = info.n_frames

    # -- lifetime ----------------------------------------------------------------------------------
    def close(self):
        if self._h is not None and self._h.value:
            _ffi.lib().kspec_plan_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------------
    def _samples(self, samples, n_units):
        a = np.ascontiguousarray(samples)
        if _ffi.in_format(a) != self.in_fmt:
            raise TypeError("plan was created for ingest format %d, got dtype %s" % (self.in_fmt, a.dtype))
        per = self.full_size * (2 if self.in_fmt == _ffi.IN_U8_IQ else 1)
        if a.size != per * n_units:
            raise ValueError("expected %d x %d elements, got %d" % (n_units, per, a.size))
        return a

    def frame_offsets(self):
        """Start index of every frame of a scan (K:386-390)."""
        n = C.c_int(0)
        check(_ffi.lib().kspec_plan_frames(self._h, None, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        check(_ffi.lib().kspec_plan_frames(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)))
        return out

    # -- sdr_curscan (K:351-397) ----------------------------------------------------------------------
    def curscan(self, samples):
        a = self._samples(samples, 1)
        out = np.empty(self.fft_size, dtype=np.float64)
        check(_ffi.lib().kspec_curscan(self._h, vptr(a), dptr(out)))
        return out

    # -- zero_span loop body (K:464-484) over n scans ---------------------------------------------------
    def zerospan_batch(self, samples, n_scans, gain, x_res, hm_mode="MAX", adj=None, rows=None, want_hm=True,
                       state=None, scan_index_base=0, n_scans_total=None, out=None):
        """rows: None | "linear" | "db".  state: (max, min, avg) float64 arrays carried in, or None.
        out: optional dict(rows=, hm_rows=) of caller-owned float64 arrays to receive the rows (views of a
        ``_ffi.PinnedBuffer`` make the device-to-host copies run at the PCIe rate instead of through pageable staging).
        Returns dict(rows, hm_rows, max, min, avg)."""
        a = self._samples(samples, n_scans)
        F = self.fft_size
        kind = {None: _ffi.ROWS_NONE, "linear": _ffi.ROWS_LINEAR, "db": _ffi.ROWS_DB}[rows]
        W = heatmap_width(F, x_res, hm_mode)

        def _out(key, shape, wanted):
            if not wanted:
                return None
            buf = None if out is None else out.get(key)
            if buf is None:
                return np.empty(shape, dtype=np.float64)
            if buf.dtype != np.float64 or not buf.flags["C_CONTIGUOUS"] or buf.shape != shape:
                raise ValueError("out[%s] must be a contiguous float64 array of shape %s" % (key, shape))
            return buf

        rows_out = _out("rows", (n_scans, F), bool(kind))
        hm_out = _out("hm_rows", (n_scans, W), want_hm)
        if state is not None:
            mx, mn, av = (np.array(s, dtype=np.float64, copy=True) for s in state)
        else:
            mx, mn, av = (np.empty(F, dtype=np.float64) for _ in range(3))
        adj_a = None if adj is None else np.ascontiguousarray(adj, dtype=np.float64)
        total = n_scans if n_scans_total is None else n_scans_total
        check(_ffi.lib().kspec_zerospan_batch(self._h, vptr(a), n_scans, float(gain), dptr(adj_a), _ffi.COMPRESS[hm_mode.upper()],
                                              int(x_res), kind, dptr(rows_out), dptr(hm_out), dptr(mx), dptr(mn), dptr(av),
                                              1 if state is not None else 0, int(scan_index_base), int(total)))
        return dict(rows=rows_out, hm_rows=hm_out, max=mx, min=mn, avg=av)

    # -- zero_span loop body on existing spectra: zeroSpanPlay (K:547-564 -> K:469-484) ----------------------
    def zerospan_rows_batch(self, lin_rows, gain, x_res, hm_mode="MAX", adj=None, want_rows=True, want_hm=True, state=None):
        lin = np.ascontiguousarray(lin_rows, dtype=np.float64)
        n, F = lin.shape
        if F != self.fft_size:
            raise ValueError("rows must have fftSize=%d bins" % self.fft_size)
        rows_out = np.empty((n, F), dtype=np.float64) if want_rows else None
        hm_out = np.empty((n, heatmap_width(F, x_res, hm_mode)), dtype=np.float64) if want_hm else None
        if state is not None:
            mx, mn, av = (np.array(s, dtype=np.float64, copy=True) for s in state)
        else:
            mx, mn, av = (np.empty(F, dtype=np.float64) for _ in range(3))
        adj_a = None if adj is None else np.ascontiguousarray(adj, dtype=np.float64)
        check(_ffi.lib().kspec_zerospan_rows_batch(self._h, dptr(lin), n, float(gain), dptr(adj_a), _ffi.COMPRESS[hm_mode.upper()],
                                                   int(x_res), dptr(rows_out), dptr(hm_out), dptr(mx), dptr(mn), dptr(av),
                                                   1 if state is not None else 0))
        return dict(rows=rows_out, hm_rows=hm_out, max=mx, min=mn, avg=av)

    # -- _scan_range step loop (K:619-668) ---------------------------------------------------------------
    def scan_batch(self, samples, n_steps, i_start, i_done, total_entries, min_amp, gain, state, pass_index,
                   step_ok=None, base_is_raw=False):
        """state: dict(cur, max, min, avg) of float64[totalEntries], updated in place."""
        a = self._samples(samples, n_steps)
        i_start = np.ascontiguousarray(i_start, dtype=np.int64)
        i_done = np.ascontiguousarray(i_done, dtype=np.int64)
        ok = None if step_ok is None else np.ascontiguousarray(step_ok, dtype=np.uint8)
        for k in ("cur", "max", "min", "avg"):
            if state[k].dtype != np.float64 or not state[k].flags["C_CONTIGUOUS"] or state[k].shape != (total_entries,):
                raise ValueError("state[%s] must be a contiguous float64[%d]" % (k, total_entries))
        check(_ffi.lib().kspec_scan_batch(self._h, vptr(a), int(n_steps),
                                          None if ok is None else ok.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          i_start.ctypes.data_as(C.POINTER(C.c_int64)), i_done.ctypes.data_as(C.POINTER(C.c_int64)),
                                          int(total_entries), float(min_amp), float(gain), 1 if base_is_raw else 0, int(pass_index),
                                          dptr(state["cur"]), dptr(state["max"]), dptr(state["min"]), dptr(state["avg"])))
        return state

    # -- the same pass with the state resident on the device (kspec_scan_state_* / kspec_scan_pass) ------------------------
    def scan_state_init(self, state):
        """state: dict(cur, max, min, avg) of float64[totalEntries] (K:602-608), uploaded once"""
        a = {k: np.ascontiguousarray(state[k], dtype=np.float64) for k in ("cur", "max", "min", "avg")}
        n = len(a["cur"])
        check(_ffi.lib().kspec_scan_state_init(self._h, n, dptr(a["cur"]), dptr(a["max"]), dptr(a["min"]), dptr(a["avg"])))
        self._scan_total = n

    def scan_pass(self, samples, n_steps, i_start, i_done, min_amp, gain, pass_index, step_ok=None, base_is_raw=False, on_device=False):
        """one pass of _scan_range on the device-resident state; samples: host array (chunked, overlapped H2D) or, with
        on_device=True, a device pointer"""
        i_start = np.ascontiguousarray(i_start, dtype=np.int64)
        i_done = np.ascontiguousarray(i_done, dtype=np.int64)
        ok = None if step_ok is None else np.ascontiguousarray(step_ok, dtype=np.uint8)
        okp = None if ok is None else ok.ctypes.data_as(C.POINTER(C.c_uint8))
        args = (int(n_steps), okp, i_start.ctypes.data_as(C.POINTER(C.c_int64)), i_done.ctypes.data_as(C.POINTER(C.c_int64)),
                float(min_amp), float(gain), 1 if base_is_raw else 0, int(pass_index))
        if on_device:
            check(_ffi.lib().kspec_scan_pass_dev(self._h, samples, *args))
        else:
            a = self._samples(samples, n_steps)
            check(_ffi.lib().kspec_scan_pass(self._h, vptr(a), *args))

    def scan_state_fetch(self, which=("cur", "max", "min", "avg")):
        n = self._scan_total
        out = {k: (np.empty(n, dtype=np.float64) if k in which else None) for k in ("cur", "max", "min", "avg")}
        check(_ffi.lib().kspec_scan_state_fetch(self._h, dptr(out["cur"]), dptr(out["max"]), dptr(out["min"]), dptr(out["avg"])))
        return {k: v for k, v in out.items() if v is not None}

    # -- the same pass sharded by frequency step ---------------------------------------------------------
    def scan_shard(self, samples, n_local, step_base, i_start_all, total_entries, min_amp, gain, step_ok=None):
        """this shard's share of the stitched Fft.Cur (float64[totalEntries]); SUM over shards = Fft.Cur"""
        a = self._samples(samples, n_local)
        i_start = np.ascontiguousarray(i_start_all, dtype=np.int64)
        ok = None if step_ok is None else np.ascontiguousarray(step_ok, dtype=np.uint8)
        out = np.empty(total_entries, dtype=np.float64)
        check(_ffi.lib().kspec_scan_shard(self._h, vptr(a), int(n_local), int(step_base), len(i_start),
                                          None if ok is None else ok.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          i_start.ctypes.data_as(C.POINTER(C.c_int64)), int(total_entries), float(min_amp), float(gain),
                                          dptr(out)))
        return out

    def scan_stats_update(self, cur, last_done, pass_index, state):
        """K:657-668 on a finished Fft.Cur: state['max'/'min'/'avg'] updated in place on bins < last_done, state['cur'] = cur"""
        cur = np.ascontiguousarray(cur, dtype=np.float64)
        check(_ffi.lib().kspec_scan_stats_update(self._h, dptr(cur), len(cur), int(last_done), int(pass_index),
                                                 dptr(state["max"]), dptr(state["min"]), dptr(state["avg"])))
        state["cur"][:] = cur
        return state

    # -- _data_plotcompress (K:168-202) ------------------------------------------------------------------
    def plotcompress(self, y, x_res, mode):
        y = np.ascontiguousarray(y, dtype=np.float64)
        mode = mode.upper()
        cols = len(y) // x_res
        n_out = len(y) if (mode == "RAW" or cols == 0) else x_res
        out = np.empty(n_out, dtype=np.float64)
        check(_ffi.lib().kspec_plotcompress(self._h, dptr(y), len(y), int(x_res), _ffi.COMPRESS[mode], dptr(out)))
        return out

    # -- plot_highs peak picking (K:243-272) and the Conv display mode (K:113-120) ------------------------------
    def plot_highs(self, freqs, levels, num_markers=5, delta4marking=0.025):
        x = np.ascontiguousarray(freqs, dtype=np.float64)
        y = np.ascontiguousarray(levels, dtype=np.float64)
        idx = np.zeros(64, dtype=np.int64)
        n = C.c_int(0)
        check(_ffi.lib().kspec_plot_highs(self._h, dptr(x), dptr(y), len(x), int(num_markers), float(delta4marking),
                                          idx.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)))
        return idx[:n.value].copy()

    def conv_smooth(self, vals, taps, edge=12):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        t = np.ascontiguousarray(taps, dtype=np.float64)
        out = np.empty_like(v)
        check(_ffi.lib().kspec_conv_smooth(self._h, dptr(v), len(v), dptr(t), len(t), int(edge), dptr(out)))
        return out

    # -- device-resident pipeline (bench / zero-copy callers) -----------------------------------------------
    def dev_alloc(self, n_bytes):
        p = C.c_void_p()
        check(_ffi.lib().kspec_dev_alloc(self._h, int(n_bytes), C.byref(p)))
        return p

    def dev_free(self, p):
        check(_ffi.lib().kspec_dev_free(self._h, p))

    def dev_upload(self, p, host, offset=0):
        host = np.ascontiguousarray(host)
        check(_ffi.lib().kspec_dev_upload(self._h, C.c_void_p(p.value + offset), vptr(host), host.nbytes))

    def dev_download(self, p, n_bytes, offset=0):
        out = np.empty(int(n_bytes), dtype=np.uint8)
        check(_ffi.lib().kspec_dev_download(self._h, vptr(out), C.c_void_p(p.value + offset), int(n_bytes)))
        return out

    def zerospan_batch_dev(self, d_samples, n_scans, gain, x_res, hm_mode="MAX", adj=None, rows=None, want_hm=True,
                           state=None, scan_index_base=0, n_scans_total=None):
        kind = {None: _ffi.ROWS_NONE, "linear": _ffi.ROWS_LINEAR, "db": _ffi.ROWS_DB}[rows]
        adj_a = None if adj is None else np.ascontiguousarray(adj, dtype=np.float64)
        mx = mn = av = None
        if state is not None:
            mx, mn, av = (np.ascontiguousarray(s, dtype=np.float64) for s in state)
        total = n_scans if n_scans_total is None else n_scans_total
        check(_ffi.lib().kspec_zerospan_batch_dev(self._h, d_samples, int(n_scans), float(gain), dptr(adj_a),
                                                  _ffi.COMPRESS[hm_mode.upper()], int(x_res), kind, 1 if want_hm else 0,
                                                  dptr(mx), dptr(mn), dptr(av), 1 if state is not None else 0,
                                                  int(scan_index_base), int(total)))
        self._last = (n_scans, kind, heatmap_width(self.fft_size, x_res, hm_mode) if want_hm else 0)

    def zerospan_fetch(self, rows=True, hm=True):
        n, kind, W = self._last
        F = self.fft_size
        rows_out = np.empty((n, F), dtype=np.float64) if (rows and kind) else None
        hm_out = np.empty((n, W), dtype=np.float64) if (hm and W) else None
        mx, mn, av = (np.empty(F, dtype=np.float64) for _ in range(3))
        check(_ffi.lib().kspec_zerospan_fetch(self._h, dptr(rows_out), dptr(hm_out), dptr(mx), dptr(mn), dptr(av)))
        return dict(rows=rows_out, hm_rows=hm_out, max=mx, min=mn, avg=av)

    def fill_l2(self):
        check(_ffi.lib().kspec_dev_fill_l2(self._h))

    def sync(self):
        check(_ffi.lib().kspec_sync(self._h))

    def timer_start(self):
        check(_ffi.lib().kspec_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        check(_ffi.lib().kspec_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def kernel_times(self, cap=64):
        """ms per fused scan-kernel launch (most recent ``cap``), CUDA events on the plan's stream."""
        ms = (C.c_float * cap)()
        n = C.c_int(0)
        check(_ffi.lib().kspec_kernel_times(self._h, ms, cap, C.byref(n)))
        return [ms[i] for i in range(n.value)]

    def reserve_sms(self, n):
        """leave n SMs to concurrent kernels (the asynchronous NCCL exchange)"""
        check(_ffi.lib().kspec_plan_reserve_sms(self._h, int(n)))

    def launch_count(self):
        n = C.c_int64(0)
        check(_ffi.lib().kspec_launch_count(self._h, C.byref(n)))
        return n.value
