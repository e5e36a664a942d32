"""ORACLE — CPU restatement (numpy, float64/complex128) of kspecanal's spectrum hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``prgs-sdr-kspecanal_b200/kspec``) never does; it fails loudly when ``libkspec.so`` is missing.

Parity status: the reference ships no tests / golden vectors, so "pinned" here
means: every function below is checked against the UNMODIFIED reference functions executed in the
build container (``oracle/ref_loader.py`` + ``oracle/make_golden.py``), and the resulting vectors are
committed under ``tests/golden/`` (``tests/test_oracle_golden.py`` re-checks them without the
reference).  The uint8 -> complex scale of the un-vendored ``pyrtlsdr`` (version unpinned by the
reference) is "parity unpinned": it is a parameter everywhere.

All ``K:`` citations are ``/root/reference/python/kspecanal.py`` line numbers.
"""
import numpy as np

CUMU_RAW, CUMU_AVG, CUMU_MAX, CUMU_MIN = "RAW", "AVG", "MAX", "MIN"
MIN_AMP4CLIP = (1 / 256) * 0.00001          # K:53
HEATMAP_ROWS = 128                          # K:448, K:611


# ------------------------------------------------------------------------------------------------
# derived configuration (K:926-949)
# ------------------------------------------------------------------------------------------------
def full_size(fft_size, sampling_rate):
    """K:926-929: capture 8 frames worth when fftSize is small against the rate, else 2."""
    return fft_size * 8 if fft_size < (sampling_rate // 8) else fft_size * 2


def window_table(name, fft_size):
    """K:932-936.  ``name`` is the CLI value (ones|hanning|hamming|kaiser), case-insensitive."""
    name = name.upper().replace("WIN.", "")
    if name == "HAMMING":
        return np.hamming(fft_size)
    if name == "HANNING":
        return np.hanning(fft_size)
    if name == "KAISER":
        return np.kaiser(fft_size, 64)
    if name == "ONES":
        return np.ones(fft_size)
    raise KeyError(name)


def adjust_xres(fft_size, x_res):
    """K:938-949 without the blocking input(): clamp to fftSize, else largest sub-multiple with
    at least ~300 points when fftSize is not a multiple of xRes."""
    if x_res > fft_size:
        return fft_size
    if fft_size % x_res != 0:
        for i in range(int(fft_size / 300), 0, -1):
            if fft_size % i == 0:
                return fft_size // i
    return x_res


# ------------------------------------------------------------------------------------------------
# frames (K:368, K:385-390)
# ------------------------------------------------------------------------------------------------
def frame_offsets(fft_size, full, non_overlap):
    """Start index of every frame sdr_curscan actually transforms.

    numLoops = int(S/(F*r)) (K:368); frame i starts at int(i*F*r) -- the product is evaluated left to
    right in float64, (i*F) exact then *r, then truncated (K:386); the loop stops at the first frame
    that would run past the capture (K:388-390)."""
    n_loops = int(full / (fft_size * non_overlap))
    offs = []
    for i in range(n_loops):
        start = int(i * fft_size * non_overlap)
        if start + fft_size > full:
            break
        offs.append(start)
    return np.asarray(offs, dtype=np.int64)


# ------------------------------------------------------------------------------------------------
# cumulate (K:124-147)
# ------------------------------------------------------------------------------------------------
def cumulate(mode, cur, new):
    """One data_cumu step on whole arrays: None -> copy; RAW overwrite; AVG = (cur+new)/2 (a halving
    recurrence, not a mean); MAX / MIN elementwise."""
    if cur is None:
        return np.array(new, dtype=np.float64, copy=True)
    if mode == CUMU_RAW:
        return np.array(new, dtype=np.float64, copy=True)
    if mode == CUMU_AVG:
        return (cur + new) / 2
    if mode == CUMU_MAX:
        return np.maximum(cur, new)
    if mode == CUMU_MIN:
        return np.minimum(cur, new)
    raise ValueError(mode)


# ------------------------------------------------------------------------------------------------
# the core (K:351-397)
# ------------------------------------------------------------------------------------------------
def curscan(samples, fft_size, non_overlap, win, cumu_mode=CUMU_AVG):
    """sdr_curscan on an in-memory capture of fullSize complex samples -> float64[F], linear, shifted.

    per frame  winAdj*2*|FFT(x*w)|/F  with winAdj = F/sum(w)  (K:372-373, K:391); cumulated by
    ``cumu_mode`` (K:392-395); finally fftshift (K:396)."""
    samples = np.asarray(samples, dtype=np.complex128)
    win = np.asarray(win, dtype=np.float64)
    win_adj = len(win) / np.sum(win)
    acc = None
    for start in frame_offsets(fft_size, len(samples), non_overlap):
        seg = samples[start:start + fft_size]
        mag = win_adj * 2 * np.abs(np.fft.fft(seg * win)) / fft_size
        acc = cumulate(cumu_mode, acc, mag)
    return np.fft.fftshift(acc)


def psd_segments(fft_size, full_size, non_overlap):
    """segment starts of the bUsePSD branch (K:375, K:381): noverlap = fftSize*(1-curScanNonOverlap) is a float;
    matplotlib.mlab's segmenting truncates it (``int(noverlap)``), steps by NFFT - noverlap and keeps
    (len(x) - noverlap) // step segments."""
    noverlap = int(fft_size * (1 - non_overlap))
    step = fft_size - noverlap
    return [i * step for i in range((full_size - noverlap) // step)]


def curscan_psd(samples, fft_size, non_overlap, win, fs=2.0):
    """sdr_curscan with bUsePSD (K:374-384): ``plt.psd(samples, NFFT=fftSize, window=win, noverlap=noverlap)[0]``.

    matplotlib is NOT under /root/reference and not installed here (parity unpinned by the reference; pinned
    operationally against scipy.signal.welch in tests/test_oracle_golden.py).  Restated from
    matplotlib.mlab._spectral_helper / csd / psd (3.x): complex input -> two-sided, no detrend, pad_to = NFFT,
    per segment conj(X)*X with X = FFT(x*w), divided by Fs (default 2) and by sum(w^2) (scale_by_freq), mean over the
    segments, real part, rolled so that frequency 0 sits at index NFFT//2 (= fftshift)."""
    samples = np.asarray(samples, dtype=np.complex128)
    win = np.asarray(win, dtype=np.float64)
    acc = np.zeros(fft_size)
    starts = psd_segments(fft_size, len(samples), non_overlap)
    for s in starts:
        X = np.fft.fft(samples[s:s + fft_size] * win)
        acc += (np.conj(X) * X).real / fs / np.sum(win ** 2)
    return np.fft.fftshift(acc / len(starts))


# ------------------------------------------------------------------------------------------------
# display processing (K:88-121, K:150-165)
# ------------------------------------------------------------------------------------------------
def clip_min(vals, min_amp=MIN_AMP4CLIP):
    """'Clip2MinAmp' (K:100-101)."""
    return np.clip(vals, min_amp, None)


def log_nogain(vals, gain, inf_to=None):
    """'LogNoGain' (K:106-112): 10*log10(amplitude) - gain; optional +-inf replacement."""
    with np.errstate(divide="ignore"):
        out = 10 * np.log10(vals) - gain
    if inf_to is not None:
        out[np.isinf(out)] = inf_to
    return out


def plotcompress(data, x_res, mode):
    """_data_plotcompress (K:168-202): RAW identity; MAX/AVG over x_res groups of adjacent bins.
    (MIN is documented but unreachable in the reference, K:188; offered here for row N3.)"""
    mode = mode.upper()
    if mode == "RAW":
        return data
    cols = len(data) // x_res
    if cols == 0:
        return data
    t = np.asarray(data)[: x_res * cols].reshape(x_res, cols)
    if mode == "MAX":
        return np.max(t, axis=1)
    if mode == "AVG":
        return np.average(t, axis=1)
    if mode == "MIN":
        return np.min(t, axis=1)
    raise ValueError(mode)


def heatmap_width(fft_size, x_res, hm_mode):
    """K:449-457."""
    if hm_mode.upper() in ("MAX", "MIN", "AVG") and fft_size > x_res:
        return x_res
    return fft_size


# ------------------------------------------------------------------------------------------------
# zeroSpan outer loop, compute lines only (K:464-484)
# ------------------------------------------------------------------------------------------------
def zerospan(lin_rows, gain, x_res, hm_mode="MAX", adj=None, state=None):
    """Run the zero_span loop body over a sequence of sdr_curscan outputs (linear, shifted).

    Returns dict(cur_rows[n,F] dB, hm_rows[n,W], max, min, avg).  ``state`` = (max, min, avg) carried
    in from earlier scans (None = fresh, K:438-441).  dB has no low clip here (K:469)."""
    mx, mn, av = state if state is not None else (None, None, None)
    cur_rows, hm_rows = [], []
    for lin in lin_rows:
        pr = log_nogain(np.asarray(lin, dtype=np.float64), gain)        # K:469
        mx = cumulate(CUMU_MAX, mx, pr)                                  # K:471-472
        mn = cumulate(CUMU_MIN, mn, pr)                                  # K:473-474
        av = cumulate(CUMU_AVG, av, pr)                                  # K:475-476
        pr_tmp = pr - adj if adj is not None else pr                     # K:400-411
        hm_rows.append(np.array(plotcompress(pr_tmp, x_res, hm_mode)))   # K:480
        cur_rows.append(pr)
    return dict(cur_rows=np.array(cur_rows), hm_rows=np.array(hm_rows), max=mx, min=mn, avg=av)


# ------------------------------------------------------------------------------------------------
# stepped scan (K:569-709)
# ------------------------------------------------------------------------------------------------
def fixup_scan_range(start_freq, end_freq, sampling_rate):
    """_fixupfreqs_scanrange (K:701-709): round the end up to a whole number of bands."""
    bands = (end_freq - start_freq) / sampling_rate
    if (bands % 1) != 0:
        end_freq = start_freq + np.ceil(bands) * sampling_rate
    return start_freq, end_freq, start_freq + (end_freq - start_freq) / 2


def scan_geometry(start_freq, end_freq, sampling_rate, fft_size, range_non_overlap):
    """Index arithmetic of _scan_range (K:594-600, K:621-629, K:688-689).

    Returns (num_groups, total_entries, steps) with steps = list of dicts
    (i, cur_freq, i_start, i_end, i_done, s_end)."""
    span = sampling_rate
    num_groups = int((end_freq - start_freq) / span)
    total = num_groups * fft_size
    cur_freq = start_freq + span / 2
    s_freq = cur_freq - span / 2
    steps = []
    i = 0
    while s_freq < end_freq:
        i_start = int(i * fft_size * range_non_overlap)
        i_end = i_start + fft_size
        i_done = int((i + 1) * fft_size * range_non_overlap)
        s_end = fft_size - max(0, i_end - total)
        steps.append(dict(i=i, cur_freq=cur_freq, i_start=i_start, i_end=i_end, i_done=i_done, s_end=s_end))
        cur_freq += span * range_non_overlap
        s_freq = cur_freq - span / 2
        i += 1
    return num_groups, total, steps


def scan_init_state(total, gain, min_amp=MIN_AMP4CLIP):
    """First-call initialisation (K:602-608): Cur/Max/Avg = dB(minAmp4Clip)-gain, Min = dB(1)-gain."""
    floor = log_nogain(np.ones(total) * min_amp, gain, inf_to=0)
    return dict(cur=floor.copy(), max=floor.copy(), avg=floor.copy(),
                min=log_nogain(np.ones(total), gain, inf_to=0))


def scan_pass(lin_rows, step_ok, geometry, gain, state, pass_index, min_amp=MIN_AMP4CLIP, base_is_raw=False):
    """One full pass of _scan_range's step loop (K:619-668) given every step's sdr_curscan output.

    ``lin_rows[i]`` linear shifted spectrum of step i (ignored when ``step_ok[i]`` is false: the
    reference substitutes ones(F), K:635-639).  ``state`` is updated in place and returned."""
    num_groups, total, steps = geometry
    avg_mode = CUMU_RAW if pass_index == 0 else CUMU_AVG              # K:615-618
    cur, mx, mn, av = state["cur"], state["max"], state["min"], state["avg"]
    i_old_end = 0
    for st in steps:
        i, i_start, i_end, i_done, s_end = st["i"], st["i_start"], st["i_end"], st["i_done"], st["s_end"]
        fft_size = i_end - i_start
        lin = np.asarray(lin_rows[i], dtype=np.float64) if step_ok[i] else np.ones(fft_size)
        pr = log_nogain(clip_min(lin, min_amp), gain, inf_to=0)        # K:640-641
        # RAW copy of the not-yet-seen tail, then halving average over the overlap (K:643-650)
        s_raw = fft_size - (i_end - i_old_end)
        cur[i_old_end:i_end] = pr[s_raw:s_end]
        if i_old_end != 0:
            if i_old_end > total:
                i_old_end = total
            cur[i_start:i_old_end] = (cur[i_start:i_old_end] + pr[0:i_old_end - i_start]) / 2
        i_old_end = i_end
        if base_is_raw:                                                # K:651-656
            d0, d1, src = i_start, i_end, pr[0:s_end]
        else:                                                          # K:657-662
            d0, d1, src = i_start, i_done, cur[i_start:i_done]
        n = len(mx[d0:d1])
        src = src[:n]
        mx[d0:d1] = np.maximum(mx[d0:d1], src)                         # K:663-664
        mn[d0:d1] = np.minimum(mn[d0:d1], src)                         # K:665-666
        if avg_mode == CUMU_RAW:                                       # K:667-668
            av[d0:d1] = src
        else:
            av[d0:d1] = (av[d0:d1] + src) / 2
    return state


def scan_freq_axis(start_freq, sampling_rate, fft_size, num_groups):
    """K:609."""
    total = num_groups * fft_size
    span = num_groups * sampling_rate
    return np.fft.fftshift(np.fft.fftfreq(total, 1 / span) + start_freq + span / 2)


def zerospan_freq_axis(center_freq, sampling_rate, fft_size):
    """K:444-445."""
    return np.fft.fftshift(np.fft.fftfreq(fft_size, 1 / sampling_rate) + center_freq)


# ------------------------------------------------------------------------------------------------
# chunked device read (K:311-347)
# ------------------------------------------------------------------------------------------------
def sdr_read_plan(length, unit=2 ** 18):
    """The (request, keep) pairs sdr_read issues for ``length`` samples: whole ``unit`` reads, then one
    read rounded UP to a power of two of which only ``remaining`` samples are kept (K:340-346)."""
    plan = []
    if length > unit:
        plan += [(unit, unit)] * (length // unit)
        remaining = length % unit
    else:
        remaining = length
    if remaining > 0:
        plan.append((int(2 ** np.ceil(np.log2(remaining))), remaining))
    return plan
