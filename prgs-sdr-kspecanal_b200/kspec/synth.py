"""Synthetic IQ sources for headless runs, tests and the bench.

The reference has one fake device, ``python/testfft.py:13-81`` (a tone generator that stands in for
``rtlsdr.RtlSdr``); it is bit-rotted (``testfft.py:66`` passes a float ``num`` to ``np.linspace``).
This module is the replacement: seeded tones + white noise, optional uint8 quantisation in the
``rtl_sdr`` raw-file layout (interleaved I,Q bytes, ``octave/load_rtlsdr.m:8-12``), and a file/array
backed ``RtlSdr`` look-alike that honours the read pattern of ``sdr_read`` (``kspecanal.py:311-347``).

Nothing here touches the GPU; it only produces host arrays.
"""
import numpy as np

DEFAULT_AMPS = (0.5, 0.25, 0.05)
DEFAULT_FREQS = (300e3, -700e3, 1.0e6)
DEFAULT_SIGMA = 0.01

# pyrtlsdr's packed-bytes -> complex convention is (b - 127.5) / 127.5  (== b/127.5 - 1); the reference
# repo's other conventions are (b-127)/128 (kspecanal.old.py:126-135) and b-127 (octave/load_rtlsdr.m:11).
U8_OFFSET = 127.5
U8_SCALE = 1.0 / 127.5


def tones_noise(n, seed, fs=2.4e6, amps=DEFAULT_AMPS, freqs=DEFAULT_FREQS, sigma=DEFAULT_SIGMA,
                t0=0, dtype=np.complex64, gate=None):
    """n complex samples: sum_k A_k exp(2 pi j f_k t) + sigma (N(0,1) + j N(0,1)), t = (t0 + i)/fs.

    ``gate``: optional (period_samples, duty) that switches the first tone on/off in time
    (used by the long-capture config so that Max/Min/Avg differ).
    """
    rng = np.random.default_rng(seed)
    idx = np.arange(t0, t0 + n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    for k, (a, f) in enumerate(zip(amps, freqs)):
        # phase accumulated in cycles, reduced mod 1 before the 2 pi to keep float64 exact-ish
        cyc = np.mod(idx * (f / fs), 1.0)
        tone = a * np.exp(2j * np.pi * cyc)
        if gate is not None and k == 0:
            period, duty = gate
            tone = tone * ((np.mod(idx, period) < duty * period).astype(np.float64))
        x += tone
    if sigma:
        x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(dtype)


def to_u8_iq(x):
    """complex -> interleaved uint8 I,Q (rtl_sdr raw layout), round((v+1)*127.5) clipped to 0..255."""
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    return np.clip(np.rint((iq + 1.0) * 127.5), 0, 255).astype(np.uint8)


def from_u8_iq(b, offset=U8_OFFSET, scale=U8_SCALE, dtype=np.complex128):
    """interleaved uint8 I,Q -> complex, (b - offset) * scale per component (float64 arithmetic)."""
    v = (b.astype(np.float64) - offset) * scale
    return (v[0::2] + 1j * v[1::2]).astype(dtype)


def step_tones(step, n, fs=2.4e6, sigma=DEFAULT_SIGMA, dtype=np.complex64):
    """Per-step buffer for stepped scans: one tone whose offset depends on the step index
    (seed = step index), so every step of a stitched range is distinguishable."""
    f = ((step * 37) % 19 - 9) * (fs / 24.0) + fs / 96.0
    a = 0.1 + 0.05 * (step % 7)
    return tones_noise(n, seed=step, fs=fs, amps=(a,), freqs=(f,), sigma=sigma, dtype=dtype)


class ArrayRtlSdr:
    """Array/file backed stand-in for ``rtlsdr.RtlSdr`` (duck type used at ``kspecanal.py:287-347``).

    * ``read_samples(n)`` hands out consecutive complex128 samples of the capture (``n`` may be a
      float, ``kspecanal.py:343-346``); a read that starts at the end raises ``EOFError``, one that
      merely runs past it is zero padded.
    * every retune (``sample_rate``/``center_freq``/``gain`` assignment, ``kspecanal.py:297-299``) arms
      a "settling" flag: the following read (the 16Ki discard of ``sdr_setup``, ``kspecanal.py:301``)
      returns zeros and does not consume the capture, so scan ``k`` of a capture is always
      samples ``[k*fullSize, (k+1)*fullSize)``.
    * ``per_tune``: optional callable ``(tune_index, center_freq, n) -> complex array`` that supplies a
      fresh buffer after every retune (stepped scans).
    """

    valid_gains_db = [0.0, 19.1, 48.0]
    bandwidth = 0
    freq_correction = 0

    def __init__(self, capture=None, per_tune=None, fail_tunes=()):
        self._cap = None if capture is None else np.asarray(capture)
        self._pos = 0
        self._per_tune = per_tune
        self._tune = -1
        self._buf = None
        self._bufpos = 0
        self._settle = False
        self._fail = set(fail_tunes)
        self._fs = 2.4e6
        self._fc = 92e6
        self._gain = 19.1

    @classmethod
    def from_u8_file(cls, path, offset=U8_OFFSET, scale=U8_SCALE):
        return cls(from_u8_iq(np.fromfile(path, dtype=np.uint8), offset, scale))

    # -- tuning ----------------------------------------------------------------------------------
    @property
    def sample_rate(self):
        return self._fs

    @sample_rate.setter
    def sample_rate(self, v):
        self._fs = v
        self._settle = True

    @property
    def center_freq(self):
        return self._fc

    @center_freq.setter
    def center_freq(self, v):
        self._fc = v
        self._settle = True
        self._tune += 1
        if self._tune in self._fail:
            raise IOError("synthetic tune failure at tune %d" % self._tune)
        if self._per_tune is not None:
            self._buf = None

    @property
    def gain(self):
        return self._gain

    @gain.setter
    def gain(self, v):
        self._gain = v
        self._settle = True

    # -- data ------------------------------------------------------------------------------------
    def read_samples(self, n):
        n = int(n)
        if self._settle:
            self._settle = False
            return np.zeros(n, dtype=np.complex128)
        if self._per_tune is not None:
            if self._buf is None:
                self._buf = np.asarray(self._per_tune(self._tune, self._fc, n))
                self._bufpos = 0
            out = self._buf[self._bufpos:self._bufpos + n]
            self._bufpos += n
        else:
            out = self._cap[self._pos:self._pos + n]
            self._pos += n
        if len(out) == 0:
            raise EOFError("capture exhausted")
        if len(out) < n:
            # sdr_read rounds a non power-of-two tail UP and discards the excess (K:340-346); a finite
            # capture answers the over-read with zero padding, which the caller then drops.
            out = np.concatenate([out, np.zeros(n - len(out), dtype=out.dtype)])
        return out.astype(np.complex128)

    def close(self):
        pass
