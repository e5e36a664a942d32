#!/usr/bin/env python3
"""Index-level model of the R32 frame transform (csrc/curscan_r32.cuh): 2048 = 32 x 2 x 32.

A team of 64 threads owns one 2048-point frame, 32 complex values per thread:
  stage 0   thread j holds x[j + 64 m]; DFT32 over m, boundary twiddle W_2048^(j k1)
  radix 2   partner threads j, j+32 (lane xor 16) swap half of their values; u = a + b, v = (a - b) W_64^j
  exchange  one pass through shared memory
  stage 1   thread rho holds row rho over j' = 0..31; DFT32 -> bins rho + 64 kappa
Run: python tools/r32_model.py   (asserts against numpy.fft)."""
import numpy as np

F, NT, P = 2048, 64, 32


def lane_maps():
    t = np.arange(NT)
    w, l = t >> 5, t & 31
    h, i = l >> 4, l & 15
    j = i + 16 * w + 32 * h
    return t, h, j, j % 32


def tables(win):
    t, h, j, jp = lane_maps()
    m = np.arange(P)
    # window with the signs that rotate the stage-0 outputs of the upper threads by 16 slots
    wtab = win[j[:, None] + 64 * m[None, :]] * np.where(h[:, None] == 1, (-1.0) ** m[None, :], 1.0)
    s = np.arange(P)
    k1 = (s[None, :] + 16 * h[:, None]) % 32                     # what slot s of thread t holds after stage 0
    tw = np.exp(-2j * np.pi * (j[:, None] * k1) / F)             # TW[t][s]
    omega = np.where(h == 1, -1.0, 1.0) * np.exp(-2j * np.pi * jp / 64)
    return wtab, tw, omega


def frame_fft(x, win):
    t, h, j, jp = lane_maps()
    wtab, tw, omega = tables(win)
    m = np.arange(P)
    b = x[j[:, None] + 64 * m[None, :]] * wtab                   # [t][m]
    b = np.fft.fft(b, axis=1)                                    # DFT32 over m; the sign trick rotates the upper threads' slots
    b = b * tw
    recv = b[t ^ 16][:, 16:]                                     # shfl.xor 16 of slots 16..31
    own = b[:, :16]
    u = own + recv
    v = (own - recv) * omega[:, None]
    S = np.zeros((64, 32), complex)                              # exchange: row rho, column j'
    for tt in range(NT):
        for s in range(16):
            k1 = s + 16 * h[tt]
            S[k1, jp[tt]] = u[tt, s]
            S[k1 + 32, jp[tt]] = v[tt, s]
    X = np.fft.fft(S, axis=1)                                    # thread rho: DFT32 over j' -> slot kappa
    out = np.zeros(F, complex)
    for rho in range(64):
        out[rho + 64 * np.arange(32)] = X[rho]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(F) + 1j * rng.standard_normal(F)
    for win in (np.ones(F), np.hanning(F), np.kaiser(F, 64)):
        ref = np.fft.fft(x * win)
        got = frame_fft(x, win)
        err = np.max(np.abs(ref - got)) / np.max(np.abs(ref))
        assert err < 1e-12, err
    print("R32 index model matches numpy.fft")
