#!/usr/bin/env python3
"""Can a tensor-core DFT stage hold 1e-3 dB?  Numerical model (CPU), BASELINE cfg-1 input.

The 2048-point frame transform is 16 x 128: stage 0 = 128 radix-16 butterflies = one [128 x 32] x [32 x 32] real GEMM
(complex DFT16 written out in re/im).  This script evaluates that stage the way tcgen05.mma kind::tf32 would -- operands
rounded to TF32 (10-bit mantissa), exact products, float32 accumulation over K = 8 chunks -- with 1, 2 and 3 operand-split
terms (x = x_hi + x_lo, F = F_hi + F_lo;  1: x_hi F_hi;  2: + x_lo F_hi;  3: + x_hi F_lo), everything AFTER the stage in
float64, and reports the error of the final cfg-1 spectrum (15 frames, hanning, cumulate AVG, dB) against the float64
reference, next to the same stage in plain float32 (what the SIMT kernel does).  Also fp16 operands (2-term split).

    python tools/tc_accuracy.py            (a few seconds; writes nothing)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))
from kspec import synth  # noqa: E402

F, S, R = 2048, 16384, 0.5


def to_tf32(a):
    """round to nearest even on a 10-bit mantissa (TF32), keep float32 storage"""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def to_f16(a):
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def split(a, rnd, terms):
    hi = rnd(a)
    if terms == 1:
        return [hi]
    lo = rnd(np.asarray(a, dtype=np.float32) - hi)
    return [hi, lo]


def gemm_tc(A, B, rnd, nterms, kchunk):
    """sum of the split products with float32 accumulation per K chunk (products of rounded operands are exact in float64)"""
    As, Bs = split(A, rnd, 2 if nterms >= 2 else 1), split(B, rnd, 2 if nterms >= 3 else 1)
    pairs = [(0, 0)] + ([(1, 0)] if nterms >= 2 else []) + ([(0, 1)] if nterms >= 3 else [])
    acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.float32)
    for ia, ib in pairs:
        a, b = As[ia].astype(np.float64), Bs[ib].astype(np.float64)
        for k0 in range(0, A.shape[1], kchunk):
            acc = (acc.astype(np.float64) + a[:, k0:k0 + kchunk] @ b[k0:k0 + kchunk]).astype(np.float32)
    return acc


def dft16_matrix():
    m = np.arange(16)
    th = 2 * np.pi * np.outer(m, m) / 16
    B = np.zeros((32, 32))
    B[0::2, 0::2] = np.cos(th)
    B[1::2, 0::2] = np.sin(th)
    B[0::2, 1::2] = -np.sin(th)
    B[1::2, 1::2] = np.cos(th)
    return B


def frame_spectrum(xw, stage0):
    """|FFT_2048(xw)| with stage 0 (DFT16 over m, n = j + 128 m) evaluated by `stage0`, the rest in float64"""
    X = xw.reshape(16, 128).T                                     # X[j][m]
    A = np.empty((128, 32))
    A[:, 0::2], A[:, 1::2] = X.real, X.imag
    D = stage0(A)
    Y = D[:, 0::2].astype(np.float64) + 1j * D[:, 1::2].astype(np.float64)       # Y[j][k1]
    j, k1 = np.arange(128)[:, None], np.arange(16)[None, :]
    Y = Y * np.exp(-2j * np.pi * j * k1 / F)
    Z = np.fft.fft(Y, axis=0)                                     # over j -> k2;  bin = k1 + 16 k2
    return np.abs(Z).reshape(-1)                                  # index k2*16 + k1 = bin


def scan_db(x, win, stage0):
    offs = [int(i * F * R) for i in range(int(S / (F * R)))]
    offs = [o for o in offs if o + F <= S]
    acc = None
    for o in offs:
        m = frame_spectrum(x[o:o + F] * win, stage0) * (F / win.sum()) * 2 / F
        acc = m if acc is None else (acc + m) / 2
    return 10 * np.log10(np.fft.fftshift(acc))


def main():
    n_scans = 6
    x = synth.tones_noise(n_scans * S, seed=1).astype(np.complex128)
    win = np.hanning(F)
    B = dft16_matrix()
    variants = [("float64 (reference)", lambda A: A @ B),
                ("float32 FMA chain (SIMT stage)", lambda A: (A.astype(np.float32) @ B.astype(np.float32))),
                ("tcgen05 kind::tf32, 1 term", lambda A: gemm_tc(A, B, to_tf32, 1, 8)),
                ("tcgen05 kind::tf32, 2 terms (x_hi, x_lo) F_hi", lambda A: gemm_tc(A, B, to_tf32, 2, 8)),
                ("tcgen05 kind::tf32, 3 terms", lambda A: gemm_tc(A, B, to_tf32, 3, 8)),
                ("tcgen05 kind::f16, 1 term", lambda A: gemm_tc(A, B, to_f16, 1, 16)),
                ("tcgen05 kind::f16, 3 terms", lambda A: gemm_tc(A, B, to_f16, 3, 16))]
    ref = None
    print("%-48s %12s %12s" % ("stage 0 arithmetic", "max |dB err|", "p99 |dB err|"))
    for name, fn in variants:
        rows = np.array([scan_db(x[k * S:(k + 1) * S], win, fn) for k in range(n_scans)])
        if ref is None:
            ref = rows
            continue
        e = np.abs(rows - ref)
        print("%-48s %12.3e %12.3e" % (name, e.max(), np.percentile(e, 99)))


if __name__ == "__main__":
    main()
