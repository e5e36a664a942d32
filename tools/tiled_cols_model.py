#!/usr/bin/env python3
"""Index-arithmetic model of cols_tiled_kernel (prgs-sdr-kspecanal_b200/csrc/bigfft_kernels.cuh): the tile walk, the swizzled
slot function, the in-place column transforms, the twiddle recurrence W^(n2*tid) * (W^(n2*NT))^m and the Z layout, in numpy,
followed by the row pass, against np.fft.fft of the whole frame.  It checks the arithmetic the kernel was written from (and that
both shared-memory access patterns are free of bank conflicts for 16-byte elements), not the kernel: that is what
tests/test_gpu_parity.py::test_fourstep_tiled_column_pass_matches_the_elementwise_one and test_big_engines do on the GPU."""
import numpy as np


def tile_cfg(l1):
    log2tc = 12 - l1
    tc = 1 << log2tc
    shift = 0 if tc >= 8 else (1 if tc == 4 else 2)
    mask = (8 if tc >= 8 else tc) - 1
    return log2tc, tc, lambda e, c: e * tc + (c ^ ((e >> shift) & mask))


def four_step_tiled(x, l1, l2):
    L1, L2 = 1 << l1, 1 << l2
    M = L1 * L2
    log2tc, TC, slot = tile_cfg(l1)
    NT = L1 // 16                                   # threads per team, 16 points each
    tw = np.exp(-2j * np.pi * np.arange(M) / M)
    Z = np.zeros(M, complex)
    for t in range(L2 >> log2tc):
        n2_0 = t << log2tc
        tile = np.full(L1 * TC, np.nan, complex)
        for i in range(L1 * TC):                    # phase 1: rows of the tile, columns fastest
            e, c = i >> log2tc, i & (TC - 1)
            tile[slot(e, c)] = x[(e << l2) + n2_0 + c]
        for c in range(TC):                         # phase 2: in place, bins times W^(n2*k) by recurrence
            idx = np.array([slot(e, c) for e in range(L1)])
            X = np.fft.fft(tile[idx])
            n2 = n2_0 + c
            for tid in range(NT):
                w, step = tw[n2 * tid], tw[n2 * NT]
                for m in range(16):
                    X[tid + NT * m] *= w
                    w = w * step
            tile[idx] = X
        for i in range(L1 * TC):                    # phase 3: rows of the tile to Z
            k, c = i >> log2tc, i & (TC - 1)
            Z[(k << l2) + n2_0 + c] = tile[slot(k, c)]
    out = np.zeros(M, complex)
    for k1 in range(L1):                            # row pass: bin k1 + L1*k2
        out[k1 + L1 * np.arange(L2)] = np.fft.fft(Z[k1 * L2:(k1 + 1) * L2])
    return out


def bank_conflict_free(l1):
    """quarter warps (8 lanes x 16 bytes = all 32 banks): column reads of a team and row-wise accesses of the CTA"""
    L1 = 1 << l1
    log2tc, TC, slot = tile_cfg(l1)
    for c in range(TC):
        for e0 in range(0, L1, 8):
            if len({slot(e, c) % 8 for e in range(e0, e0 + 8)}) != 8:
                return False
    for i0 in range(0, L1 * TC, 8):
        if len({slot(i >> log2tc, i & (TC - 1)) % 8 for i in range(i0, i0 + 8)}) != 8:
            return False
    return True


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for l1, l2 in ((8, 8), (8, 9), (9, 9), (10, 10)):
        x = rng.standard_normal(1 << (l1 + l2)) + 1j * rng.standard_normal(1 << (l1 + l2))
        ref = np.fft.fft(x)
        err = np.max(np.abs(four_step_tiled(x, l1, l2) - ref)) / np.max(np.abs(ref))
        print("L1 = 2^%d, L2 = 2^%d: max relative error %.2e, bank-conflict free: %s" % (l1, l2, err, bank_conflict_free(l1)))
