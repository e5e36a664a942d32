"""Scan-range / step-range sharding arithmetic for multi-GPU runs.  Pure host integer logic, no GPU.

zeroSpan: the unit is one scan (fullSize contiguous samples, K:370); overlap exists only inside a scan, so shards
need no halo.  Each rank gets a contiguous range; Max/Min combine with MAX/MIN, Avg partials are pre-weighted with
the global halving weights (data_cumu AVG, K:137-139) so that a plain SUM reproduces the sequential recurrence.
"""
import numpy as np


def shard_bounds(n_units, n_ranks):
    """contiguous [a, b) per rank, sizes differ by at most one, earlier ranks get the larger shards"""
    base, extra = divmod(int(n_units), int(n_ranks))
    out, a = [], 0
    for r in range(n_ranks):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def avg_weights(n_total):
    """weights w_k of the halving recurrence A_0 = x_0, A_k = (A_{k-1} + x_k)/2 after n_total items (float64)"""
    k = np.arange(n_total)
    w = np.ldexp(1.0, -(n_total - k).astype(np.int64))
    if n_total:
        w[0] = np.ldexp(1.0, -(n_total - 1))
    return w


def combine_stats(parts):
    """host-side equivalent of kspec_comm_allreduce_stats over a list of per-shard dicts (max, min, avg partial)"""
    return dict(max=np.max([p["max"] for p in parts], axis=0), min=np.min([p["min"] for p in parts], axis=0),
                avg=np.sum([p["avg"] for p in parts], axis=0))


def scan_cover(i_start, fft_size, total):
    """first / last step covering every bin of a stepped scan (steps cover [iStart, iStart+F), K:622-624)"""
    i_start = np.asarray(i_start, dtype=np.int64)
    b = np.arange(total, dtype=np.int64)
    i1 = np.searchsorted(i_start, b, side="right") - 1
    i0 = np.searchsorted(i_start + fft_size, b, side="right")
    return i0, i1


def stitch_partial(db_rows, step_base, i_start, fft_size, total):
    """numpy statement of kspec_scan_shard's arithmetic (host-side reference for the sharding logic, not a product path):
    the share of steps [step_base, step_base+len(db_rows)) in the stitched Fft.Cur.  SUM over shards == K:643-650."""
    i0, i1 = scan_cover(i_start, fft_size, total)
    out = np.zeros(total)
    for k, row in enumerate(db_rows):
        i = step_base + k
        b = np.arange(i_start[i], min(i_start[i] + fft_size, total))
        sh = np.where(i == i0[b], i1[b] - i0[b], i1[b] - i + 1)
        out[b] += np.ldexp(np.asarray(row, dtype=np.float64)[:len(b)], -sh.astype(np.int64))
    return out


def keep_uncovered(cur_sum, cur_prev, i_start, fft_size):
    """after the SUM over shards of kspec_scan_shard's partials: bins that no step covers keep the previous Fft.Cur, as the
    single-plan stitch does (K:643-650 never touches them); none exist in the reference's own geometries (K:598-600)"""
    i0, i1 = scan_cover(i_start, fft_size, len(cur_sum))
    unc = (i1 < 0) | (i0 > i1)
    out = np.array(cur_sum, dtype=np.float64, copy=True)
    out[unc] = np.asarray(cur_prev, dtype=np.float64)[unc]
    return out
