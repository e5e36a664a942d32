"""Host logic of the multi-GPU path on CPU: scan-range sharding, the global halving-average weights, and the
MAX / MIN / SUM combination -- once with plain numpy and once across two real processes (torch.distributed, gloo,
world_size 2), which is what kspec_comm_allreduce_stats does over NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest

from kspec.sharding import avg_weights, combine_stats, shard_bounds
from oracle import kspec_oracle as O


def test_shard_bounds_cover_everything_once():
    for n, w in ((146, 8), (2197, 4), (5, 8), (16384, 3), (1, 1)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
        sizes = [y - x for x, y in b]
        assert max(sizes) - min(sizes) <= 1


def test_avg_weights_equal_the_recurrence():
    rng = np.random.default_rng(0)
    for n in (1, 2, 15, 64, 100):
        x = rng.normal(size=(n, 7)) * 20 - 50
        a = None
        for row in x:
            a = O.cumulate("AVG", a, row)
        assert np.allclose(avg_weights(n) @ x, a, rtol=0, atol=1e-12)
        assert abs(avg_weights(n).sum() - 1.0) < 1e-15


def _partial(rows, a, b, n):
    """what kspec_zerospan_batch returns for shard [a,b) of n scans (float64 semantics)"""
    w = avg_weights(n)[a:b]
    return dict(max=rows[a:b].max(axis=0), min=rows[a:b].min(axis=0), avg=w @ rows[a:b])


def test_combine_equals_sequential():
    rng = np.random.default_rng(1)
    rows = rng.normal(size=(40, 33)) * 10 - 60
    ref = O.zerospan(10 ** ((rows + 19.1) / 10), 19.1, 33, "RAW")
    for w in (1, 2, 4, 8):
        parts = [_partial(rows, a, b, 40) for a, b in shard_bounds(40, w)]
        got = combine_stats(parts)
        assert np.allclose(got["max"], ref["max"], atol=1e-9) and np.allclose(got["min"], ref["min"], atol=1e-9)
        assert np.allclose(got["avg"], ref["avg"], atol=1e-9)


@pytest.mark.parametrize("R", [1.0, 0.5, 0.25])
def test_stitch_partials_sum_to_the_sequential_stitch(R):
    """closed-form weights of the sharded stepped scan == the reference's RAW-then-halving stitch (K:643-650)"""
    from kspec.sharding import stitch_partial
    F, fs = 64, 2.4e6
    geo = O.scan_geometry(30e6, 30e6 + 6 * fs, fs, F, R)
    _, total, steps = geo
    n = len(steps)
    rng = np.random.default_rng(5)
    lin = np.abs(rng.normal(size=(n, F))) + 1e-3
    ref = O.scan_pass(lin, [True] * n, geo, 19.1, O.scan_init_state(total, 19.1), 0)
    db = O.log_nogain(O.clip_min(lin), 19.1, inf_to=0)
    i_start = [s["i_start"] for s in steps]
    for w in (1, 2, 3, 5):
        parts = [stitch_partial(db[a:b], a, i_start, F, total) for a, b in shard_bounds(n, w) if b > a]
        assert np.max(np.abs(np.sum(parts, axis=0) - ref["cur"])) < 1e-10, (R, w)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    rows = rng.normal(size=(40, 33)) * 10 - 60
    a, b = shard_bounds(40, world)[rank]
    p = _partial(rows, a, b, 40)
    mx, mn, av = (torch.from_numpy(p[k].copy()) for k in ("max", "min", "avg"))
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(av, op=dist.ReduceOp.SUM)
    q.put((rank, mx.numpy(), mn.numpy(), av.numpy()))
    dist.destroy_process_group()


def test_two_process_gloo_allreduce_matches_single_process():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    rows = rng.normal(size=(40, 33)) * 10 - 60
    ref = O.zerospan(10 ** ((rows + 19.1) / 10), 19.1, 33, "RAW")
    for _, mx, mn, av in res:
        assert np.allclose(mx, ref["max"], atol=1e-9) and np.allclose(mn, ref["min"], atol=1e-9)
        assert np.allclose(av, ref["avg"], atol=1e-9)


def test_uncovered_bins_keep_the_previous_cur():
    """kspec_scan_shard writes 0 into bins that no step covers; after the SUM over shards the caller restores the previous
    Fft.Cur there, which is what the single-plan stitch (K:643-650 never touches such bins) leaves behind"""
    from kspec.sharding import keep_uncovered, scan_cover
    F, total = 8, 40
    i_start = [0, 4, 8, 12]                               # covers bins [0, 20): 20..39 are never written
    prev = np.arange(total, dtype=np.float64) - 100.0
    summed = np.where(np.arange(total) < 20, 7.0, 0.0)
    out = keep_uncovered(summed, prev, i_start, F)
    assert np.array_equal(out[:20], summed[:20]) and np.array_equal(out[20:], prev[20:])
    i0, i1 = scan_cover(i_start, F, total)
    assert (i0[:20] <= i1[:20]).all() and (i0[20:] > i1[20:]).all()


def test_bdata_switches_gate_what_reaches_the_dict():
    """bDataMax / bDataMin / bDataAvg (K:471-476): the batch always computes all three (they are carried between batches), the
    switches decide which of them are published in d['Fft.*']"""
    from kspec import hotpath
    d = {"bDataMax": True, "bDataMin": False, "bDataAvg": True, "Fft.Max": None, "Fft.Min": None, "Fft.Avg": None}
    out = {"max": np.ones(4), "min": -np.ones(4), "avg": np.zeros(4)}
    hotpath._publish_stats(d, out)
    assert d["Fft.Max"] is out["max"] and d["Fft.Avg"] is out["avg"] and d["Fft.Min"] is None
    st = hotpath._carried_state(d)
    assert st[0] is out["max"] and st[1] is out["min"] and st[2] is out["avg"]


def test_tiled_column_pass_arithmetic_model():
    """tools/tiled_cols_model.py: the tile walk, swizzle, twiddle recurrence and Z layout cols_tiled_kernel was written from give
    the transform of the whole frame, and the swizzle keeps both shared-memory access patterns conflict-free (host logic only;
    the kernel itself is compared with the element-wise pass and the oracle in the GPU tests)"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("tiled_cols_model", os.path.join(os.path.dirname(__file__), "..", "tools", "tiled_cols_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(3)
    for l1, l2 in ((8, 8), (9, 9)):
        x = rng.standard_normal(1 << (l1 + l2)) + 1j * rng.standard_normal(1 << (l1 + l2))
        ref = np.fft.fft(x)
        assert np.max(np.abs(m.four_step_tiled(x, l1, l2) - ref)) < 1e-12 * np.max(np.abs(ref))
    for l1 in (8, 9, 10):
        assert m.bank_conflict_free(l1)
