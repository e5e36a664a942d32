"""Two real GPUs, one process each: scan-range shards + NCCL MAX/MIN/SUM (kspec_comm_*) == the single-GPU result.
Skipped on boxes with one GPU (the sharding arithmetic itself is covered on CPU by tests/test_sharding_cpu.py and on one
GPU by test_sharded_partials_combine_to_the_single_gpu_result)."""
import os

import numpy as np
import pytest

from kspec.engine import device_count

pytestmark = pytest.mark.gpu

F, R, GAIN, XRES, N = 256, 0.5, 19.1, 64, 48


def _worker(rank, world, uid_path, q):
    import time
    import numpy as np
    from kspec import synth
    from kspec.comm import Comm
    from kspec.engine import Plan
    from kspec.sharding import shard_bounds
    S = F * 8
    if rank == 0:
        uid = Comm.unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.05)
        uid = open(uid_path, "rb").read()
    comm = Comm(world, rank, uid, rank)
    x = synth.tones_noise(N * S, seed=77, gate=(5000, 0.5))
    a, b = shard_bounds(N, world)[rank]
    with Plan(F, S, R, np.hanning(F), "AVG", precision="f64", device=rank) as plan:
        # host-vector path
        out = plan.zerospan_batch(x[a * S:b * S], b - a, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=N)
        comm.allreduce_host(out["max"], out["min"], out["avg"])
        # device-resident asynchronous path
        d = plan.dev_alloc((b - a) * S * 8)
        plan.dev_upload(d, x[a * S:b * S])
        plan.zerospan_batch_dev(d, b - a, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=N)
        comm.allreduce_plan_stats(plan)
        comm.join(plan)
        dev = plan.zerospan_fetch()
        # overlapped order: batch k, all-reduce (asynchronous), batch k+1 -> join must refuse (the plan now holds batch k+1),
        # the reduced vectors of batch k come from fetch_reduced, the plan's own statistics are those of batch k+1
        plan.zerospan_batch_dev(d, b - a, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=N)
        comm.allreduce_plan_stats(plan)
        half = (b - a) // 2
        plan.zerospan_batch_dev(d, half, GAIN, XRES, "MAX")
        from kspec._ffi import KspecError
        try:
            comm.join(plan)
            refused = False
        except KspecError:
            refused = True
        red = comm.fetch_reduced(F)
        nxt = plan.zerospan_fetch()
        plan.dev_free(d)
        local_half = plan.zerospan_batch(x[a * S:(a + half) * S], half, GAIN, XRES, "MAX")
        overl_ok = bool(refused and np.array_equal(nxt["max"], local_half["max"]) and np.array_equal(nxt["avg"], local_half["avg"]))
    # the exchange folded into the statistics kernel: peer-memory writes over NVLink (kspec_comm_peer_setup), three batches in
    # a row (both epochs of the symmetric buffer get reused), float64 and float32 plans
    peer_ok = True
    for prec in ("f64", "f32"):
        with Plan(F, S, R, np.hanning(F), "AVG", precision=prec, device=rank) as plan:
            comm.peer_setup(plan)
            d = plan.dev_alloc((b - a) * S * 8)
            plan.dev_upload(d, x[a * S:b * S])
            for _ in range(3):
                plan.zerospan_batch_dev(d, b - a, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=N)
                got = plan.zerospan_fetch()
            plan.dev_free(d)
            one = plan.zerospan_batch(x, N, GAIN, XRES, "MAX")            # not a shard: no exchange, this rank alone
            peer_ok = peer_ok and np.array_equal(got["max"], one["max"]) and np.array_equal(got["min"], one["min"])
            peer_ok = peer_ok and float(np.max(np.abs(got["avg"] - one["avg"]))) < (1e-9 if prec == "f64" else 1e-4)
    peer_ok = bool(peer_ok and not comm.peer_timed_out())
    # stepped scan sharded by frequency step: SUM of the stitch partials over NVLink
    from oracle import kspec_oracle as O
    Fs, rs = 64, 0.1
    Ss = Fs * 8
    geo = O.scan_geometry(30e6, 30e6 + 11 * 2.4e6, 2.4e6, Fs, 0.5)
    _, total, steps = geo
    ns = len(steps)
    bufs = np.concatenate([synth.step_tones(s, Ss) for s in range(ns)])
    sa, sb = shard_bounds(ns, world)[rank]
    with Plan(Fs, Ss, rs, np.ones(Fs), "AVG", precision="f64", device=rank) as plan:
        cur = plan.scan_shard(bufs[sa * Ss:sb * Ss], sb - sa, sa, [s["i_start"] for s in steps], total, O.MIN_AMP4CLIP, 19.1)
        comm.allreduce_sum(cur)
        st = O.scan_init_state(total, 19.1)
        plan.scan_stats_update(cur, steps[-1]["i_done"], 0, st)
    comm.close()
    q.put((rank, out["max"], out["min"], out["avg"], dev["max"], dev["min"], dev["avg"], st["cur"], st["max"], st["avg"],
           red[0], red[1], red[2], overl_ok, peer_ok))


@pytest.mark.skipif(device_count() < 2, reason="needs two GPUs")
def test_two_gpu_shards_match_single_gpu(tmp_path):
    import multiprocessing as mp
    from kspec import synth
    from kspec.engine import Plan
    S = F * 8
    x = synth.tones_noise(N * S, seed=77, gate=(5000, 0.5))
    with Plan(F, S, R, np.hanning(F), "AVG", precision="f64") as plan:
        ref = plan.zerospan_batch(x, N, GAIN, XRES, "MAX")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid_path = str(tmp_path / "nccl.uid")
    procs = [ctx.Process(target=_worker, args=(r, 2, uid_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import kspec_oracle as O
    geo = O.scan_geometry(30e6, 30e6 + 11 * 2.4e6, 2.4e6, 64, 0.5)
    _, total, steps = geo
    lin = [O.curscan(synth.step_tones(s, 512).astype(np.complex128), 64, 0.1, np.ones(64)) for s in range(len(steps))]
    sref = O.scan_pass(lin, [True] * len(steps), geo, 19.1, O.scan_init_state(total, 19.1), 0)
    for r in res:
        assert r[13], "overlapped order: join must refuse and leave the plan's batch k+1 statistics alone"
        assert r[14], "peer-memory exchange: the sharded batches must leave the statistics of the whole capture on every rank"
        for got in (r[1:4], r[4:7], r[10:13]):
            assert np.array_equal(got[0], ref["max"]) and np.array_equal(got[1], ref["min"])
            assert np.max(np.abs(got[2] - ref["avg"])) < 1e-9
        assert np.max(np.abs(r[7] - sref["cur"])) < 1e-9 and np.max(np.abs(r[8] - sref["max"])) < 1e-9
        assert np.max(np.abs(r[9] - sref["avg"])) < 1e-9
