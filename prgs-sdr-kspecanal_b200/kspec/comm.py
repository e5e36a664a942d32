"""Multi-GPU exchange of the per-bin statistics (one process per GPU, NCCL over NVLink/NVSwitch).

Captures are sharded by scan range (``kspec.sharding``); every rank runs its shard with
``Plan.zerospan_batch(..., scan_index_base=a, n_scans_total=n)`` and then calls ``allreduce_*`` once:
MAX on Fft.Max, MIN on Fft.Min, SUM on the pre-weighted Fft.Avg partials.  Waterfall / Cur rows stay on the rank
that produced them.  The reference has no counterpart (it is a single process, kspecanal.py:1139-1155).
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import check, dptr


class Comm:
    @staticmethod
    def unique_id():
        """128-byte NCCL id; rank 0 creates it and ships it to the other ranks (any out-of-band channel)."""
        buf = C.create_string_buffer(128)
        check(_ffi.lib().kspec_comm_unique_id(buf))
        return buf.raw

    def __init__(self, n_ranks, rank, uid, device):
        self._h = C.c_void_p()
        assert len(uid) == 128
        check(_ffi.lib().kspec_comm_init(C.byref(self._h), int(n_ranks), int(rank), C.create_string_buffer(uid, 128), int(device)))
        self.n_ranks, self.rank = n_ranks, rank

    def allreduce_host(self, mx, mn, av):
        """in place on float64 host vectors"""
        for a in (mx, mn, av):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        check(_ffi.lib().kspec_comm_allreduce_stats(self._h, dptr(mx), dptr(mn), dptr(av), len(mx)))

    def allreduce_sum(self, v):
        """SUM over ranks of a float64 host vector, in place (sharded stepped scan)"""
        assert v.dtype == np.float64 and v.flags["C_CONTIGUOUS"]
        check(_ffi.lib().kspec_comm_allreduce_sum(self._h, dptr(v), len(v)))

    def allreduce_plan_stats(self, plan):
        """reduce the statistics plan.zerospan_batch_dev left on the device: asynchronous, on the communicator's own
        stream (the plan can start its next batch at once); ``join(plan)`` brings the result back into the plan"""
        check(_ffi.lib().kspec_comm_allreduce_plan(self._h, plan._h))

    def join(self, plan):
        """the plan's stream waits for the last allreduce_plan_stats and takes the reduced vectors (fetch returns them).
        Only while the plan still holds the batch that was reduced; otherwise KSPEC_ERR_STATE -> use fetch_reduced."""
        check(_ffi.lib().kspec_comm_join(self._h, plan._h))

    def fetch_reduced(self, n):
        """(max, min, avg) float64[n] of the last allreduce_plan_stats, whatever the plan has done since"""
        mx, mn, av = (np.empty(n, dtype=np.float64) for _ in range(3))
        check(_ffi.lib().kspec_comm_fetch_reduced(self._h, dptr(mx), dptr(mn), dptr(av), int(n)))
        return mx, mn, av

    def peer_setup(self, plan):
        """collective: fold the exchange into the plan's statistics kernel (peer-memory writes over NVLink, kspec_comm_peer_setup):
        every sharded zerospan_batch_dev of ``plan`` then leaves the statistics of the whole capture in the plan"""
        check(_ffi.lib().kspec_comm_peer_setup(self._h, plan._h))

    def peer_timed_out(self):
        t = C.c_int(0)
        check(_ffi.lib().kspec_comm_peer_status(self._h, C.byref(t)))
        return bool(t.value)

    def close(self):
        if self._h is not None and self._h.value:
            _ffi.lib().kspec_comm_finalize(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
