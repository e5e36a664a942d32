#!/usr/bin/env python3
"""bench.py — headline benchmark of the kspec spectrum hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): the cfg-1 kernel configuration of BASELINE.json -- zeroSpan, fftSize 2048, hanning window,
50 % overlap (curScanNonOverlap 0.5), curScanCumuMode AVG, per-scan dB + Max/Min/Avg + one 512-bin MAX waterfall row --
on a long synthetic capture (tones + noise): 2^28 complex64 IQ samples (16384 scans, 2 GiB) PER GPU per step, so the
input is far larger than L2 (126 MB) and every step streams from HBM.  A "step" = one pass of the hot path over that
capture.  N > 1: the capture is N times longer and sharded by scan range (weak scaling); only the per-bin
Max/Min/Avg vectors are all-reduced (NCCL).

  value   IQ Msamples/s, whole job, inputs resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e     same metric through the public host-buffer call kspec_zerospan_batch: pinned host IQ -> H2D -> kernels ->
          D2H of waterfall rows + Max/Min/Avg, all inside the timed region
  roofline  HBM: algorithmic bytes per launch of the fused scan kernel / its measured duration / measured HBM peak
  cpu_baseline  the oracle port (numpy float64, the reference's algorithm) on the host cores, bounded sample

  value_f64 / roofline_f64  the same device-resident step in KSPEC_PREC_AUTO (= float64: 1e-8 dB on every bin), fewer steps
  parity_check  (N > 1, outside the timed region) a cfg-3 shaped capture (fftSize 8192, kaiser, 75 % overlap, float64)
          sharded by scan range over the ranks + NCCL MAX/MIN/SUM against ONE plan over the whole capture on rank 0:
          Max/Min bit-exact, Avg <= 1e-9 dB; the run exits non-zero if it fails

--impl reference times the reference's own CPU implementation on all host cores for the same metric/config: the unmodified
kspecanal.py functions when /root/reference is present (build container), else the oracle port (the Python reference cannot
travel to the GPU box); cpu_baseline.kind says which.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "prgs-sdr-kspecanal_b200"))
sys.path.insert(0, ROOT)

F, R_NONOVERLAP, GAIN, XRES, FS = 2048, 0.5, 19.1, 512, 2.4e6
S = F * 8
LOG2_SAMPLES = int(os.environ.get("KSPEC_BENCH_LOG2_SAMPLES", "28"))
N_SCANS = (1 << LOG2_SAMPLES) // S
BASE_SCANS = min(1024, N_SCANS)          # synthetic block that is tiled to the full capture
WORKLOAD = ("zeroSpan fftSize 2048 hanning 50%% overlap cumuAVG, complex64 ingest, %d scans x %d samples (2^%d IQ samples, "
            "%.2f GiB) per GPU per step, dB + Max/Min/Avg + 512-bin MAX waterfall row per scan" % (N_SCANS, S, LOG2_SAMPLES, N_SCANS * S * 8 / 2 ** 30))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled through NVML every ~2 ms DURING the timed region
    (an nvidia-smi subprocess is too slow for a region of a few tens of milliseconds)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.err = gpu, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[gpu])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:      # no NVML: the line says so instead of inventing numbers
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, rs, pw, time.perf_counter()))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def summary(self, t0=None, t1=None):
        """t0/t1: perf_counter bounds of the timed region (samples inside it are counted separately)"""
        self.stop_flag = True
        if self.nv is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": self.err}
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        inside = [r for r in self.rows if t0 is not None and t0 <= r[3] <= t1]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm, "power_w_max": max(r[2] for r in self.rows),
                "samples": len(self.rows), "samples_in_timed_region": len(inside), "reasons": reasons,
                "how": "NVML, every ~2 ms while the GPU runs this workload: last warm-up steps, the timed region and, because the "
                       "timed region lasts only milliseconds, the same steps repeated untimed right after it"}


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel, per launch, from the committed ncu capture of this
    very command (profiles/README.md); None when no capture is committed."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*final*.csv")))
    if not files:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for row in csv.reader(open(files[-1])):
        if row and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(row[2]) * scale.get(row[1], 1.0)
    return (tot or None), os.path.relpath(files[-1], ROOT)


def profiled_pipes():
    """what bounds the fused kernel instead of HBM (SURVEY 8d asks for it next to the HBM fraction): busy share of the
    FP32 (FMA) pipe, the shared-memory / L1 data pipe and the issue slots, from the same committed ncu capture"""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*final*.csv")))
    if not files:
        return None
    want = {"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fp32_fma_pipe_pct",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "shared_l1_data_pipe_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct"}
    out = {}
    for row in csv.reader(open(files[-1])):
        if row and row[0] in want:
            out[want[row[0]]] = round(float(row[2]), 1)
    out["source"] = os.path.relpath(files[-1], ROOT)
    return out


def bind_to_gpu_numa_node(local):
    """pin this rank (and therefore its pinned host buffers, first touch) to the NUMA node of its GPU: with 8 ranks the
    end-to-end leg is bound by host memory / PCIe root complexes, not by the GPUs"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                     # 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def base_capture():
    from kspec import synth
    return synth.tones_noise(BASE_SCANS * S, seed=1)


def algorithmic_bytes(n_scans):
    """every input sample once (complex64, 8 B), one 512-bin float32 waterfall row per scan,
    Max/Min/Avg vectors and the window table once."""
    return n_scans * S * 8 + n_scans * XRES * 4 + 3 * F * 4 + F * 4


# ---------------------------------------------------------------------------------------------------------------
# CPU: the reference algorithm (oracle port) on host cores
# ---------------------------------------------------------------------------------------------------------------
CPU_BLOCK = 64      # scans synthesised per process; the timed loop walks this block repeatedly
CPU_WHAT = {"reference": "the unmodified kspecanal.py functions (sdr_curscan, fftvals_dispproc, data_cumu x3, _data_plotcompress) loaded from /root/reference",
            "port": "numpy float64 oracle port of kspecanal.py:351-397,464-484 (no /root/reference on this host)"}


def cpu_kind():
    from oracle import ref_loader
    return "reference" if ref_loader.available() else "port"


def _cpu_worker_reference(seed, n_scans):
    """the UNMODIFIED kspecanal.py loop body (K:464-480) on its own numpy sdr_curscan: build container only"""
    import contextlib
    import io
    from kspec import synth
    from oracle import ref_loader
    ns = ref_loader.load()
    with contextlib.redirect_stdout(io.StringIO()):
        d = ref_loader.base_dict(ns, ["zeroSpan", "fftSize", F, "window", "hanning", "curScanNonOverlap", R_NONOVERLAP, "xRes", XRES])
    d["Fft.Max"] = d["Fft.Min"] = d["Fft.Avg"] = None
    x = synth.tones_noise(CPU_BLOCK * S, seed=seed).astype(np.complex128)

    class LoopSdr:                                   # read_samples(n) walks the 64-scan block again and again (K:339-346)
        pos = 0

        def read_samples(self, n):
            n = int(n)
            if self.pos + n > len(x):
                self.pos = 0
            out = x[self.pos:self.pos + n]
            self.pos += n
            return out

    d["sdr"] = LoopSdr()
    cumu, dispproc, compress, curscan = ns["data_cumu"], ns["fftvals_dispproc"], ns["_data_plotcompress"], ns["sdr_curscan"]
    t0 = time.perf_counter()
    for _ in range(n_scans):
        lin = curscan(d)
        pr = dispproc(d, lin, "LogNoGain")
        d["Fft.Max"] = cumu(d, "MAX", d["Fft.Max"], 0, len(pr), pr, 0, len(pr))
        d["Fft.Min"] = cumu(d, "MIN", d["Fft.Min"], 0, len(pr), pr, 0, len(pr))
        d["Fft.Avg"] = cumu(d, "AVG", d["Fft.Avg"], 0, len(pr), pr, 0, len(pr))
        compress(d, pr, "MAX")
    return time.perf_counter() - t0


def _cpu_worker(args):
    seed, n_scans = args
    from kspec import synth
    from oracle import kspec_oracle as O
    if cpu_kind() == "reference":
        return _cpu_worker_reference(seed, n_scans)
    win = O.window_table("hanning", F)
    x = synth.tones_noise(CPU_BLOCK * S, seed=seed).astype(np.complex128)    # the reference's dtype (K:335)
    t0 = time.perf_counter()
    state = None
    done = 0
    while done < n_scans:
        n = min(CPU_BLOCK, n_scans - done)
        lin = [O.curscan(x[k * S:(k + 1) * S], F, R_NONOVERLAP, win, "AVG") for k in range(n)]
        z = O.zerospan(lin, GAIN, XRES, "MAX", state=state)
        state = (z["max"], z["min"], z["avg"])
        done += n
    return time.perf_counter() - t0


def cpu_reference(n_procs, scans_per_proc, reps=1):
    """throughput (IQ Msamples/s) of the reference algorithm with n_procs processes each doing scans_per_proc scans"""
    import multiprocessing as mp
    best = None
    ctx = mp.get_context("fork")
    for _ in range(reps):
        if n_procs == 1:
            wall = _cpu_worker((1, scans_per_proc))          # the worker times its compute loop only
        else:
            with ctx.Pool(n_procs) as pool:
                times = pool.map(_cpu_worker, [(i + 1, scans_per_proc) for i in range(n_procs)])
            wall = max(times)
        v = n_procs * scans_per_proc * S / wall / 1e6
        best = v if best is None else max(best, v)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    scans = 640                                  # ~0.7 s of CPU work per process per step
    for _ in range(max(args.warmup, 0)):
        cpu_reference(procs, 8)
    t0 = time.perf_counter()
    vals = [cpu_reference(procs, scans) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "IQ Msamples/s via window+FFT+max/min/avg at fftSize 2048", "value": v, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d processes x %d scans of the same configuration per step" % (procs, scans)},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": procs, "kind": cpu_kind(),
                         "sample": "%d scans (%d IQ samples) per step, %s" % (procs * scans, procs * scans * S, CPU_WHAT[cpu_kind()])},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def mark(msg):
    """progress marker on stderr (KSPEC_BENCH_VERBOSE=1): where a multi-rank run is when something stalls"""
    if os.environ.get("KSPEC_BENCH_VERBOSE"):
        print("[bench rank %s %.1fs] %s" % (os.environ.get("RANK", "0"), time.perf_counter() - T_START, msg), file=sys.stderr, flush=True)


T_START = time.perf_counter()


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = bind_to_gpu_numa_node(local) if not os.environ.get("KSPEC_NO_NUMA_BIND") else None
    from kspec import _ffi
    from kspec.engine import Plan                 # product path: no oracle import here (only _cpu_worker uses it)

    win = np.hanning(F)                           # the host builds the window with numpy (K:933), as kspecanal.py does
    prec = os.environ.get("KSPEC_BENCH_PRECISION", "f32")
    plan = Plan(F, S, R_NONOVERLAP, win, "AVG", _ffi.IN_C64, precision=prec, device=local)
    info = plan.info
    base = base_capture()
    # device-resident capture: the synthetic block tiled N_SCANS/BASE_SCANS times (content does not affect timing)
    d_samples = plan.dev_alloc(N_SCANS * S * 8)
    for i in range(N_SCANS // BASE_SCANS):
        plan.dev_upload(d_samples, base, offset=i * base.nbytes)
    # pinned host copy for the end-to-end leg
    host = host_out = None
    if not os.environ.get("KSPEC_BENCH_FAST"):
        pinned = _ffi.PinnedBuffer(N_SCANS * S * 8)
        host = pinned.view(np.complex64)
        for i in range(N_SCANS // BASE_SCANS):
            host[i * len(base):(i + 1) * len(base)] = base
        # the waterfall rows come back into a pinned buffer the caller owns and reuses (64 MiB per step)
        pinned_out = _ffi.PinnedBuffer(N_SCANS * XRES * 8)
        host_out = {"hm_rows": pinned_out.view(np.float64).reshape(N_SCANS, XRES)}

    comm = None
    if world > 1:
        comm = _make_comm(world, rank, local, dist)
        plan.reserve_sms(int(os.environ.get("KSPEC_SM_RESERVE", "0")))      # measured: reserving SMs for the NCCL kernels does not pay (profiles/README.md)
    comm_sync = bool(int(os.environ.get("KSPEC_COMM_SYNC", "0")))
    mark("buffers ready")
    parity = None
    if comm is not None and not os.environ.get("KSPEC_BENCH_FAST"):
        parity = parity_check(world, rank, local, comm)
        mark("parity check done")
        dist.barrier()
    total_scans = N_SCANS * world
    base_idx = rank * N_SCANS

    def barrier():
        plan.sync()
        if dist is not None:
            dist.barrier()

    def step_dev():
        plan.zerospan_batch_dev(d_samples, N_SCANS, GAIN, XRES, "MAX", rows=None, want_hm=True,
                                scan_index_base=base_idx, n_scans_total=total_scans)
        if comm is not None:
            comm.allreduce_plan_stats(plan)          # asynchronous: overlaps the next step's kernel
            if comm_sync:
                comm.join(plan)

    def step_e2e():
        out = plan.zerospan_batch(host, N_SCANS, GAIN, XRES, "MAX", rows=None, want_hm=True,
                                  scan_index_base=base_idx, n_scans_total=total_scans, out=host_out)
        if comm is not None:
            comm.allreduce_host(out["max"], out["min"], out["avg"])
        return out

    # ---- device-resident timing -----------------------------------------------------------------------------
    mark("device-resident timing")
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_dev()
    barrier()
    l0 = plan.launch_count()
    t_region0 = time.perf_counter()
    plan.timer_start()
    for _ in range(args.steps):
        step_dev()
    if comm is not None:
        comm.join(plan)              # the last exchange is inside the timed region
    ms = plan.timer_stop()
    t_region1 = time.perf_counter()
    launches = plan.launch_count() - l0
    kt = plan.kernel_times(min(args.steps, 64))
    # keep the identical load running (untimed) so that the clock sampler sees it for long enough.  A FIXED number of steps:
    # every step is a collective at N > 1, so all ranks must issue the same count
    for _ in range(12):
        for _ in range(8):
            step_dev()
        plan.sync()
    if comm is not None:
        comm.join(plan)
    barrier()
    clocks = sampler.summary(t_region0, t_region1)
    ms_all = _max_over_ranks(ms, dist, local)
    value = world * N_SCANS * S * args.steps / (ms_all * 1e-3) / 1e6

    if os.environ.get("KSPEC_BENCH_FAST"):          # kernel tuning runs: device-resident number only
        if rank == 0:
            k_ms = float(np.mean(kt)) if kt else ms / args.steps
            print(json.dumps({"fast": True, "variant": os.environ.get("KSPEC_VARIANT", "0"), "value": value, "ms_per_step": ms_all / args.steps,
                              "kernel_ms": k_ms, "frac": algorithmic_bytes(N_SCANS) / (k_ms * 1e-3) / 1e9 / peaks()[0],
                              "ctas_per_sm": info.ctas_per_sm, "smem": info.smem_bytes, "stages": info.tma_stages, "clocks": clocks}))
        return 0
    mark("float64 leg")
    f64 = f64_leg(plan, d_samples, local, world, dist, min(args.steps, 5))
    mark("end-to-end leg")
    # ---- end to end -----------------------------------------------------------------------------------------------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    plan.sync()
    e2e_s = time.perf_counter() - t0
    e2e_s = _max_over_ranks(e2e_s, dist, local)
    e2e_value = world * N_SCANS * S * args.steps / e2e_s / 1e6
    h2d = N_SCANS * S * 8
    d2h = N_SCANS * XRES * 8 + 3 * F * 8

    # supplementary: the same end-to-end call on the wire format of the reference's device, interleaved uint8 I/Q (2 B per
    # sample: octave/load_rtlsdr.m:8-12; the conversion pyrtlsdr does on the host is fused into the first FFT stage here)
    # (at N > 1 as well: 2 B per sample is the format whose end-to-end rate can still scale when the host's PCIe / memory
    # system is the limit; every rank runs its shard, the slowest rank sets the time)
    mark("uint8 end-to-end leg")
    e2e_u8 = e2e_uint8_leg(win, S, host_out, args.steps, local, world, dist)

    mark("cpu baseline (rank 0)")
    if rank == 0:
        peak, peak_src = peaks()
        k_ms = float(np.mean(kt)) if kt else ms / args.steps
        ach = algorithmic_bytes(N_SCANS) / (k_ms * 1e-3) / 1e9
        # cpu baseline: bounded sample, all host cores
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))      # the CPU baseline may use every core again
        except Exception:
            pass
        cores = min(os.cpu_count() or 1, 64)
        cpu_v = cpu_reference(cores, 640, reps=2)
        cpu_1 = cpu_reference(1, 640)
        line = {
            "metric": "IQ Msamples/s via window+FFT+max/min/avg at fftSize 2048", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": plan.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "input %.2f GiB per step >> 126 MB L2, no flush needed" % (N_SCANS * S * 8 / 2 ** 30),
                       "frames_per_s": value * 1e6 * info.n_frames / S, "parallelism": "scan-range shards x%d" % world, "numa_node_rank0": numa,
                       "cta_threads": info.cta_threads, "ctas_per_sm": info.ctas_per_sm, "smem_bytes": info.smem_bytes},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": profiled_traffic()[0], "traffic_source": profiled_traffic()[1], "other_pipes": profiled_pipes(),
                         "peak_source": peak_src, "kernel": ("curscan_r32_kernel<C64>" if plan.precision == "f32" else "curscan_smem_kernel<double,C64,11>"), "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": algorithmic_bytes(N_SCANS)},
            "cpu_baseline": {"value": cpu_v, "unit": "Msamples/s", "cores": cores, "kind": cpu_kind(), "single_core_value": cpu_1,
                             "sample": "%d scans per process x %d processes (%.0f M IQ samples), best of 2, %s" % (640, cores, 640 * cores * S / 1e6, CPU_WHAT[cpu_kind()])},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_uint8_iq": e2e_u8,
            "value_f64": f64["value"], "roofline_f64": f64["roofline"], "f64": f64,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity is not None:
            line["parity_check"] = parity
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            sys.stdout.flush()
            os._exit(3)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    plan.dev_free(d_samples)
    plan.close()
    return 0


def parity_check(world, rank, local, comm, peer=False):
    """N > 1, outside every timed region: BASELINE cfg-3 shape (fftSize 8192, kaiser(64), 75 % overlap, float64), 96 scans
    per rank sharded by scan range.  Every rank runs its shard (host path + NCCL on host vectors, then device-resident path
    + asynchronous NCCL on the plan's vectors); rank 0 also runs ONE plan over the whole capture.  K:460-476."""
    from kspec import _ffi, synth
    from kspec.engine import Plan
    # 96 scans per rank: above the 64-scan limit of the frame-parallel small-batch form, so that the shards and the single plan
    # run the same kernel form and the comparison can be bit-exact
    F3, r3, per = 8192, 0.25, 96
    S3, n = F3 * 8, per * world
    x = synth.tones_noise(n * S3, seed=3, gate=(40000, 0.5))           # seeded: every rank builds the same capture
    win = np.kaiser(F3, 64)
    a = rank * per
    shard = x[a * S3:(a + per) * S3]
    with Plan(F3, S3, r3, win, "AVG", _ffi.IN_C64, precision="f64", device=local) as plan:
        mark("parity: capture built")
        out = plan.zerospan_batch(shard, per, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=n)
        comm.allreduce_host(out["max"], out["min"], out["avg"])
        mark("parity: host all-reduce done")
        d = plan.dev_alloc(shard.nbytes)
        plan.dev_upload(d, shard)
        plan.zerospan_batch_dev(d, per, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=n)
        comm.allreduce_plan_stats(plan)
        comm.join(plan)
        dev = plan.zerospan_fetch(rows=False, hm=False)
        mark("parity: device all-reduce done")
        pex = None
        if peer:
            comm.peer_setup(plan)                # same fftSize as the timed plan: the exchange moves to this plan
            plan.zerospan_batch_dev(d, per, GAIN, XRES, "MAX", scan_index_base=a, n_scans_total=n)
            pex = plan.zerospan_fetch(rows=False, hm=False)
            if comm.peer_timed_out():
                pex = None
            mark("parity: peer exchange done")
        plan.dev_free(d)
        if rank != 0:
            return None
        one = plan.zerospan_batch(x, n, GAIN, XRES, "MAX")
    err = {}
    ok = True
    if peer and pex is None:
        ok = False
    for name, got in (("host_allreduce", out), ("device_allreduce", dev)) + ((("peer_exchange", pex),) if pex is not None else ()):
        e = {k: float(np.max(np.abs(got[k] - one[k]))) for k in ("max", "min", "avg")}
        ok = ok and np.array_equal(got["max"], one["max"]) and np.array_equal(got["min"], one["min"]) and e["avg"] <= 1e-9
        err[name] = e
    return {"n_ranks": world, "shape": "fftSize 8192 kaiser 75 %% overlap float64, %d scans per rank, sharded by scan range" % per,
            "criterion": "Max/Min bit-exact, Avg <= 1e-9 dB against one plan over the whole capture", "max_abs_err": err,
            "max_abs_err_all": max(max(e.values()) for e in err.values()), "ok": bool(ok)}


def f64_leg(plan32, d_samples, local, world, dist, steps):
    """the same device-resident step in the default precision (KSPEC_PREC_AUTO = float64)"""
    from kspec import _ffi
    from kspec.engine import Plan
    plan32.sync()
    with Plan(F, S, R_NONOVERLAP, np.hanning(F), "AVG", _ffi.IN_C64, precision="auto", device=local) as plan:
        def step():
            plan.zerospan_batch_dev(d_samples, N_SCANS, GAIN, XRES, "MAX", rows=None, want_hm=True)
        for _ in range(3):
            step()
        plan.sync()
        if dist is not None:
            dist.barrier()
        plan.timer_start()
        for _ in range(steps):
            step()
        ms = plan.timer_stop()
        kt = plan.kernel_times(min(steps, 64))
        info = plan.info
    ms_all = _max_over_ranks(ms, dist, local)
    k_ms = float(np.mean(kt)) if kt else ms / steps
    alg = N_SCANS * S * 8 + N_SCANS * XRES * 8 + 4 * F * 8            # float64 outputs are 8 bytes each
    peak, peak_src = peaks()
    return {"value": world * N_SCANS * S * steps / (ms_all * 1e-3) / 1e6, "unit": "Msamples/s", "steps": steps, "ms_per_step": ms_all / steps,
            "dtype": "f64", "cta_threads": info.cta_threads, "ctas_per_sm": info.ctas_per_sm,
            "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "kernel": "curscan_smem_kernel<double,C64,11>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}}


def run_cfg3(args):
    """--workload cfg3: BASELINE config 3 as a STRONG-scaling run.  60 s capture at 2.4 MS/s = 2197 scans x 65536 samples
    (fftSize 8192, kaiser(64), 75 % overlap, cumulate AVG, float64 = KSPEC_PREC_AUTO) sharded by scan range over the ranks
    (K:460-476 per scan, K:471-476 combined with one NCCL MAX/MIN/SUM on 3 x 8192 float64).  A step = the whole capture.
    Prints one JSON line: device-resident and end-to-end rates, the exchange's share, and the result check against the
    N = 1 values (Max/Min bit-exact, Avg <= 1e-9) of a shorter capture."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from kspec import _ffi, synth
    from kspec.engine import Plan
    from kspec.sharding import shard_bounds
    F3, r3, n_total = 8192, 0.25, int(60 * FS) // (8192 * 8)
    S3 = F3 * 8
    a, b = shard_bounds(n_total, world)[rank]
    n = b - a
    base = synth.tones_noise(64 * S3, seed=3, gate=(40000, 0.5))        # 64-scan block tiled over the shard (timing only)
    plan = Plan(F3, S3, r3, np.kaiser(F3, 64), "AVG", _ffi.IN_C64, precision="auto", device=local)
    d = plan.dev_alloc(n * S3 * 8)
    pinned = _ffi.PinnedBuffer(n * S3 * 8)
    host = pinned.view(np.complex64)
    for i in range(0, n, 64):
        m = min(64, n - i)
        plan.dev_upload(d, base[:m * S3], offset=i * S3 * 8)
        host[i * S3:(i + m) * S3] = base[:m * S3]
    pinned_out = _ffi.PinnedBuffer(n * XRES * 8)
    host_out = {"hm_rows": pinned_out.view(np.float64).reshape(n, XRES)}
    comm = _make_comm(world, rank, local, dist) if world > 1 else None
    # KSPEC_PEER_EXCHANGE=1: the exchange is the tail of the statistics kernel (peer-memory writes over NVLink) instead of NCCL
    peer = comm is not None and os.environ.get("KSPEC_PEER_EXCHANGE", "1") == "1"
    if peer:
        comm.peer_setup(plan)

    def barrier():
        plan.sync()
        if dist is not None:
            dist.barrier()

    def step_dev(exchange=True):
        if peer:
            # sharded (n_scans_total) -> the batch ends with the peer exchange; "without exchange" = the same shard as a capture of its own
            plan.zerospan_batch_dev(d, n, GAIN, XRES, "MAX", rows=None, want_hm=True, scan_index_base=a if exchange else 0,
                                    n_scans_total=n_total if exchange else n)
            return
        plan.zerospan_batch_dev(d, n, GAIN, XRES, "MAX", rows=None, want_hm=True, scan_index_base=a, n_scans_total=n_total)
        if comm is not None and exchange:
            comm.allreduce_plan_stats(plan)
            comm.join(plan)                      # a capture is finished only when every rank holds the combined Max/Min/Avg

    def timed(fn, steps):
        for _ in range(max(args.warmup, 3)):
            fn()
        barrier()
        plan.timer_start()
        for _ in range(steps):
            fn()
        ms = plan.timer_stop()
        return _max_over_ranks(ms, dist, local) / steps

    ms_dev = timed(step_dev, args.steps)
    ms_noex = timed(lambda: step_dev(False), args.steps)

    def step_e2e():
        out = plan.zerospan_batch(host, n, GAIN, XRES, "MAX", rows=None, want_hm=True, scan_index_base=a, n_scans_total=n_total, out=host_out)
        if comm is not None:
            comm.allreduce_host(out["max"], out["min"], out["avg"])
        return out

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    plan.sync()
    e2e_s = _max_over_ranks((time.perf_counter() - t0) / args.steps, dist, local)
    parity = parity_check(world, rank, local, comm, peer=peer) if comm is not None else None
    if rank == 0:
        total = n_total * S3
        line = {"metric": "IQ Msamples/s via window+FFT+max/min/avg, BASELINE cfg 3 (fftSize 8192 kaiser 75 % overlap, 60 s capture)",
                "value": total / (ms_dev * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_dev, "ms_per_step_without_exchange": ms_noex, "exchange_share": max(0.0, 1.0 - ms_noex / ms_dev),
                "exchange": "peer-memory writes from stats_finish_kernel + peer_combine_kernel (kspec_comm_peer_setup)" if peer else ("NCCL all-reduce (MAX, MIN, SUM)" if comm is not None else "none"),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "zeroSpan fftSize 8192 kaiser(64) 75 %% overlap cumuAVG float64, %d scans x %d samples (60 s at 2.4 MS/s) "
                                       "sharded by scan range over %d GPU(s), complex64 ingest" % (n_total, S3, world), "scans_rank0": n},
                "e2e": {"value": total / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": n * S3 * 8, "d2h_bytes_per_step": n * XRES * 8 + 3 * F3 * 8},
                "frames_per_s": total / (ms_dev * 1e-3) * plan.n_frames / S3}
        if parity is not None:
            line["parity_check"] = parity
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    plan.dev_free(d)
    plan.close()
    return 0


def e2e_uint8_leg(win, S, host_out, steps, local=0, world=1, dist=None):
    """kspec_zerospan_batch on pinned interleaved uint8 I/Q of the same shape (supplementary), every rank its own shard"""
    from kspec import _ffi, synth
    from kspec.engine import Plan
    plan = Plan(F, S, R_NONOVERLAP, win, "AVG", _ffi.IN_U8_IQ, precision=os.environ.get("KSPEC_BENCH_PRECISION", "f32"), device=local)
    pinned = _ffi.PinnedBuffer(N_SCANS * S * 2)
    host = pinned.view(np.uint8)
    base = synth.to_u8_iq(synth.tones_noise(BASE_SCANS * S, seed=1).astype(np.complex128))
    for i in range(N_SCANS // BASE_SCANS):
        host[i * len(base):(i + 1) * len(base)] = base

    def step():
        return plan.zerospan_batch(host, N_SCANS, GAIN, XRES, "MAX", rows=None, want_hm=True, out=host_out)

    step()
    plan.sync()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    plan.sync()
    dt = _max_over_ranks(time.perf_counter() - t0, dist, local)
    plan.close()
    pinned.free()
    return {"value": world * N_SCANS * S * steps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": N_SCANS * S * 2,
            "d2h_bytes_per_step": N_SCANS * XRES * 8 + 3 * F * 8}


def _max_over_ranks(v, dist, local):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _make_comm(world, rank, local, dist):
    import torch
    from kspec.comm import Comm
    uid = Comm.unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda:%d" % local)
    dist.broadcast(t, 0)
    return Comm(world, rank, bytes(t.cpu().tolist()), local)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=["cfg1", "cfg3"],
                    help="cfg1: the headline benchmark (default, what the driver runs); cfg3: BASELINE config 3, strong scaling")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg3":
        return run_cfg3(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
