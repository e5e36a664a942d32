#!/usr/bin/env python3
"""Condense one kernel of an ncu report into the 'metric,unit,value' form kept under profiles/ (bench.py reads the DRAM bytes and
the pipe utilisations from the newest profiles/r*_ncu_*final*.csv):  tools/ncu_to_csv.py report.ncu-rep kernel-substring out.csv"""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, needle, out = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if needle in r[hdr.index("Kernel Name")])
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k in ("Kernel Name", "Block Size", "Grid Size"):
            w.writerow([k, "", row[hdr.index(k)]])
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
                w.writerow([h, units[i], row[i]])


if __name__ == "__main__":
    main()
