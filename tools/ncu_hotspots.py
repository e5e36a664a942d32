#!/usr/bin/env python3
"""Stall samples of an ncu report by region of the SASS:  tools/ncu_hotspots.py report.ncu-rep [bucket=64]
Prints, per bucket of consecutive instructions, the samples, the dominant stall reasons and the first instruction."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[2:]
    total = sum(int(r[ix["# Samples"]] or 0) for r in body)
    print("instructions %d, samples %d" % (len(body), total))
    for b0 in range(0, len(body), bucket):
        chunk = body[b0:b0 + bucket]
        n = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
        if n < total * 0.004:
            continue
        st = {h: sum(int(r[ix[h]] or 0) for r in chunk) for h in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in chunk)
        print("%5d  %5.1f%%  exec %9d  %s   | %s" % (b0, 100.0 * n / total, ex, " ".join("%s=%d" % (k[6:], v) for k, v in top if v),
                                                     chunk[0][ix["Source"]].strip()[:50]))


if __name__ == "__main__":
    main()
