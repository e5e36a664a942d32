#!/usr/bin/env python3
"""Issue-stall histogram of a SASS region from the control words (bits 105..108 of each 128-bit instruction):
   tools/sass_stalls.py file.sass [first_line last_line]     (file = cuobjdump -sass output)
Sums the stall fields = cycles one warp needs to issue the region when no scoreboard wait binds (B300_MICROARCH.md)."""
import re
import sys
import collections

lines = open(sys.argv[1]).read().split("\n")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(lines)
ins = []
i = 0
pat = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/")
while i < len(lines) - 1:
    m = pat.match(lines[i])
    if m:
        m2 = re.match(r"^\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2 and lo <= i <= hi:
            hiw = int(m2.group(1), 16)
            stall = (hiw >> 41) & 0xf
            yld = (hiw >> 45) & 1
            wbar = (hiw >> 46) & 7
            rbar = (hiw >> 49) & 7
            wait = (hiw >> 52) & 0x3f
            ins.append((m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1], stall, wait, m.group(2)))
        i += 2
    else:
        i += 1
tot = sum(s for _, s, _, _ in ins)
print("instructions %d, sum of stall fields %d (%.2f per instruction)" % (len(ins), tot, tot / max(len(ins), 1)))
by = collections.defaultdict(lambda: [0, 0])
for op, s, w, _ in ins:
    k = op.split(".")[0]
    by[k][0] += 1
    by[k][1] += s
for k, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %-10s n=%4d stall sum=%5d avg=%.2f" % (k, n, s, s / n))
if len(sys.argv) > 4:
    for op, s, w, txt in ins:
        print("%2d %02x %s" % (s, w, txt[:90]))
