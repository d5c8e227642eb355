#!/usr/bin/env python
"""Hot SASS of a .ncu-rep: per instruction executed count (per `unit` launches of work) and stall samples.
usage: tools/ncu_sass.py rep.ncu-rep [units=1e6] [min_per_unit=20]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows:
    if len(r) != len(hdr) or r is hdr: continue
    try: data.append((r[si], int(r[ie] or 0), int(r[ws] or 0)))
    except ValueError: pass
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("instructions %d (%.0f per unit), samples %d" % (tot, tot / units, tots))
# contiguous regions with similar execution count
i = 0
while i < len(data):
    if data[i][1] / units < thr: i += 1; continue
    j = i; n = 0; smp = 0
    while j < len(data) and data[j][1] / units >= thr:
        n += data[j][1]; smp += data[j][2]; j += 1
    print("--- region [%d,%d) %d instr, %.0f exec/unit total (%.1f%%), %.1f%% of stall samples" % (i, j, j - i, n / units, 100.0 * n / tot, 100.0 * smp / tots))
    for k in range(i, j):
        print("   %6.1f  %5d  %s" % (data[k][1] / units, data[k][2], data[k][0][:100]))
    i = j
