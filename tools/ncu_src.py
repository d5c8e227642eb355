#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass` of a .ncu-rep: per CUDA source line, stall samples
and executed warp instructions (SASS rows are attributed to the source line they follow).
usage: tools/ncu_src.py rep.ncu-rep [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = ("?", 0, ""); fname = "?"
S = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip())
        if r[2] == "": continue
    if not r[2].startswith("0x"): continue
    ws, ni = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    try: s, n = int(r[ws] or 0), int(r[ni] or 0)
    except ValueError: continue
    e = S[(cur[0], cur[1])]; e[0] += s; e[1] += n; e[2] = cur[2]
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: e[3][h[6:]] += int(r[i] or 0)
            except ValueError: pass
ts = sum(e[0] for e in S.values()); ti = sum(e[1] for e in S.values())
print("total stall samples %d, warp instructions %d" % (ts, ti))
tot = collections.Counter()
for e in S.values(): tot.update(e[3])
print("stall reasons:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, ts)) for k, v in tot.most_common(8)))
print("--- by stall samples")
for (f, l), e in sorted(S.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%-4d %-80s %s" % (100.0 * e[0] / max(1, ts), 100.0 * e[1] / max(1, ti), f, l, e[2][:80], " ".join("%s=%d" % kv for kv in e[3].most_common(2))))
print("--- by instructions")
for (f, l), e in sorted(S.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%-4d %s" % (100.0 * e[1] / max(1, ti), 100.0 * e[0] / max(1, ts), f, l, e[2][:100]))
