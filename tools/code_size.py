#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source function (nvdisasm -g line info): what fills the
instruction cache.  usage: tools/code_size.py [lib.so] [kernel-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else "rappas_b200/librappas_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "place_kernelILb0ELi0E"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.startswith("rp_place.")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
src = open("rappas_b200/csrc/rp_place.cu").read().split("\n")
starts = []  # (line, name) of top-level function definitions
for i, l in enumerate(src, 1):
    m = re.match(r"(?:static |__device__ |__forceinline__ |__global__ |inline |int |void |bool )*.*?\b([A-Za-z_]\w+)\s*\(", l)
    if m and not l.startswith((" ", "/", "#", "}")) and ("__device__" in l or "__global__" in l or l.startswith(("static", "int ", "place_kernel"))):
        starts.append((i, m.group(1)))
def fn_of(line):
    name = "?"
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
inside = False; cur = None; cnt = collections.Counter(); tot = 0
for l in dis:
    if l.startswith(".text."): inside = pat in l; continue
    if not inside: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,6}\*/", l) and cur:
        cnt[(cur[0], fn_of(cur[1]) if cur[0] == "rp_place.cu" else "")] += 1; tot += 1
print("kernel *%s*: %d SASS instructions = %.1f KB" % (pat, tot, tot * 16 / 1024))
for k, v in cnt.most_common(30): print("%6d  %5.1f KB  %s %s" % (v, v * 16 / 1024, k[0], k[1]))
