#!/bin/bash
# SASS evidence for the placement kernels of the in-tree library (no GPU needed): per kernel variant the count of the
# instructions that show what the hot path is built from -- bulk async copies (UBLKCP), mbarrier waits / arrivals
# (SYNCS), 256-bit probe loads, warp reductions -- and the lines themselves for the main variant.
so=${1:-rappas_b200/librappas_b200.so}
out=${2:-profiles/r02_sass_evidence.txt}
cuobjdump -sass "$so" > /tmp/all.sass
{
  echo "# cuobjdump -sass $so  ($(date -u +%F)); registers / spills from nvcc -Xptxas -v are in DESIGN.md"
  echo "# kernel                                        instrs UBLKCP SYNCS LDG.256 LDG.64(direct) REDUX ELECT ATOMG.128 STL/LDL"
  awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{n[name]++; if($0~/UBLKCP/)u[name]++; if($0~/SYNCS/)s[name]++; if($0~/LDG\.E\.ENL2\.256/)l[name]++; if($0~/LDG\.E\.64\.CONSTANT/)d[name]++; if($0~/REDUX/)r[name]++; if($0~/ELECT/)e[name]++; if($0~/ATOMG\.E\.(CAS|EXCH)\.128/)a[name]++; if($0~/STL|LDL/)sp[name]++}
       END{for(k in n) if (k ~ /rp/) printf "%-120s %6d %4d %5d %6d %6d %5d %5d %5d %5d\n", k, n[k], u[k], s[k], l[k], d[k], r[k], e[k], a[k], sp[k]}' /tmp/all.sass | sort
  echo
  echo "# place_kernel<false, kDirect> (configs 1-2): the lines"
  awk '/Function : .*place_kernelILb0ELi1E/{p=1;next} /Function :/{p=0} p' /tmp/all.sass | grep -E "UBLKCP|SYNCS|LDG\.E\.64\.CONSTANT|REDUX|ELECT|R2UR" | cut -c1-100
  echo
  echo "# place_kernel<true, kCuckoo> (big trees, k > 12): probe loads and bulk copy"
  awk '/Function : .*place_kernelILb1ELi0E/{p=1;next} /Function :/{p=0} p' /tmp/all.sass | grep -E "UBLKCP|LDG\.E\.ENL2\.256" | cut -c1-100
  echo
  echo "# synth_build_kernel: 128-bit atomics of the device-side cuckoo build"
  awk '/Function : .*synth_build_kernel/{p=1;next} /Function :/{p=0} p' /tmp/all.sass | grep -E "ATOMG" | cut -c1-100
} > "$out"
wc -l "$out"
