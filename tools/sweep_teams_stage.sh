# teams of C consumers x stage size on a big-tree workload (default cfg3, 1 M reads)
export RP_DEBUG_GEOM=1
for c in ${CONS:-1 2 4}; do
  for sb in ${STAGES:-0 4096 5632 12288}; do
    if [ "$sb" = 0 ]; then unset RP_STAGE_BYTES; else export RP_STAGE_BYTES=$sb; fi
    RP_CONSUMERS=$c timeout 300 python bench.py ${BENCH_ARGS:---config 3 --reads 1000000} --steps 4 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/geom.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('consumers=$c stage=$sb', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
    grep -m1 geometry gpurun_out/geom.err
  done
done
