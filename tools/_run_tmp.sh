export RP_DEBUG_GEOM=1
run() { # consumers pairs
  RP_CONSUMERS=$1 RP_PAIRS_PER_SM=$2 timeout 300 python bench.py --config 2 --steps 4 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/geom.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('consumers=$1 teams<=$2', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'])
"; grep -m1 geometry gpurun_out/geom.err | cut -c20-130; }
run 1 6; run 2 6; run 1 4; run 2 4; run 4 4; run 1 3; run 4 3
