# the whole GPU suite with 1, 2 and 4 consumers per team forced, then cfg3 / cfg2 timings per variant
for c in 1 2 4; do
  echo "== RP_CONSUMERS=$c"; RP_CONSUMERS=$c timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
done
for c in 1 2 4; do
  for cfg in 3 2; do
    RP_CONSUMERS=$c timeout 600 python bench.py --config $cfg --reads 1000000 --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('consumers=$c cfg$cfg', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
  done
done
