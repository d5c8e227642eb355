# how does the kernel scale with resident producer/consumer pairs per SM?
for p in ${SWEEP:-4 6 8 10 12}; do
  export RP_PAIRS_PER_SM=$p
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('pairs=$p', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
done
