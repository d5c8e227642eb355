# sweep the per-warp posting stage size (RP_STAGE_BYTES) on the default bench workload
for sb in ${SWEEP:-2048 3072 4096 0 7168}; do
  if [ "$sb" = 0 ]; then unset RP_STAGE_BYTES; else export RP_STAGE_BYTES=$sb; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('stage=$sb', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
done
