# end-to-end (host buffers through the C ABI) throughput vs pipeline chunk size
for c in ${SWEEP:-32768 65536 131072 262144 524288}; do
  export RP_CHUNK_READS=$c
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('chunk=$c', 'kernel ms=%.3f'%j['ms_per_step'], 'e2e ms=%.3f'%j['e2e']['ms_per_step'], 'e2e reads/s=%.3e'%j['e2e']['value'])
"
done
