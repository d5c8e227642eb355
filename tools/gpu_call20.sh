# 1 GPU: e2e of the default workload against the chunk size of rp_place_batch's H2D / kernel / D2H pipeline
mkdir -p gpurun_out
for ch in 65536 262144 1048576; do
  RP_CHUNK_READS=$ch timeout 400 python bench.py --reads 4000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/g20.json 2> gpurun_out/g20.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/g20.json').read().strip().split('\n')[-1])
    print('chunk=$ch', 'value=%.4e'%j['value'], 'e2e=%.4e'%j['e2e']['value'], 'e2e ms=%.1f'%j['e2e']['ms_per_step'], 'pageable=%.4e'%j['e2e']['pageable']['value'])
except Exception as e:
    print('chunk=$ch FAILED', e); print(open('gpurun_out/g20.err').read()[-800:])
PY
done
