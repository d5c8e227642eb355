# 1 GPU: fast form of ambiguous windows -- parity suites, then the configs it could move
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exchange.py -x -q 2>&1 | tail -3
for c in 5 2 4 3; do
  timeout 600 python bench.py --config $c --no-cpu --steps 5 --warmup 3 > gpurun_out/g17_cfg$c.json 2> gpurun_out/g17_cfg$c.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/g17_cfg$c.json').read().strip().split('\n')[-1])
    print('cfg$c', 'reads/s=%.4e'%j['value'], 'ms=%.3f'%j['ms_per_step'], 'e2e=%.4e'%j['e2e']['value'], 'frac=%.3f'%j['roofline']['frac'])
except Exception as e:
    print('cfg$c FAILED', e); print(open('gpurun_out/g17_cfg$c.err').read()[-800:])
PY
done
