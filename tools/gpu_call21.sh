# 1 GPU: adaptive chunk size of rp_place_batch on sliced trees -- invariance tests, then the default line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -k "pinned or split or device_pointer or large_trees or config3" -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/g21.json 2> gpurun_out/g21.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/g21.json').read().strip().split('\n')[-1])
    print('default', 'value=%.4e'%j['value'], 'e2e=%.4e'%j['e2e']['value'], 'e2e ms=%.1f'%j['e2e']['ms_per_step'], 'pageable=%.4e'%j['e2e']['pageable']['value'], 'frac=%.3f'%j['roofline']['frac'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/g21.err').read()[-800:])
PY
