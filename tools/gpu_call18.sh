# 1 GPU: A/B of the fast form of ambiguous windows (build/variants/nofast.so = without), same box
mkdir -p gpurun_out
for v in "" build/variants/nofast.so "" build/variants/nofast.so; do
  for c in 5 4; do
    extra=""; [ $c = 5 ] && extra="--reads 300000"
    RAPPAS_B200_LIB=$v timeout 600 python bench.py --config $c $extra --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/g18.json 2> gpurun_out/g18.err
    python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/g18.json').read().strip().split('\n')[-1])
    print('lib=[$v] cfg$c', 'reads/s=%.4e'%j['value'], 'ms=%.3f'%j['ms_per_step'])
except Exception as e:
    print('cfg$c FAILED', e); print(open('gpurun_out/g18.err').read()[-800:])
PY
  done
done
