# producer's poll interval while it waits for a released stage (0 = spin)
for ns in ${POLLS:-0 100 400 1500}; do
  python - <<PY
from rappas_b200 import build as b
b.build(force=True, extra=["-DRP_POLL_NS=$ns"])
PY
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('poll_ns=$ns', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
done
python -m rappas_b200.build --force > /dev/null
