# parity suite, then the configs an ambiguity-path change can move: cfg2 (no ambiguity), the stress stand-in with and without
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 2 4; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('cfg$c', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"; done
for extra in "" "--no-ambiguity"; do timeout 400 python bench.py --config 5 $extra --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('cfg5 $extra', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"; done
tail -5 gpurun_out/sweep.err
