export RP_DEBUG_GEOM=1
for c in 4 1 2 3; do
extra=""; [ $c = 3 ] && extra="--reads 1000000"
python bench.py --config $c $extra --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/geom.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('cfg$c', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"; grep -m1 geometry gpurun_out/geom.err
done
unset RP_DEBUG_GEOM
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
