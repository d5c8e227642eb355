# time builds of the library that differ by a -D switch (built beforehand into build/variants/*.so)
for v in ${VARIANTS:-OLD A B C D E}; do
  export RAPPAS_B200_LIB=build/variants/$v.so
  for c in ${CONFIGS:-2 4}; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('$v cfg$c', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"; done
done
tail -3 gpurun_out/sweep.err
