timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 5; do extra=""; [ $c = 5 ] && extra="--reads 200000"; timeout 600 python bench.py --config $c $extra --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('cfg$c', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'lookups/s=%.3e'%j['kmer_lookups_per_sec'], 'frac=%.3f'%j['roofline']['frac'])
"; done
