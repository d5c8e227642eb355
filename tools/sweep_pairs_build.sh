# rebuild with different kMaxPairsPerCta (register budget changes with it) and sweep the stage size
for p in ${PAIRS:-8 9 10}; do
  sed -i "s/constexpr int kMaxPairsPerCta = [0-9]*;/constexpr int kMaxPairsPerCta = $p;/" rappas_b200/csrc/rp_place.cu
  python -m rappas_b200.build --force --ptxas 2>&1 | grep -A3 "place_kernel" | grep -E "Used" | sed "s/^/pairs=$p /"
  for sb in ${STAGES:-5120 6144 7168 8192}; do
    export RP_STAGE_BYTES=$sb
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('pairs=$p stage=$sb', 'ms=%.3f'%j['ms_per_step'], 'reads/s=%.3e'%j['value'], 'frac=%.3f'%j['roofline']['frac'])
"
  done
done
