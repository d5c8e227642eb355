# final 1-GPU session of the round: the whole -m gpu suite, the default bench line, the reference arm, the other configs, the launch list
mkdir -p gpurun_out; rm -f gpurun_out/fin_*
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/fin_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/fin_bench_default.json 2> gpurun_out/fin_bench_default.err; tail -c 1800 gpurun_out/fin_bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/fin_bench_reference.json 2> gpurun_out/fin_bench_reference.err; tail -c 700 gpurun_out/fin_bench_reference.json
for c in 1; do
  timeout 600 python bench.py --config $c --no-cpu > gpurun_out/fin_bench_cfg$c.json 2> gpurun_out/fin_bench_cfg$c.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/fin_bench_cfg$c.json').read().strip().split('\n')[-1])
    print('cfg$c', 'reads/s=%.4e'%j['value'], 'ms=%.3f'%j['ms_per_step'], 'e2e=%.4e'%j['e2e']['value'], 'frac=%.3f'%j['roofline']['frac'])
except Exception as e:
    print('cfg$c FAILED', e); print(open('gpurun_out/fin_bench_cfg$c.err').read()[-800:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/fin_ncu.log 2>&1; tail -2 gpurun_out/fin_ncu.log | cut -c1-300
