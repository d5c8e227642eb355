mkdir -p gpurun_out; rm -f gpurun_out/sweep7.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exchange.py -x -q > gpurun_out/t_par7.log 2>&1; tail -5 gpurun_out/t_par7.log
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1" >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
timeout 300 python tools/sweep_geom.py --config 5 --reads 100000 --tag NEWgenome >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
RAPPAS_B200_LIB=build/variants/OLD.so timeout 300 python tools/sweep_geom.py --config 5 --reads 100000 --tag OLDgenome >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
cat gpurun_out/sweep7.jsonl
timeout 300 python bench.py --config 5 --reads 200000 --steps 2 --warmup 1 --k5 13 > gpurun_out/c7_one_k13.json 2> gpurun_out/c7_one_k13.err; python -c "
import json; j=json.loads(open('gpurun_out/c7_one_k13.json').read().strip().split('\n')[-1]); print('one_k13 value=%.3e e2e=%.3e ok=%s'%(j['value'],j['e2e']['value'],j['matches_oracle']))"
timeout 300 python bench.py --config 5 --reads 200000 --steps 2 --warmup 1 --k5 13 --no-ambiguity > gpurun_out/c7_one_k13_noamb.json 2> gpurun_out/c7_one_k13_noamb.err; python -c "
import json; j=json.loads(open('gpurun_out/c7_one_k13_noamb.json').read().strip().split('\n')[-1]); print('one_k13_noamb value=%.3e e2e=%.3e ok=%s'%(j['value'],j['e2e']['value'],j['matches_oracle']))"
