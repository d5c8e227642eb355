mkdir -p gpurun_out; rm -f gpurun_out/sweep11.jsonl
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all11.log 2>&1; tail -4 gpurun_out/t_all11.log
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep11.jsonl 2>> gpurun_out/sweep11.err
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1" >> gpurun_out/sweep11.jsonl 2>> gpurun_out/sweep11.err
timeout 600 python tools/sweep_geom.py --config 3 --reads 1000000 --tag NEW --envs ";RP_AMB_BATCH=1" >> gpurun_out/sweep11.jsonl 2>> gpurun_out/sweep11.err
timeout 300 python tools/sweep_geom.py --config 5 --reads 100000 --tag NEWgenome --envs ";RP_AMB_BATCH=0">> gpurun_out/sweep11.jsonl 2>> gpurun_out/sweep11.err
cat gpurun_out/sweep11.jsonl
for a in "" "--no-ambiguity"; do
timeout 300 python bench.py --config 5 --reads 200000 --steps 2 --warmup 1 --k5 13 $a > gpurun_out/c11_one_k13$a.json 2> gpurun_out/c11_one_k13$a.err; python -c "
import json; j=json.loads(open('gpurun_out/c11_one_k13$a.json').read().strip().split('\n')[-1]); print('one_k13 $a value=%.3e e2e=%.3e ok=%s'%(j['value'],j['e2e']['value'],j['matches_oracle']))" || tail -3 gpurun_out/c11_one_k13$a.err
done
