# 4 GPUs: NCCL exchange test, config 5 at k=14 through the exchange form and (diagnosis) the peer-memory form, config 3 strong scaling
mkdir -p gpurun_out; rm -f gpurun_out/c4_*.json
nvidia-smi topo -m | head -8 > gpurun_out/c4_topo.txt
timeout 600 python -m pytest tests/test_gpu_multiproc.py -x -q > gpurun_out/t_multi4.log 2>&1; tail -3 gpurun_out/t_multi4.log
run() { # tag args...
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 4 "$@" > gpurun_out/c4_$tag.json 2> gpurun_out/c4_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/c4_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'), j['config'].get('db_bytes_per_gpu'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/c4_$tag.err').read()[-1500:])
PY
}
run xchg_k14 --config 5 --k5 14 --reads 400000 --steps 2 --warmup 1
run peer_k14 --config 5 --k5 14 --reads 400000 --steps 1 --warmup 1 --peer
run xchg_k14_noamb --config 5 --k5 14 --reads 400000 --steps 2 --warmup 1 --no-ambiguity
run cfg3 --steps 5 --warmup 3 --no-cpu
