# 2 GPUs: NCCL exchange tests, then the footprint diagnosis of the peer-memory form vs the exchange form
mkdir -p gpurun_out; rm -f gpurun_out/c6_*.json
nvidia-smi topo -m | head -8 > gpurun_out/c6_topo.txt
timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_parity.py -x -q -k "process or peer or multi_device" > gpurun_out/t_multi.log 2>&1; tail -8 gpurun_out/t_multi.log
run() { # tag args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 2 --config 5 --reads 200000 --steps 2 --warmup 1 "$@" > gpurun_out/c6_$tag.json 2> gpurun_out/c6_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/c6_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j['matches_oracle'], 'GB/gpu=%.2f'%(j['config']['db_bytes_per_gpu']/1e9))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/c6_$tag.err').read()[-1500:])
PY
}
run peer_k11 --peer --k5 11
run peer_k12 --peer --k5 12
run peer_k13 --peer --k5 13
run xchg_k12 --k5 12
run xchg_k13 --k5 13
run xchg_k13_noamb --k5 13 --no-ambiguity
run peer_k13_noamb --peer --k5 13 --no-ambiguity
timeout 300 python bench.py --config 5 --reads 200000 --steps 2 --warmup 1 --k5 13 > gpurun_out/c6_one_k13.json 2> gpurun_out/c6_one_k13.err; tail -c 600 gpurun_out/c6_one_k13.json
