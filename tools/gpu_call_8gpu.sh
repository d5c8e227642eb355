# 8 GPUs: config 5 as specified (k=15, 805 M keys, ~280 GB over the 8 GPUs) through the exchange form; config 3 strong scaling; config 2 weak e2e
mkdir -p gpurun_out; rm -f gpurun_out/c8g_*.json
nvidia-smi topo -m | head -12 > gpurun_out/c8g_topo.txt
run() { # tag args...
  tag=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 8 "$@" > gpurun_out/c8g_$tag.json 2> gpurun_out/c8g_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/c8g_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'), j['config'].get('db_bytes_all_gpus'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/c8g_$tag.err').read()[-2000:])
PY
}
export RP_XCHG_DEBUG=1
run cfg5_xchg --config 5 --steps 2 --warmup 1
grep "rp_xchg\[0\]" gpurun_out/c8g_cfg5_xchg.err | tail -7 | cut -c1-80
run cfg3 --steps 5 --warmup 3 --no-cpu
