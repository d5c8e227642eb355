# 8 GPUs: config 5 as specified (k=15, 805 M keys, ~280 GB over the 8 GPUs) through the exchange form in push mode
mkdir -p gpurun_out; rm -f gpurun_out/c8h_*.json
run() { # tag args...
  tag=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 8 "$@" > gpurun_out/c8h_$tag.json 2> gpurun_out/c8h_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/c8h_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'), j['config'].get('db_bytes_all_gpus'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/c8h_$tag.err').read()[-2000:])
PY
}
export RP_XCHG_DEBUG=1
run cfg5_push --config 5 --steps 3 --warmup 1
grep "rp_xchg\[0\]" gpurun_out/c8h_cfg5_push.err | tail -14 | cut -c1-90
run cfg5_push_noamb --config 5 --steps 3 --warmup 1 --no-ambiguity
grep "rp_xchg\[0\]" gpurun_out/c8h_cfg5_push_noamb.err | tail -6 | cut -c1-90
