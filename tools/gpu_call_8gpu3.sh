# 8 GPUs: config 5 as specified, NCCL all-to-all of the blocks against the push form, same box, both with the flattened pack
# kernel and two placement streams
mkdir -p gpurun_out; rm -f gpurun_out/c8i_*.json
run() { # tag envs... -- args
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 8 --config 5 --steps 3 --warmup 1 > gpurun_out/c8i_$tag.json 2> gpurun_out/c8i_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/c8i_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/c8i_$tag.err').read()[-2000:])
PY
  grep "rp_xchg\[0\] [a-z]" gpurun_out/c8i_$tag.err | tail -5 | cut -c1-90
}
export RP_XCHG_DEBUG=1
run nccl RP_XCHG_PUSH=0
run push RP_XCHG_PUSH=1
run push_r24 RP_XCHG_PUSH=1 RP_XCHG_RESERVE_SMS=24
