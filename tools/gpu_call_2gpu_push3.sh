# 2 GPUs: flattened pack + two placement streams; NCCL-rank tests, then the time line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiproc.py -x -q 2>&1 | tail -3
run() { # tag envs... 
  tag=$1; shift
  env RP_XCHG_DEBUG=1 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 2 --config 5 --k5 13 --reads 200000 --steps 2 --warmup 1 $EXTRA > gpurun_out/p4_$tag.json 2> gpurun_out/p4_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/p4_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/p4_$tag.err').read()[-1500:])
PY
  grep "rp_xchg\[0\]" gpurun_out/p4_$tag.err | tail -$((J+6)) | cut -c1-90
}
EXTRA=""
J=6; run amb
J=0; run amb_one RP_XCHG_ONE_STREAM=1
J=0; run amb_r0 RP_XCHG_RESERVE_SMS=0
EXTRA="--no-ambiguity"
J=5; run noamb
