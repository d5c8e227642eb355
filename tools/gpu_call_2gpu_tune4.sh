mkdir -p gpurun_out
run() { # tag envs... -- args
  tag=$1; shift
  env RP_XCHG_DEBUG=1 RP_XCHG_PROBES=16000000 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29$((RANDOM % 800 + 100)) \
    bench.py --gpus 2 --config 5 --k5 13 --reads 200000 --steps 2 --warmup 1 $EXTRA > gpurun_out/t4_$tag.json 2> gpurun_out/t4_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/t4_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'reads/s=%.3e'%j['value'], 'ms=%.1f'%j['ms_per_step'], 'e2e=%.3e'%j['e2e']['value'], 'ok=',j.get('matches_oracle'))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/t4_$tag.err').read()[-1500:])
PY
  grep "rp_xchg\[0\]" gpurun_out/t4_$tag.err | tail -2 | cut -c1-70
}
EXTRA=""
run ch8_r16 NCCL_MAX_NCHANNELS=8 RP_XCHG_RESERVE_SMS=16
run ch4_r8 NCCL_MAX_NCHANNELS=4 RP_XCHG_RESERVE_SMS=8
run ce_r4 NCCL_P2P_USE_CUDA_MEMCPY=1 RP_XCHG_RESERVE_SMS=4
run ch16_r24 NCCL_MAX_NCHANNELS=16 RP_XCHG_RESERVE_SMS=24
EXTRA="--no-ambiguity"
run noamb_ch8_r16 NCCL_MAX_NCHANNELS=8 RP_XCHG_RESERVE_SMS=16
