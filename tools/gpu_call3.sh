mkdir -p gpurun_out; rm -f gpurun_out/sweep3.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/t_parity.log 2>&1; tail -3 gpurun_out/t_parity.log
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1;RP_STAGE_BYTES=5248" >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
RAPPAS_B200_LIB=build/variants/NOPAIR.so timeout 300 python tools/sweep_geom.py --config 2 --tag NOPAIR --envs ";RP_NO_DIRECT=1" >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
timeout 300 python tools/sweep_geom.py --config 1 --reads 10000 --tag NEW >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
RAPPAS_B200_LIB=build/variants/OLD.so timeout 300 python tools/sweep_geom.py --config 1 --reads 10000 --tag OLD >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
timeout 600 python tools/sweep_geom.py --config 3 --tag NEW --envs "RP_PASSES=1;RP_PASSES=2;RP_PASSES=3;RP_PASSES=2,RP_STAGE_BYTES=3840;RP_PASSES=2,RP_STAGE_BYTES=5632;RP_PASSES=2,RP_PAIRS_PER_SM=6" >> gpurun_out/sweep3.jsonl 2>> gpurun_out/sweep3.err
cat gpurun_out/sweep3.jsonl
