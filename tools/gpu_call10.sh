mkdir -p gpurun_out; rm -f gpurun_out/sweep10.jsonl
for v in NODUMP CK10 CK11; do
RAPPAS_B200_LIB=build/variants/$v.so timeout 300 python tools/sweep_geom.py --config 4 --tag $v >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err
RAPPAS_B200_LIB=build/variants/$v.so timeout 300 python tools/sweep_geom.py --config 2 --tag $v --envs "RP_NO_DIRECT=1" >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err
done
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err
timeout 600 python tools/sweep_geom.py --config 3 --reads 1000000 --tag NEW_1M >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err
timeout 600 python tools/sweep_geom.py --config 3 --reads 4000000 --tag NEW_4M >> gpurun_out/sweep10.jsonl 2>> gpurun_out/sweep10.err
cat gpurun_out/sweep10.jsonl
