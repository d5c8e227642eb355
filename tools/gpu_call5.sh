mkdir -p gpurun_out; rm -f gpurun_out/sweep5.jsonl
for v in STG3 STG4; do
RAPPAS_B200_LIB=build/variants/$v.so timeout 600 python tools/sweep_geom.py --config 3 --tag $v --envs "RP_PASSES=2;RP_PASSES=3;RP_PASSES=4;RP_PASSES=2,RP_STAGE_BYTES=3200;RP_PASSES=3,RP_STAGE_BYTES=2560" >> gpurun_out/sweep5.jsonl 2>> gpurun_out/sweep5.err
done
RAPPAS_B200_LIB=build/variants/STG3.so timeout 300 python tools/sweep_geom.py --config 2 --tag STG3 --envs ";RP_STAGE_BYTES=3200" >> gpurun_out/sweep5.jsonl 2>> gpurun_out/sweep5.err
cat gpurun_out/sweep5.jsonl; grep geometry gpurun_out/sweep5.err | sort | uniq
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
