mkdir -p gpurun_out; rm -f gpurun_out/sweep2.jsonl
for v in HASHOLD SELOLD EXPOLD NOPAIR; do
  RAPPAS_B200_LIB=build/variants/$v.so timeout 300 python tools/sweep_geom.py --config 2 --tag $v --envs ";RP_NO_DIRECT=1" >> gpurun_out/sweep2.jsonl 2>> gpurun_out/sweep2.err
done
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1;RP_PAIRS_PER_SM=10;RP_STAGE_BYTES=5248" >> gpurun_out/sweep2.jsonl 2>> gpurun_out/sweep2.err
cat gpurun_out/sweep2.jsonl
export RP_NO_DIRECT=1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_r2a.log 2>&1
tail -2 gpurun_out/ncu_r2a.log
