mkdir -p gpurun_out; rm -f gpurun_out/sweep4.jsonl
timeout 900 python -m pytest tests/test_gpu_exchange.py -x -q > gpurun_out/t_xchg.log 2>&1; tail -30 gpurun_out/t_xchg.log
timeout 600 python tools/sweep_geom.py --config 3 --tag NEW --envs "RP_PASSES=1;RP_PASSES=2" >> gpurun_out/sweep4.jsonl 2>> gpurun_out/sweep4.err
cat gpurun_out/sweep4.jsonl; grep geometry gpurun_out/sweep4.err | sort | uniq
export RP_PASSES=2
python bench.py --config 3 --reads 1000000 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2b_cfg3 python bench.py --config 3 --reads 1000000 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_r2b.log 2>&1
tail -2 gpurun_out/ncu_r2b.log
