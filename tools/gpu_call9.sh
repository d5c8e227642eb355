mkdir -p gpurun_out; rm -f gpurun_out/sweep9.jsonl
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all9.log 2>&1; tail -5 gpurun_out/t_all9.log
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep9.jsonl 2>> gpurun_out/sweep9.err
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1" >> gpurun_out/sweep9.jsonl 2>> gpurun_out/sweep9.err
cat gpurun_out/sweep9.jsonl
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_reference.json
