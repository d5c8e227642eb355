# the round's measurement set: default bench line, ncu launch list of the same command, cfg4 and cfg3 data points
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v7_cfg4.json 2> gpurun_out/bench_v7_cfg4.err
python bench.py --config 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v7_cfg1.json 2> gpurun_out/bench_v7_cfg1.err
python bench.py --config 3 --reads 1000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v7_cfg3.json 2> gpurun_out/bench_v7_cfg3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v7_reference.json 2> gpurun_out/bench_v7_reference.err
tail -2 gpurun_out/*.err
