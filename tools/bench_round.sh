# the round's measurement set: default bench line, ncu launch list of the same command, the other configs, the reference arm
set -x
V=${V:-v10}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$V.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${V}_cfg4.json 2> gpurun_out/bench_${V}_cfg4.err
python bench.py --config 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${V}_cfg1.json 2> gpurun_out/bench_${V}_cfg1.err
python bench.py --config 3 --reads 1000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${V}_cfg3.json 2> gpurun_out/bench_${V}_cfg3.err
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${V}_cfg5.json 2> gpurun_out/bench_${V}_cfg5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${V}_reference.json 2> gpurun_out/bench_${V}_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 3 -c 1 -f -o gpurun_out/prof_$V python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$V.log 2>&1
tail -2 gpurun_out/*.err
