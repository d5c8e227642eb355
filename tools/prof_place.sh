set -e
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 3 -c 1 -f -o gpurun_out/prof_v8 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v8.log 2>&1
tail -2 gpurun_out/ncu_v8.log
