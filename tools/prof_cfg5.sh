set -e
mkdir -p gpurun_out
python bench.py --config 5 --reads 200000 --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/plain5.log 2>&1 && tail -c 400 gpurun_out/plain5.log
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 4 -c 2 -f -o gpurun_out/prof_cfg5 python bench.py --config 5 --reads 200000 --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_cfg5.log 2>&1
tail -2 gpurun_out/ncu_cfg5.log
