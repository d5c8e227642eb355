for v in OLD MAIN; do
export RAPPAS_B200_LIB=build/variants/$v.so
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_$v.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:place_kernel -s 3 -c 1 -f -o gpurun_out/prof_$v python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$v.log 2>&1
tail -1 gpurun_out/ncu_$v.log
done
ls -la gpurun_out/*.ncu-rep
