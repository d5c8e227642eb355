set -e
python tools/bench_dbbuild.py 10 60 300 2 > gpurun_out/dbbuild_plain.json 2> gpurun_out/dbbuild_plain.err
ncu --set full --clock-control none --import-source on -k regex:explore_kernel -s 2 -c 2 -f -o gpurun_out/prof_dbbuild python tools/bench_dbbuild.py 10 60 300 2 > gpurun_out/ncu_dbbuild.log 2>&1
tail -2 gpurun_out/ncu_dbbuild.log; cat gpurun_out/dbbuild_plain.json | cut -c1-400
