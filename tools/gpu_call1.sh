# round 2, GPU call 1: java probe, parity of the restructured kernel, geometry sweeps against the round-1 build
mkdir -p gpurun_out
( which java javac; java -version; ls /usr/lib/jvm; nproc; free -g | head -2; nvidia-smi -L ) > gpurun_out/java_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/t_parity.log 2>&1; tail -5 gpurun_out/t_parity.log
timeout 300 python tools/sweep_geom.py --config 2 --tag NEW --envs ";RP_NO_DIRECT=1;RP_PASSES=2" >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
RAPPAS_B200_LIB=build/variants/OLD.so timeout 300 python tools/sweep_geom.py --config 2 --tag OLD >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
timeout 300 python tools/sweep_geom.py --config 4 --tag NEW >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
RAPPAS_B200_LIB=build/variants/OLD.so timeout 300 python tools/sweep_geom.py --config 4 --tag OLD >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
timeout 600 python tools/sweep_geom.py --config 3 --tag NEW --envs "RP_PASSES=1;RP_PASSES=2;RP_PASSES=3;RP_PASSES=4;RP_PASSES=2,RP_NO_DIRECT=1;RP_PASSES=2,RP_STAGE_BYTES=7424" >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
RAPPAS_B200_LIB=build/variants/OLD.so timeout 400 python tools/sweep_geom.py --config 3 --tag OLD >> gpurun_out/sweep1.jsonl 2>> gpurun_out/sweep1.err
cat gpurun_out/sweep1.jsonl; grep geometry gpurun_out/sweep1.err | sort | uniq -c | head -30
