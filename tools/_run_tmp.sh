i=0
for extra in "" "--replicate-table"; do
i=$((i+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$i bench.py --gpus 2 --partitioned $extra --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v10_n2_partitioned_layout$i.json 2>gpurun_out/part$i.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2958$i bench.py --gpus 2 --config 5 --reads 200000 --partitioned $extra --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_v10_cfg5_n2_partitioned_layout$i.json 2>gpurun_out/part5$i.err
done
grep -h -o '"value": [0-9.e+]*, "unit": "reads/s", "n_gpus": 2[^}]*"ms_per_step": [0-9.]*' gpurun_out/bench_v10_*n2_partitioned_layout*.json
tail -3 gpurun_out/part*.err | cut -c1-200
