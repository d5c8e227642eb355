mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exchange.py -x -q 2>&1 | tail -3
RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 2 --reads 50000 2>&1 | grep "rp_xchg\[0\] [a-z]" | tail -5
