# 1 GPU: the flattened pack kernel -- exchange tests, the two new parity tests, and the push time line through the buffers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exchange.py "tests/test_gpu_parity.py::test_the_full_config3_database" "tests/test_gpu_parity.py::test_k15_long_reads_of_the_config5_shape" -x -q 2>&1 | tail -5
echo "== amb through the buffers"; RP_XCHG_COPY_LOCAL=1 RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -7
echo "== 2 virtual ranks"; RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 2 --reads 50000 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -7
