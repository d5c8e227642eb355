# 1 GPU: two placement streams in the push pipeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exchange.py -x -q 2>&1 | tail -3
for one in 0 1; do
echo "== amb in place, one_stream=$one"; RP_XCHG_ONE_STREAM=$one RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -7
echo "== amb through the buffers, one_stream=$one"; RP_XCHG_ONE_STREAM=$one RP_XCHG_COPY_LOCAL=1 RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -7
done
echo "== noamb through the buffers"; RP_XCHG_COPY_LOCAL=1 RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 --no-ambiguity 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -6
