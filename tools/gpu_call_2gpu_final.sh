# 2 GPUs, final build: the tests that need more than one GPU
timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_parity.py -k "multiproc or multi_device or peer_memory or one_process_per_gpu" -x -q 2>&1 | tail -3
