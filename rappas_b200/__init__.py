"""rappas_b200 -- B200-native (sm_100a) implementation of the RAPPAS placement hot path.

Layout:
  csrc/         CUDA kernels + the C ABI (include/rappas_b200.h) -> librappas_b200.so
  _abi/_lib     ctypes view / loader of that library
  engine        Database: load a phylo-kmer DB onto the GPU(s), place batches of reads
  synth         seeded synthetic DBs / reads of the BASELINE.json shapes
"""
from ._abi import place_cfg  # noqa: F401
from .engine import Database, device_count, kernel_launch_count, partition_of_keys  # noqa: F401
