"""Read sharding for the one-process-per-GPU deployment.

Reads are independent units (no state survives a read: PlacementProcess.java:1067-1075), the DB is
replicated on every GPU, so rank r places the contiguous slice [n*r/W, n*(r+1)/W) of the batch -- the
same split `rp_place_batch` uses across the devices of one process -- and the result rows are
concatenated in rank order.  There is NO collective on the data path; `gather_results` is only the
optional convenience of bringing every rank's rows to rank 0 (e.g. to write one .jplace).
"""
from __future__ import annotations

import numpy as np

RESULT_KEYS = ("n_rows", "node", "score", "lwr", "counts", "status")


def shard_bounds(n_reads: int, world: int):
    """[(lo, hi)] per rank; identical to the device split inside rp_place_batch (rp_place.cu place_host)."""
    return [(n_reads * r // world, n_reads * (r + 1) // world) for r in range(world)]


def place_local_shard(place_fn, reads, cfg, rank: int, world: int):
    """Place this rank's slice. place_fn(reads_slice, cfg) -> dict of RESULT_KEYS arrays."""
    lo, hi = shard_bounds(reads.n_reads, world)[rank]
    return place_fn(reads.slice(lo, hi), cfg), (lo, hi)


def merge_results(parts):
    """Concatenate per-rank result dicts (rank order = read order)."""
    out = {}
    for k in RESULT_KEYS:
        vals = [p[k] for p in parts if p.get(k) is not None]
        out[k] = np.concatenate(vals, axis=0) if vals else None
    return out


def gather_results(local, dst: int = 0, group=None):
    """torch.distributed gather of the per-rank rows to `dst` (any backend; tests use gloo on CPU,
    bench.py/NCCL never needs it because every rank reports only its timing)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object({k: local.get(k) for k in RESULT_KEYS}, bucket, dst=dst, group=group)
    return merge_results(bucket) if rank == dst else None
