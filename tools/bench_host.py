#!/usr/bin/env python
"""Throughput of the host rows either side of the kernel, on CPU only:
native FASTA ingest + duplicate structure (rp_reads_load_fasta) and the .jplace writer (rp_jplace_write).
usage: python tools/bench_host.py [n_reads] [read_len]"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rappas_b200 import _abi  # noqa: E402
from rappas_b200._lib import check, load  # noqa: E402
from rappas_b200.ingest import QueryFile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rng = np.random.default_rng(1)
n_distinct = int(n * 0.9)  # 10 % exact duplicates
seqs = rng.integers(0, 4, (n_distinct, L), dtype=np.uint8)
pick = np.concatenate([np.arange(n_distinct), rng.integers(0, n_distinct, n - n_distinct)])
rng.shuffle(pick)
letters = np.frombuffer(b"ACGT", np.uint8)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "q.fa")
t0 = time.perf_counter()
with open(fa, "wb") as f:
    body = letters[seqs]
    for lo in range(0, n, 100000):
        hi = min(n, lo + 100000)
        f.write(b"".join(b">read_%d some description\n%s\n" % (i, body[pick[i]].tobytes()) for i in range(lo, hi)))
size = os.path.getsize(fa)
print("wrote %d records, %.1f MB in %.1f s" % (n, size / 1e6, time.perf_counter() - t0))

fn = load()
best = 1e9
for _ in range(3):
    h = C.c_void_p()
    t0 = time.perf_counter()
    check(fn["reads_load_fasta"](fa.encode(), C.byref(h)))
    dt = time.perf_counter() - t0
    best = min(best, dt)
    fn["reads_free"](h)
print("rp_reads_load_fasta: %.3f s  -> %.1f M records/s, %.0f MB/s" % (best, n / best / 1e6, size / best / 1e6))

q = QueryFile.from_file(fa)
K = 7
nu = q.n_unique
res = dict(n_rows=np.full(nu, 3, np.int32), node=rng.integers(0, 1999, (nu, K)).astype(np.uint16),
           score=(-rng.random((nu, K)) * 600).astype(np.float32), lwr=rng.random((nu, K)),
           status=np.zeros(nu, np.int32))
edge = np.arange(1999, dtype=np.int32)
bl = rng.random(1999).astype(np.float32)
out = os.path.join(tmp, "o.jplace")
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    m = q.write_jplace(out, res, K, edge, bl, tree_newick="(a,b);", invocation="bench")
    best = min(best, time.perf_counter() - t0)
print("rp_jplace_write: %d placements (%d records), %.1f MB in %.3f s -> %.1f M records/s, %.0f MB/s" %
      (m, n, os.path.getsize(out) / 1e6, best, n / best / 1e6, os.path.getsize(out) / best / 1e6))
