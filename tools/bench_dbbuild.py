#!/usr/bin/env python
"""GPU phylo-k-mer generation against the recursive CPU oracle: same synthetic posteriors,
the oracle timed on a slice of the nodes (it is single-threaded, like the reference), the GPU on all of them.
usage: python tools/bench_dbbuild.py [k] [n_nodes] [n_sites] [oracle_nodes]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dbbuild_lib as D  # noqa: E402
import oracle_lib as O  # noqa: E402
from rappas_b200 import dbbuild  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n_sites = int(sys.argv[3]) if len(sys.argv) > 3 else 500
o_nodes = int(sys.argv[4]) if len(sys.argv) > 4 else 4
pp, states, oid, _, _ = D.make_inputs(0, k, n_nodes, n_sites, seed=9, peak=0.95)
thr = float(O.threshold(1.5, 0, k)[1])
dbbuild.build_db(0, k, pp[:2], states[:2], oid[:2], thr)  # warm-up (context, module load)
t0 = time.perf_counter()
g = dbbuild.build_db(0, k, pp, states, oid, thr)
wall = time.perf_counter() - t0
t0 = time.perf_counter()
o = D.oracle_build(0, k, pp[:o_nodes], states[:o_nodes], oid[:o_nodes], thr)
cpu = time.perf_counter() - t0
gs = dbbuild.build_db(0, k, pp[:o_nodes], states[:o_nodes], oid[:o_nodes], thr)
D.assert_csr_equal({f: gs[f] for f in ("keys", "offsets", "post_node", "post_score")},
                   {f: o[f] for f in ("keys", "offsets", "post_node", "post_score")})
tasks = n_nodes * (n_sites - k + 2)
print(json.dumps({
    "workload": "dbbuild nucl k=%d, %d nodes x %d sites (%d explorers)" % (k, n_nodes, n_sites, tasks),
    "tuples": g["n_tuples"], "keys": int(g["keys"].size), "postings": int(g["post_node"].size),
    "gpu_device_ms": g["kernel_ms"], "gpu_wall_s": wall, "gpu_tuples_per_s": g["n_tuples"] / (g["kernel_ms"] / 1e3),
    "gpu_explorers_per_s": tasks / (g["kernel_ms"] / 1e3),
    "cpu_oracle": {"kind": "port", "cores": 1, "sample": "%d of the %d nodes" % (o_nodes, n_nodes), "seconds": cpu,
                   "tuples_per_s": o["n_tuples"] / cpu, "explorers_per_s": o_nodes * (n_sites - k + 2) / cpu},
    "oracle_check": "the sample's CSR equals the GPU's bit for bit"}))
