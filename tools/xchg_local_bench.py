#!/usr/bin/env python
"""One GPU, V virtual ranks of the exchange form (rp_xchg_create_local) against the same reads placed on the whole DB:
lets ncu's launch list time the exchange kernels (ncu must not wrap a multi-rank NCCL run).
  python tools/xchg_local_bench.py --k 13 --world 2 --reads 100000 [--no-ambiguity]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=13); ap.add_argument("--world", type=int, default=2)
ap.add_argument("--reads", type=int, default=100000); ap.add_argument("--no-ambiguity", action="store_true")
ap.add_argument("--whole", action="store_true", help="also place the same reads on the whole DB (n_parts = 1)")
a = ap.parse_args()
import rappas_b200 as R
from rappas_b200 import _abi, exchange, synth, synth_hash
hdb = synth_hash.HashDB(k=a.k, n_nodes=9999, seed=47, occupancy=0.75, mean_postings=48)
proxy = synth.SynthDB(0, a.k, 9999, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64), np.zeros(0, np.uint16), np.zeros(0, np.float32))
amb = dict(iupac_rate=0.0, n_rate=0.0) if a.no_ambiguity else dict(iupac_rate=0.005, n_rate=0.002)
batches = [synth.make_reads(proxy, a.reads, (50, 1500), seed=1047 + 7919 * p, mode="uniform", **amb) for p in range(a.world)]
cfg = _abi.place_cfg()
parts = [R.Database.from_hash_db(hdb, 0, p, a.world) for p in range(a.world)]
x = exchange.Exchange.local(parts)
for it in range(2):
    t0 = time.perf_counter(); outs = x.place(batches, cfg); dt = time.perf_counter() - t0
    print("exchange (local, %d ranks): %.1f ms wall, %s" % (a.world, dt * 1e3, x.stats()), flush=True)
x.close()
for p in parts: p.close()
if a.whole:
    g = R.Database.from_hash_db(hdb)
    for rb, oo in zip(batches, outs):
        for it in range(2):
            t0 = time.perf_counter(); out = g.place(rb, cfg); dt = time.perf_counter() - t0
        print("whole DB: %.1f ms wall per %d reads; kernel ms %.1f; rows equal: %s" % (dt * 1e3, rb.n_reads, g.last_kernel_ms(),
              bool(np.array_equal(out["n_rows"], oo["n_rows"]) and np.array_equal(out["score"], oo["score"], equal_nan=True))), flush=True)
