"""Ad-hoc GPU timing probe (not a test, not the bench): config-2-shaped workload, kernel-only ms."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rappas_b200 as R
from rappas_b200 import _abi, synth

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
w = synth.workload(idx)
t = time.time(); db, rb = synth.build(w, n_reads=nreads); print("gen %.1fs keys=%d postings=%d" % (time.time() - t, db.n_keys, db.n_postings))
t = time.time(); g = R.Database.from_synth(db); print("load %.1fs bytes=%s" % (time.time() - t, g.device_bytes()))
cfg = _abi.place_cfg()
for i in range(3):
    t = time.time(); out = g.place(rb, cfg); dt = time.time() - t
    print("host call %.1f ms, kernel %.2f ms -> %.2f Mreads/s kernel-only" % (dt * 1e3, g.last_kernel_ms(), nreads / g.last_kernel_ms() / 1e3))
print("status", np.bincount(out["status"]), "rows mean", out["n_rows"].mean(), "matched mean", out["counts"][:, 1].mean())
