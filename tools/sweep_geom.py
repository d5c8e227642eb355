#!/usr/bin/env python
"""Times the device-resident placement of one workload under several geometry settings in ONE process
(the synthetic DB is generated once; the library reads RP_* when a DB is loaded).

  python tools/sweep_geom.py --config 3 --reads 1000000 --envs "RP_PASSES=1;RP_PASSES=2;RP_PASSES=3,RP_STAGE_BYTES=4096"

Prints one JSON line per setting: ms per launch (CUDA events, best and mean of --steps), the geometry the
library chose (RP_DEBUG_GEOM line) and a checksum of the rows (all settings must print the same one).
Another build of the library: RAPPAS_B200_LIB=build/variants/X.so python tools/sweep_geom.py ...
"""
import argparse
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--envs", default="")
    ap.add_argument("--postings-scale", type=float, default=1.0)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import rappas_b200 as R
    from rappas_b200 import _abi, synth
    w = synth.workload(a.config)
    db = synth.make_db(w.alphabet, w.k, w.n_nodes, w.n_keys, w.mean_postings * a.postings_scale, seed=42 + w.index,
                       key_mode=w.key_mode)
    rb = synth.make_reads(db, a.reads, w.read_len, seed=1042 + w.index, iupac_rate=w.iupac_rate, n_rate=w.n_rate)
    dev = torch.device("cuda", 0)
    cfg = _abi.place_cfg()
    K, n = cfg.keep_at_most, rb.n_reads
    d_seq = torch.from_numpy(rb.seq).to(dev)
    d_off = torch.from_numpy(rb.seq_off.view(np.int64)).to(dev)
    d_n = torch.empty(n, dtype=torch.int32, device=dev)
    d_node = torch.empty((n, K), dtype=torch.int16, device=dev)
    d_score = torch.empty((n, K), dtype=torch.float32, device=dev)
    d_lwr = torch.empty((n, K), dtype=torch.float64, device=dev)
    d_cnt = torch.empty((n, 4), dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    knobs = ("RP_PASSES", "RP_STAGE_BYTES", "RP_PAIRS_PER_SM", "RP_NO_DIRECT", "RP_GRID_SMS")
    for setting in (a.envs.split(";") if a.envs else [""]):
        for kn in knobs:
            os.environ.pop(kn, None)
        for kv in filter(None, setting.split(",")):
            k_, v_ = kv.split("=")
            os.environ[k_] = v_
        os.environ["RP_DEBUG_GEOM"] = "1"
        try:
            g = R.Database.from_synth(db, devices=(0,))
        except Exception as e:  # a geometry that does not fit
            print(json.dumps({"tag": a.tag, "env": setting, "error": str(e)}), flush=True)
            continue

        def step():
            g.place_device(cfg, d_seq.data_ptr(), d_off.data_ptr(), n, d_n.data_ptr(), d_node.data_ptr(),
                           d_score.data_ptr(), d_lwr.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr(),
                           stream=stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        crc = zlib.crc32(d_score.cpu().numpy().tobytes()) ^ zlib.crc32(d_n.cpu().numpy().tobytes()) ^ \
            zlib.crc32(d_cnt.cpu().numpy().tobytes()) ^ zlib.crc32(d_st.cpu().numpy().tobytes())
        print(json.dumps({"tag": a.tag, "config": a.config, "reads": n, "env": setting, "ms_best": min(ms),
                          "ms_mean": float(np.mean(ms)), "reads_per_s": n / (min(ms) / 1e3), "crc": "%08x" % crc}),
              flush=True)
        g.close()


if __name__ == "__main__":
    main()
