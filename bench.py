#!/usr/bin/env python
"""bench.py -- the RAPPAS placement hot path on N B200s (one process per GPU).

A "step" = one pass of the hot path (k-mer extraction -> DB lookup -> scoring -> top-k/LWR) over one
batch of synthetic reads.  Default workload = BASELINE.json configs[1]: 1,000-taxon tree (1,999 nodes),
nucleotide k=10 omega=1.5 DB, 1 M synthetic 150 bp reads per GPU (DB replicated, reads sharded: weak
scaling, no collective on the data path).

  value     reads placed / s, whole job, inputs resident in HBM, CUDA-event timed (max over ranks)
  e2e       the same through rp_place_batch with pinned HOST buffers (H2D + kernel + D2H in the timed region)
  roofline  algorithmic bytes (SURVEY.md 8d formula) / kernel time  vs the measured HBM copy peak
  cpu_baseline  the CPU oracle (a C restatement of the Java algorithm, 1 thread) on a bounded sample

`--impl reference` times the reference algorithm on the host cores (oracle port, all threads; the Java
program itself cannot run here: no JRE, fastutil jar absent).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "reads_placed_per_sec"
UNIT = "reads/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config index 1..5 (default 2)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the config's count)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ambiguity", action="store_true", help="experiment: generate the reads without IUPAC / N characters")
    ap.add_argument("--partitioned", action="store_true",
                    help="hash-partition the DB over the ranks (peer memory over NVLink) instead of replicating it")
    ap.add_argument("--postings-scale", type=float, default=1.0,
                    help="experiment: multiply the workload's mean postings per key (the DB's posting blocks grow with it)")
    ap.add_argument("--replicate-table", action="store_true",
                    help="with --partitioned: partition only the posting blocks, keep the whole table on every GPU")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload_name):
    """dram bytes per launch of the placement kernel from the committed ncu --set full capture, or None."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return j.get(workload_name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def algorithmic_bytes(gdb, rb, out, k, chunk=100_000):
    """bytes(read) = len + 16*lookups + 6*H + 16*rows + 20, summed over the batch.
    lookups = plain windows + alternatives of treated ambiguous windows; H = postings gathered."""
    n = rb.n_reads
    lens = np.diff(rb.seq_off.astype(np.int64))
    lookups = 0
    postings = 0
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ex = gdb.extract(rb.slice(lo, hi))
        lookups += int(ex["nalt"][ex["kind"] != 2].sum())
        postings += int(ex["hits"][ex["hits"] > 0].sum())
    rows = int(out["n_rows"].sum())
    total = int(lens.sum()) + 16 * lookups + 6 * postings + 16 * rows + 20 * n
    return total, lookups, postings


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on this box's host cores."""
    if rank != 0:
        return
    import oracle_lib as O
    from rappas_b200 import _abi, synth
    w = synth.workload(args.config)
    cores = os.cpu_count() or 1
    db, _ = synth.build(w, reads=False)
    odb = O.OracleDB(db)
    cfg = _abi.place_cfg()
    # bounded sample per step: ~2 s of work on all cores (single-thread rate ~3-15 k reads/s)
    n_sample = args.cpu_sample or min(w.n_reads, 4000 * cores)
    rb = synth.make_reads(db, n_sample, w.read_len, seed=1042 + w.index, iupac_rate=w.iupac_rate, n_rate=w.n_rate)
    for _ in range(args.warmup):
        odb.place(rb.slice(0, max(1, n_sample // 8)), cfg, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        odb.place(rb, cfg, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "k": w.k, "n_nodes": w.n_nodes, "read_len": w.read_len,
                   "keep_at_most": 7, "keep_factor": 0.01},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d reads of the workload per step, sharded over %d pthreads; C restatement of "
                                   "PlacementProcess.processQueries (oracle/), NOT the JVM: no JRE in the image, "
                                   "the reference itself is single-threaded" % (n_sample, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rappas_b200 as R
    from rappas_b200 import _abi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rappas_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: DB replicated per GPU, reads sharded (each rank draws its own shard) -------------
    w = synth.workload(args.config)
    if args.no_ambiguity:
        w.iupac_rate = w.n_rate = 0.0
    n_reads = args.reads or w.n_reads
    if args.postings_scale != 1.0:
        w.mean_postings = w.mean_postings * args.postings_scale
        w.name += "_postings_x%g" % args.postings_scale
    db = synth.make_db(w.alphabet, w.k, w.n_nodes, w.n_keys, w.mean_postings, seed=42 + w.index, key_mode=w.key_mode)
    rb = synth.make_reads(db, n_reads, w.read_len, seed=1042 + w.index + 7919 * rank, iupac_rate=w.iupac_rate,
                          n_rate=w.n_rate)
    if args.partitioned and world > 1:
        gdb = R.Database.from_synth_partitioned_dist(db, device=local_rank, replicate_table=args.replicate_table)
    else:
        gdb = R.Database.from_synth(db, devices=(local_rank,))
    cfg = _abi.place_cfg()
    K = cfg.keep_at_most
    n = rb.n_reads

    # ---- device-resident leg ("value") -----------------------------------------------------------
    d_seq = torch.from_numpy(rb.seq).to(dev)
    d_off = torch.from_numpy(rb.seq_off.view(np.int64)).to(dev)
    d_n = torch.empty(n, dtype=torch.int32, device=dev)
    d_node = torch.empty((n, K), dtype=torch.int16, device=dev)
    d_score = torch.empty((n, K), dtype=torch.float32, device=dev)
    d_lwr = torch.empty((n, K), dtype=torch.float64, device=dev)
    d_cnt = torch.empty((n, 4), dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        gdb.place_device(cfg, d_seq.data_ptr(), d_off.data_ptr(), n, d_n.data_ptr(), d_node.data_ptr(),
                         d_score.data_ptr(), d_lwr.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr(),
                         stream=stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = R.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = R.kernel_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    ms_step = max_over_ranks(ms_total / args.steps)
    clocks = sampler.stop(t_wall0, t_wall1)
    value = world * n / (ms_step / 1e3)
    dev_out = {"n_rows": d_n.cpu().numpy(), "status": d_st.cpu().numpy(), "counts": d_cnt.cpu().numpy(),
               "score": d_score.cpu().numpy()}

    # ---- end-to-end leg through the host-buffer C ABI (pinned memory) -------------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t = torch.from_numpy(a).pin_memory()
            return t, t.numpy()
        keep = []
        t_seq, h_seq = pinned(rb.seq); t_off, h_off = pinned(rb.seq_off.view(np.int64)); keep += [t_seq, t_off]
        outs = {}
        for name, shape, dt in (("n_rows", (n,), np.int32), ("node", (n, K), np.int16), ("score", (n, K), np.float32),
                                ("lwr", (n, K), np.float64), ("counts", (n, 4), np.int32), ("status", (n,), np.int32)):
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name), pin_memory=True)
            keep.append(t)
            outs[name] = t.numpy()
        outs["node"] = outs["node"].view(np.uint16)
        hrb = synth.ReadBatch(h_seq, h_off.view(np.uint64))
        e2e_steps = max(1, min(args.steps, 10))
        for _ in range(2):
            gdb.place(hrb, cfg, out=outs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gdb.place(hrb, cfg, out=outs)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        dt = max_over_ranks(dt)
        h2d = int(rb.seq.nbytes + rb.seq_off.nbytes)
        d2h = int(sum(v.nbytes for v in outs.values()))
        e2e = {"value": world * n / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "rp_place_batch (C ABI, pinned host buffers, 2-stream 64k-read chunks: H2D / kernel / D2H overlapped)"}
        assert np.array_equal(outs["n_rows"], dev_out["n_rows"]) and np.array_equal(outs["score"], dev_out["score"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the (single) kernel: algorithmic bytes / event time vs measured HBM peak -------
    alg_bytes, lookups, postings = algorithmic_bytes(gdb, rb, dev_out, w.k)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step / 1e3) / 1e9
    tbytes, bbytes = gdb.device_bytes()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(w.name), "kernel": "rp::place_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "algorithmic bytes = len + 16*lookups + 6*postings + 16*rows + 20 per read (SURVEY 8d); "
                        "table %d MB + posting blocks %d MB vs 126 MB L2: partly cache-resident on this config"
                        % (tbytes >> 20, bbytes >> 20)}

    # ---- CPU baseline: the oracle port, 1 thread, bounded sample (rank 0, N=1 only) ------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle_lib as O
        odb = O.OracleDB(db)
        n_s = args.cpu_sample or min(n, 40_000)
        sample = rb.slice(0, n_s)
        odb.place(rb.slice(0, min(n, 500)), cfg)
        t0 = time.perf_counter()
        oo = odb.place(sample, cfg)
        dt = time.perf_counter() - t0
        # the bench's own parity spot-check on that sample (scores are expected bit-identical)
        same = bool(np.array_equal(oo["n_rows"], dev_out["n_rows"][:n_s]) and
                    np.array_equal(oo["score"].view(np.uint32), dev_out["score"][:n_s].view(np.uint32)))
        cpu = {"value": n_s / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d reads of the workload, %.1f s; C restatement of the Java loop (oracle/), 1 thread "
                         "like the reference (PlacementProcess.java:568); not the JVM (no JRE in the image)" % (n_s, dt),
               "host_cores_available": os.cpu_count(), "matches_gpu_bit_exact": same,
               "published_reference": "~417-556 reads/s, 1 desktop core, RAPPAS v1.00 (README.md:244)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "alphabet": "nucl" if w.alphabet == 0 else "amino", "k": w.k,
                   "n_nodes": w.n_nodes, "n_keys": db.n_keys, "n_postings": db.n_postings,
                   "reads_per_gpu": n, "read_len": w.read_len, "keep_at_most": K, "keep_factor": 0.01,
                   "db_layout": (("postings hash-partitioned over the GPUs, table replicated (gathers over NVLink peer memory), "
                                  "reads sharded, no collective") if args.partitioned and world > 1 and args.replicate_table else
                                 "hash-partitioned over the GPUs (peer memory over NVLink), reads sharded, no collective"
                                 if args.partitioned and world > 1 else "replicated per GPU, reads sharded, no collective"),
                   "l2_policy": "inputs larger than L2 (reads+DB+outputs %d MB per step vs 126 MB)"
                                % ((rb.seq.nbytes + tbytes + bbytes + n * (K * 14 + 24)) >> 20)},
        "kmer_lookups_per_sec": world * lookups / (ms_step / 1e3),
        "postings_per_sec": world * postings / (ms_step / 1e3),
        "hit_fraction": float(dev_out["counts"][:, 1].sum() / max(1, dev_out["counts"][:, 0].sum())),
        "placed_fraction": float((dev_out["status"] == 0).mean()),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
