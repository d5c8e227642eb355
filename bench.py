#!/usr/bin/env python
"""bench.py -- the RAPPAS placement hot path on N B200s (one process per GPU).

A "step" = one pass of the hot path (k-mer extraction -> DB lookup -> scoring -> top-k/LWR) over one batch of
synthetic reads.  Default workload = BASELINE.json configs[2], the largest single-GPU configuration and the one
SURVEY.md section 10 names for the HBM-roofline claim: 5,000-taxon tree (9,999 nodes), nucleotide k=12 DB
(12.6 M keys, 604 M postings, 4.2 GB), 10 M synthetic 150 bp reads SHARDED across the GPUs (DB replicated,
strong scaling: 10 M / N reads per GPU, no collective on the data path).

  value     reads placed / s, whole job, inputs resident in HBM, CUDA-event timed (max over ranks)
  e2e       the same through rp_place_batch with pinned HOST buffers (H2D + kernel + D2H in the timed region);
            e2e.pageable = the same call from ordinary (unpinned) host memory, what a JVM byte[] / heap buffer is
  roofline  algorithmic bytes (SURVEY.md 8d formula) / kernel time  vs the measured HBM copy peak
  cpu_baseline  the CPU oracle (a C restatement of the Java algorithm, 1 thread) on a bounded sample

--config 2 / 1 / 4: the other single-GPU configs (weak scaling under --gpus N: the config's reads per GPU).
--config 5: the stress shape -- 1 M reads of 50-1500 bp with IUPAC / N characters against the hash-defined k=15 DB
  (805 M keys, 38.7 G postings, ~280 GB: larger than one GPU's HBM), every rank generating ITS partition on the
  device, placed through the exchange form (NCCL all-to-all of k-mer probes, posting lists back; --peer = the
  peer-memory form instead).  At --gpus 1 a stand-in of the same shape that fits one GPU (k=13, same N, list
  lengths, occupancy and therefore hit rate).

`--impl reference` times the reference algorithm on the host cores (oracle port; the Java program itself cannot
run here or on the GPU box: no JRE on either, fastutil jar absent from the reference tree).
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
STRONG = (3, 5)   # configs whose read count is the whole job's (BASELINE.json: "sharded across 1/2/4/8 B200")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config index 1..5 (default 3)")
    ap.add_argument("--reads", type=int, default=0, help="reads of the whole job (strong configs) / per GPU (others)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ambiguity", action="store_true", help="experiment: generate the reads without IUPAC / N characters")
    ap.add_argument("--partitioned", action="store_true",
                    help="configs 1-4: hash-partition the host-built DB over the ranks (peer memory over NVLink)")
    ap.add_argument("--peer", action="store_true", help="config 5: peer-memory form instead of the exchange form")
    ap.add_argument("--k5", type=int, default=0, help="config 5: k of the hash-defined DB (default 15 at N > 1, 13 at N = 1)")
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


def ncu_traffic(workload_name, n_reads):
    """dram bytes per launch of the placement kernel from the committed ncu --set full capture (taken at
    `reads_per_launch` reads; scaled to this launch), or None."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(workload_name)
        if not j:
            return None
        return int(j["dram_bytes_per_launch"] * (n_reads / j.get("reads_per_launch", n_reads)))
    except Exception:
        return None


def pin_to_gpu_numa_node(local_rank):
    """CPU affinity = the cores of the NUMA node the GPU hangs off, so that the pinned staging buffers are
    first-touched there and the copy threads run there.  Best effort; returns what it did."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"numa_node": None}
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": len(ids)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": str(e)[:80]}


def config_dict(w, db_keys, db_postings, reads_per_gpu, world, layout, extra=None):
    d = {"workload": w.name, "alphabet": "nucl" if w.alphabet == 0 else "amino", "k": w.k, "n_nodes": w.n_nodes,
         "n_keys": int(db_keys), "n_postings": int(db_postings), "reads_per_gpu": int(reads_per_gpu),
         "reads_total": int(reads_per_gpu * world), "read_len": w.read_len, "keep_at_most": 7, "keep_factor": 0.01,
         "db_layout": layout}
    if extra:
        d.update(extra)
    return d


def job_reads(args, w, world):
    """(reads of this rank, scaling) -- strong configs split the job's reads, the others repeat them per GPU."""
    total = args.reads or w.n_reads
    if w.index in STRONG:
        return max(1, total // world), "strong"
    return total, "weak"


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on this box's host cores (all of them, reads sharded
    over pthreads -- the reference itself is single-threaded, PlacementProcess.java:568, which the 1-thread
    figure beside it shows)."""
    if rank != 0:
        return
    import oracle_lib as O
    from rappas_b200 import _abi, synth
    w = synth.workload(args.config)
    cores = os.cpu_count() or 1
    n_job, scaling = job_reads(args, w, args.gpus)
    if args.config == 5:
        from rappas_b200 import synth_hash
        k5 = args.k5 or (15 if args.gpus > 1 else 13)
        hdb = synth_hash.HashDB(k=k5, n_nodes=w.n_nodes, seed=42 + w.index, occupancy=0.75, mean_postings=w.mean_postings)
        proxy = synth.SynthDB(0, k5, w.n_nodes, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                              np.zeros(0, np.uint16), np.zeros(0, np.float32))
        n_sample = args.cpu_sample or 250 * cores
        rb = synth.make_reads(proxy, n_sample, w.read_len, seed=1042 + w.index, mode="uniform", iupac_rate=w.iupac_rate, n_rate=w.n_rate)
        db = hdb.sub_db(synth_hash.probed_codes(rb, k5))
        n_keys, n_post = int(hdb.expected_keys()), int(hdb.expected_keys() * w.mean_postings)
        w.k = k5
    else:
        db, _ = synth.build(w, reads=False)
        n_sample = args.cpu_sample or min(n_job, 4000 * cores)
        rb = synth.make_reads(db, n_sample, w.read_len, seed=1042 + w.index, iupac_rate=w.iupac_rate, n_rate=w.n_rate)
        n_keys, n_post = db.n_keys, db.n_postings
    odb = O.OracleDB(db)
    cfg = _abi.place_cfg()
    for _ in range(args.warmup):
        odb.place(rb.slice(0, max(1, n_sample // 8)), cfg, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        odb.place(rb, cfg, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample / dt
    n1 = max(1, n_sample // (4 * cores))
    t0 = time.perf_counter()
    odb.place(rb.slice(0, n1), cfg, threads=1)
    one = n1 / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, n_keys, n_post, n_job, args.gpus,
                              "host memory (CPU arm): one hash map, reads sharded over the host threads"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample_reads": n_sample,
                         "one_thread_value": one,
                         "sample": "%d reads of the workload per step, sharded over %d pthreads; C restatement of "
                                   "PlacementProcess.processQueries (oracle/), NOT the JVM: no JRE in the image or on the "
                                   "GPU box (probed: `java` not found), fastutil-8.2.2.jar absent from the reference tree; "
                                   "the reference itself is single-threaded (one_thread_value)" % (n_sample, cores)},
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
    affinity = pin_to_gpu_numa_node(local_rank)
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

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = dict(args=args, rank=rank, local_rank=local_rank, world=world, dev=dev, barrier=barrier,
               max_over_ranks=max_over_ranks, sum_over_ranks=sum_over_ranks, affinity=affinity)
    if args.config == 5:
        run_config5(ctx)
    else:
        run_replicated(ctx)
    if world > 1:
        dist.destroy_process_group()


def device_leg(ctx, gdb, rb, cfg):
    """value leg: reads resident in HBM, rp_place_batch_device on the current stream, CUDA events around `steps`
    launches (max over ranks), clocks sampled during the timed region.  -> (ms per step, outputs, launches, clocks)"""
    import torch
    import rappas_b200 as R
    args, local_rank, dev = ctx["args"], ctx["local_rank"], ctx["dev"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]
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
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_out = {"n_rows": d_n.cpu().numpy(), "status": d_st.cpu().numpy(), "counts": d_cnt.cpu().numpy(),
               "score": d_score.cpu().numpy()}
    return ms_step, dev_out, launches, clocks


# ------------------------------------------------------------------------------------------- configs 1-4
def run_replicated(ctx):
    import torch
    import torch.distributed as dist
    import rappas_b200 as R
    from rappas_b200 import _abi, synth
    args, rank, local_rank, world, dev = ctx["args"], ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    barrier, max_over_ranks = ctx["barrier"], ctx["max_over_ranks"]

    # ---- workload: DB replicated per GPU, reads sharded (each rank draws its own shard) -------------
    w = synth.workload(args.config)
    if args.no_ambiguity:
        w.iupac_rate = w.n_rate = 0.0
    n_reads, scaling = job_reads(args, w, world)
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

    ms_step, dev_out, launches, clocks = device_leg(ctx, gdb, rb, cfg)
    value = world * n / (ms_step / 1e3)

    # ---- end-to-end leg through the host-buffer C ABI: pinned (rp_host_alloc-style) and pageable callers ----
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t = torch.from_numpy(a).pin_memory()
            return t, t.numpy()
        keep = []
        t_seq, h_seq = pinned(rb.seq); t_off, h_off = pinned(rb.seq_off.view(np.int64)); keep += [t_seq, t_off]
        shapes = (("n_rows", (n,), np.int32), ("node", (n, K), np.int16), ("score", (n, K), np.float32),
                  ("lwr", (n, K), np.float64), ("counts", (n, 4), np.int32), ("status", (n,), np.int32))
        outs = {}
        for name, shape, dt in shapes:
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name), pin_memory=True)
            keep.append(t)
            outs[name] = t.numpy()
        outs["node"] = outs["node"].view(np.uint16)
        hrb = synth.ReadBatch(h_seq, h_off.view(np.uint64))
        e2e_steps = max(1, min(args.steps, 10))

        def timed(batch, out):
            for _ in range(2):
                gdb.place(batch, cfg, out=out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                gdb.place(batch, cfg, out=out)
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        dt = timed(hrb, outs)
        assert np.array_equal(outs["n_rows"], dev_out["n_rows"]) and np.array_equal(outs["score"], dev_out["score"])
        # the same call from ordinary host memory (numpy arrays are pageable): the library stages through its own
        # pinned ring, the copies run on its helper threads
        pouts = {name: np.empty(shape, dt) for name, shape, dt in shapes}
        pouts["node"] = pouts["node"].view(np.uint16)
        dt_page = timed(rb, pouts)
        assert np.array_equal(pouts["n_rows"], dev_out["n_rows"]) and np.array_equal(pouts["score"], dev_out["score"])
        h2d = int(rb.seq.nbytes + rb.seq_off.nbytes)
        d2h = int(sum(v.nbytes for v in outs.values()))
        e2e = {"value": world * n / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "pageable": {"value": world * n / dt_page, "ms_per_step": dt_page * 1e3,
                            "note": "caller buffers in ordinary (unpinned) memory, staged by the library"},
               "affinity": ctx["affinity"],
               "api": "rp_place_batch (C ABI, host buffers from rp_host_alloc / pinned, 2-stream 64k-read chunks: "
                      "H2D / kernel / D2H overlapped)"}

    if rank != 0:
        return

    # ---- roofline of the (single) kernel: algorithmic bytes / event time vs measured HBM peak -------
    # lookups are exact (windows - skipped, + alternatives: from the kernel's counts when the reads carry no
    # ambiguity codes, else from the sample); postings per read come from rp_extract_kmers over a 200 k-read sample
    lens = np.diff(rb.seq_off.astype(np.int64))
    ns = min(n, 200_000)
    ex = gdb.extract(rb.slice(0, ns))
    s_lookups = int(ex["nalt"][ex["kind"] != 2].sum())
    s_postings = int(ex["hits"][ex["hits"] > 0].sum())
    lookups = int(round(s_lookups * (n / ns)))
    postings = int(round(s_postings * (n / ns)))
    rows = int(dev_out["n_rows"].sum())
    alg_bytes = int(lens.sum()) + 16 * lookups + 6 * postings + 16 * rows + 20 * n
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step / 1e3) / 1e9
    tbytes, bbytes = gdb.device_bytes()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(w.name, n), "kernel": "rp::place_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "algorithmic bytes = len + 16*lookups + 6*postings + 16*rows + 20 per read (SURVEY 8d), lookups and "
                        "postings per read from a %d-read sample; table %d MB + posting blocks %d MB vs 126 MB L2"
                        % (ns, tbytes >> 20, bbytes >> 20)}

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
        cpu = {"value": n_s / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample_reads": n_s,
               "sample": "first %d reads of the workload, %.1f s; C restatement of the Java loop (oracle/), 1 thread "
                         "like the reference (PlacementProcess.java:568); not the JVM (no JRE in the image or on the GPU box)" % (n_s, dt),
               "host_cores_available": os.cpu_count(), "matches_gpu_bit_exact": same,
               "published_reference": "~417-556 reads/s, 1 desktop core, RAPPAS v1.00 (README.md:244)"}

    layout = (("postings hash-partitioned over the GPUs, table replicated (gathers over NVLink peer memory), "
               "reads sharded, no collective") if args.partitioned and world > 1 and args.replicate_table else
              "hash-partitioned over the GPUs (peer memory over NVLink), reads sharded, no collective"
              if args.partitioned and world > 1 else "replicated per GPU, reads sharded, no collective")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, db.n_keys, db.n_postings, n, world, layout,
                              {"l2_policy": "inputs larger than L2 (reads+DB+outputs %d MB per step vs 126 MB)"
                                            % ((rb.seq.nbytes + tbytes + bbytes + n * (K * 14 + 24)) >> 20)}),
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


# --------------------------------------------------------------------------------------------- config 5
def run_config5(ctx):
    """1 M variable-length reads with ambiguity codes against the hash-defined DB, one partition per GPU, generated
    on the device; exchange form (default) or peer-memory form (--peer)."""
    import torch
    import torch.distributed as dist
    import rappas_b200 as R
    from rappas_b200 import _abi, exchange, synth, synth_hash
    args, rank, local_rank, world, dev = ctx["args"], ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    barrier, max_over_ranks, sum_over_ranks = ctx["barrier"], ctx["max_over_ranks"], ctx["sum_over_ranks"]
    w = synth.workload(5)
    if args.no_ambiguity:
        w.iupac_rate = w.n_rate = 0.0
    k5 = args.k5 or (15 if world > 1 else 13)
    w.k = k5
    w.name = "cfg5_stress_k%d_var_len_hashdb" % k5 if world > 1 else "cfg5_stress_standin_k%d_one_gpu" % k5
    n, scaling = job_reads(args, w, world)
    hdb = synth_hash.HashDB(k=k5, n_nodes=w.n_nodes, seed=42 + w.index, occupancy=0.75,
                            mean_postings=w.mean_postings * args.postings_scale)
    proxy = synth.SynthDB(0, k5, w.n_nodes, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                          np.zeros(0, np.uint16), np.zeros(0, np.float32))
    rb = synth.make_reads(proxy, n, w.read_len, seed=1042 + w.index + 7919 * rank, mode="uniform", iupac_rate=w.iupac_rate,
                          n_rate=w.n_rate)
    # page-locked copies of the batch, as a caller that used rp_host_alloc would hold it
    _pin = [torch.from_numpy(rb.seq).pin_memory(), torch.from_numpy(rb.seq_off.view(np.int64)).pin_memory()]
    rb = synth.ReadBatch(_pin[0].numpy(), _pin[1].numpy().view(np.uint64))
    cfg = _abi.place_cfg()
    K = cfg.keep_at_most
    t0 = time.perf_counter()
    part = R.Database.from_hash_db(hdb, local_rank, rank, world)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    tb, bb = part.device_bytes()
    x = None
    if world == 1:
        mode = "one GPU, whole (stand-in) DB"
    elif args.peer:
        part.attach_partitions_dist(local_rank)
        mode = "peer-memory form: probes and posting gathers read the owner's HBM over NVLink inside the kernel"
    else:
        uid = exchange.unique_id() if rank == 0 else np.zeros(_abi.RP_XCHG_ID_BYTES, np.uint8)
        t = torch.from_numpy(uid).to(dev)
        dist.broadcast(t, 0)
        x = exchange.Exchange.nccl(part, rank, world, t.cpu().numpy())
        mode = ("exchange form: NCCL all-to-all of k-mer probes to the owners, posting lists back (pipelined over "
                "sub-batches), placement kernel on the home GPU")
    steps = max(1, args.steps)

    def step():
        if x is not None:
            return x.place([rb], cfg)[0]
        return part.place(rb, cfg)
    for _ in range(max(1, min(args.warmup, 2))):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = R.kernel_launch_count()
    barrier()
    t_wall0 = time.time()
    dev_ms, e2e_s = [], []
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        out = step()
        e2e_s.append(time.perf_counter() - t0)
        dev_ms.append(x.stats()["device_ms"] if x is not None else part.last_kernel_ms())
    barrier()
    t_wall1 = time.time()
    launches = R.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_step = max_over_ranks(float(np.mean(dev_ms)))
    e2e_dt = max_over_ranks(float(np.mean(e2e_s)))
    if x is None and world == 1:  # one GPU: the usual device-resident leg (the host call sums its overlapping chunks' kernel times)
        ms_step, _, launches, clocks = device_leg(ctx, part, rb, cfg)
    # ---- parity: a sample of THIS rank's reads against the oracle over the regenerated sub-DB
    import oracle_lib as O
    ns = min(n, 300)
    sample = rb.slice(0, ns)
    sub = hdb.sub_db(synth_hash.probed_codes(sample, k5))
    odb = O.OracleDB(sub)
    t0 = time.perf_counter()
    oo = odb.place(sample, cfg)
    cpu_dt = time.perf_counter() - t0
    plain = oo["counts"][:, 2] == 0
    ok = bool(np.array_equal(oo["n_rows"][plain], out["n_rows"][:ns][plain]) and
              np.array_equal(oo["status"], out["status"][:ns]) and np.array_equal(oo["counts"], out["counts"][:ns]) and
              np.array_equal(oo["score"][plain].view(np.uint32), out["score"][:ns][plain].view(np.uint32)) and
              np.allclose(oo["score"][~plain], out["score"][:ns][~plain], rtol=1e-6, atol=0, equal_nan=True))
    oks = [ok]
    if world > 1:
        g = [None] * world
        dist.all_gather_object(g, ok)
        oks = g
    # ---- work done: lookups and postings of the whole job
    if x is not None:
        st = x.stats()
        lookups, postings = sum_over_ranks(float(st["probes"])), sum_over_ranks(float(st["owner_postings"]))
        payload = sum_over_ranks(float(st["payload_bytes"]))
    else:
        nsx = min(n, 20_000)
        ex = part.extract(rb.slice(0, nsx)) if world == 1 else None
        if ex is not None:
            lookups = float(ex["nalt"][ex["kind"] != 2].sum()) * n / nsx
            postings = float(ex["hits"][ex["hits"] > 0].sum()) * n / nsx
        else:
            lookups = postings = float("nan")
        lookups, postings, payload = sum_over_ranks(lookups), sum_over_ranks(postings), None
    total_len = sum_over_ranks(float(np.diff(rb.seq_off.astype(np.int64)).sum()))
    rows = sum_over_ranks(float(out["n_rows"].sum()))
    n_all = sum_over_ranks(float(n))
    keys_all, post_all = sum_over_ranks(float(part.desc.n_keys)), sum_over_ranks(float(part.desc.n_postings))
    dbytes_all = sum_over_ranks(float(tb + bb))
    if rank != 0:
        return
    alg_bytes = total_len + 16 * lookups + 6 * postings + 16 * rows + 20 * n_all
    peak, peak_src = measured_peak()
    achieved = alg_bytes / world / (ms_step / 1e3) / 1e9  # per GPU
    line = {
        "metric": METRIC, "value": n_all / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, keys_all, post_all, n, world, mode,
                              {"db_bytes_all_gpus": int(dbytes_all), "db_bytes_per_gpu": int(tb + bb),
                               "db_generated_on_device_s": gen_s,
                               "l2_policy": "DB far larger than L2 (%.1f GB per GPU)" % ((tb + bb) / 1e9)}),
        "kmer_lookups_per_sec": lookups / (ms_step / 1e3), "postings_per_sec": postings / (ms_step / 1e3),
        "matches_oracle": oks, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "rp::place_kernel (+ exchange kernels)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / world,
                     "note": "per GPU: the job's algorithmic bytes / n_gpus / step time; in the exchange form every posting "
                             "is additionally packed, sent over NVLink and read again (%.1f GB of posting blocks exchanged per step)"
                             % ((payload or 0) / 1e9)},
        "e2e": {"value": n_all / e2e_dt, "unit": UNIT, "ms_per_step": e2e_dt * 1e3,
                "h2d_bytes_per_step": int(rb.seq.nbytes + rb.seq_off.nbytes), "d2h_bytes_per_step": int(n * (K * 14 + 24)),
                "api": "rp_xchg_place / rp_place_batch (C ABI, host buffers)"},
        "cpu_baseline": {"value": ns / cpu_dt, "unit": UNIT, "cores": 1, "kind": "port", "sample_reads": ns,
                         "sample": "first %d reads of rank 0 against the host-regenerated sub-DB of the keys they probe" % ns},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
