"""-m gpu: the exchange form of a hash-partitioned DB (rp_xchg.cu) and the device-generated synthetic DB
(rp_synthdb.cu), against the CPU oracle.

On one GPU the ranks of the exchange are "virtual": all of them live in this process (rp_xchg_create_local) and the
collectives are device copies -- every bucketing / owner / pipeline path is the one the NCCL ranks run."""
import numpy as np
import pytest

import oracle_lib as O
import parity
from rappas_b200 import _abi, synth, synth_hash

pytestmark = pytest.mark.gpu


def amb_reads(o_out):
    return o_out["counts"][:, _abi.CNT_AMBIG] > 0


def check_against_oracle(outs, batches, o, cfg, So_fn=None):
    for gg, rb in zip(outs, batches):
        oo = o.place(rb, cfg)
        amb = None if cfg.amb_with_max else amb_reads(oo)
        So, _ = o.node_scores(rb, cfg, hitcount=False)
        parity.assert_placements_equal(gg, oo, cfg.keep_at_most, amb, So=So)


@pytest.mark.parametrize("env", [dict(), dict(RP_XCHG_PROBES="20000"), dict(RP_XCHG_COPY_LOCAL="1", RP_XCHG_PROBES="50000"),
                                 dict(RP_XCHG_PUSH="0", RP_XCHG_PROBES="30000"), dict(RP_XCHG_ONE_STREAM="1", RP_XCHG_PROBES="30000")],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in sorted(e.items())) or "default")
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_exchange_form_matches_the_oracle(world, env, monkeypatch):
    """Keys travel to their owners, posting lists travel back, the home GPU accumulates in window order: rows and
    per-node scores are those of the oracle (and of the replicated DB) bit for bit.  RP_XCHG_PROBES forces many
    sub-batches (the push | placement pipeline with its two buffers), RP_XCHG_COPY_LOCAL sends a rank's own partition
    through the buffers too, RP_XCHG_PUSH=0 takes the pack | all-to-all | placement form instead of owners writing
    into the homes' receive buffers, RP_XCHG_ONE_STREAM=1 places all sub-batches on one stream instead of two."""
    import rappas_b200 as R
    from rappas_b200 import exchange
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    db = synth.make_db(0, 10, 1999, n_keys=100000, mean_postings=24, seed=7)
    parts = [R.Database.partition_of_synth(db, 0, p, world) for p in range(world)]
    x = exchange.Exchange.local(parts)
    o = O.OracleDB(db)
    # every rank its own reads (different counts; one rank has none when there are several)
    batches = [synth.make_reads(db, 0 if (p == 1 and world > 2) else 400 + 130 * p, (50, 400), seed=8 + p, iupac_rate=0.005,
                                n_rate=0.002) for p in range(world)]
    try:
        for kw in (dict(), dict(amb_with_max=True), dict(treat_amb=False, keep_at_most=3)):
            cfg = _abi.place_cfg(**kw)
            outs = x.place(batches, cfg)
            check_against_oracle(outs, batches, o, cfg)
        st = x.stats()
        assert st["probes"] > 0 and (world == 1 and not env.get("RP_XCHG_COPY_LOCAL") or st["payload_bytes"] > 0)
        # the replicated DB gives the same bytes
        whole = R.Database.from_synth(db)
        cfg = _abi.place_cfg()
        outs = x.place(batches, cfg)
        for gg, rb in zip(outs, batches):
            ref = whole.place(rb, cfg)
            for key in ("n_rows", "node", "score", "lwr", "counts", "status"):
                assert np.array_equal(gg[key], ref[key], equal_nan=True), key
    finally:
        x.close()
        o.close()


def test_exchange_form_on_a_big_tree_with_long_reads():
    """config-5 shape in small: N = 9 999 (node-range passes), reads of 50-1500 bp with IUPAC / N characters, k = 12."""
    import rappas_b200 as R
    from rappas_b200 import exchange
    db = synth.make_db(0, 12, 9999, n_keys=200000, mean_postings=48, seed=45, key_mode="genome")
    world = 4
    parts = [R.Database.partition_of_synth(db, 0, p, world) for p in range(world)]
    x = exchange.Exchange.local(parts)
    o = O.OracleDB(db)
    batches = [synth.make_reads(db, 150, (50, 1500), seed=80 + p, mutation=0.02, iupac_rate=0.005, n_rate=0.002) for p in range(world)]
    try:
        cfg = _abi.place_cfg()
        check_against_oracle(x.place(batches, cfg), batches, o, cfg)
    finally:
        x.close()
        o.close()


@pytest.mark.parametrize("k,n_nodes,mean", [(8, 299, 16), (11, 9999, 48)])
def test_device_generated_db_equals_its_host_restatement(k, n_nodes, mean):
    """rp_db_synth_partition builds the hash-defined DB on the device; synth_hash.HashDB regenerates, on the host,
    the keys a read sample probes.  The oracle over that sub-DB and the GPU over the whole DB must agree on every
    k-mer hit, score and row -- which pins the generator (presence, list length, nodes, scores) to its definition."""
    import rappas_b200 as R
    hdb = synth_hash.HashDB(k=k, n_nodes=n_nodes, seed=1234 + k, occupancy=0.75, mean_postings=mean)
    g = R.Database.from_hash_db(hdb)
    expect = hdb.expected_keys()
    assert abs(g.desc.n_keys - expect) < 6 * np.sqrt(expect) + 8
    proxy = synth.SynthDB(0, k, n_nodes, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                          np.zeros(0, np.uint16), np.zeros(0, np.float32))
    rb = synth.make_reads(proxy, 600, (20, 300), seed=5, mode="uniform", iupac_rate=0.004, n_rate=0.002)
    sub = hdb.sub_db(synth_hash.probed_codes(rb, k))
    assert 0.7 < sub.n_keys / max(1, len(synth_hash.probed_codes(rb, k))) < 0.8  # occupancy of the probed codes
    o = O.OracleDB(sub)
    try:
        parity.assert_extract_equal(g.extract(rb), o.extract(rb))
        cfg = _abi.place_cfg()
        oo = o.place(rb, cfg)
        So, _ = o.node_scores(rb, cfg, hitcount=False)
        parity.assert_scores_equal(g.node_scores(rb, cfg), So, amb_reads(oo))
        parity.assert_placements_equal(g.place(rb, cfg), oo, cfg.keep_at_most, amb_reads(oo), So=So)
    finally:
        g.close()
        o.close()


def test_device_generated_partitions_behind_the_exchange():
    """The config-5 construction in small: every rank generates ITS partition of the hash-defined DB on the device
    (no rank, and no host, ever holds the whole DB), reads are placed through the exchange, and the oracle checks
    them over the sub-DB of the probed keys."""
    import rappas_b200 as R
    from rappas_b200 import exchange
    world, k, n_nodes = 4, 11, 9999
    hdb = synth_hash.HashDB(k=k, n_nodes=n_nodes, seed=99, occupancy=0.75, mean_postings=48)
    parts = [R.Database.from_hash_db(hdb, 0, p, world) for p in range(world)]
    total = sum(p.desc.n_keys for p in parts)
    assert abs(total - hdb.expected_keys()) < 6 * np.sqrt(hdb.expected_keys())
    assert min(p.desc.n_keys for p in parts) > 0.9 * total / world
    x = exchange.Exchange.local(parts)
    proxy = synth.SynthDB(0, k, n_nodes, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                          np.zeros(0, np.uint16), np.zeros(0, np.float32))
    batches = [synth.make_reads(proxy, 120, (50, 1500), seed=40 + p, mode="uniform", iupac_rate=0.005, n_rate=0.002)
               for p in range(world)]
    allreads = synth.reads_from_strings([b.read(i) for b in batches for i in range(b.n_reads)])
    o = O.OracleDB(hdb.sub_db(synth_hash.probed_codes(allreads, k)))
    try:
        cfg = _abi.place_cfg()
        check_against_oracle(x.place(batches, cfg), batches, o, cfg)
    finally:
        x.close()
        o.close()
