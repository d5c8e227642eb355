"""-m gpu: parity of the CUDA path (called through the C ABI) against the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import parity
from rappas_b200 import _abi, synth
from rappas_b200.synth import reads_from_strings

pytestmark = pytest.mark.gpu


def both(db):
    import rappas_b200 as R
    return R.Database.from_synth(db), O.OracleDB(db)


def amb_reads(o_out):
    return o_out["counts"][:, _abi.CNT_AMBIG] > 0


CASES = {
    "nucl_k6_ambig": dict(db=dict(alphabet=0, k=6, n_nodes=41, n_keys=2500, mean_postings=5, seed=0),
                          reads=dict(n_reads=300, length=(4, 200), seed=100, iupac_rate=0.03, n_rate=0.02,
                                     gap_rate=0.01, lowercase_rate=0.2)),
    "nucl_k8_cfg1_like": dict(db=dict(alphabet=0, k=8, n_nodes=299, n_keys=49152, mean_postings=16, seed=43),
                              reads=dict(n_reads=2000, length=150, seed=1043)),
    "nucl_k10_long_lists": dict(db=dict(alphabet=0, k=10, n_nodes=1999, n_keys=100000, mean_postings=120, seed=7),
                                reads=dict(n_reads=500, length=(50, 1500), seed=8, iupac_rate=0.005,
                                           n_rate=0.002)),
    "nucl_k16_two_ambig": dict(db=dict(alphabet=0, k=16, n_nodes=23, n_keys=3000, mean_postings=4, seed=5,
                                       key_mode="genome"),
                               reads=dict(n_reads=200, length=(10, 120), seed=6, mutation=0.01, iupac_rate=0.03,
                                          n_rate=0.01)),
    "nucl_k12_big_tree": dict(db=dict(alphabet=0, k=12, n_nodes=9999, n_keys=200000, mean_postings=48, seed=45,
                                      key_mode="genome"),
                              reads=dict(n_reads=400, length=150, seed=1045, mutation=0.02)),
    "amino_k3": dict(db=dict(alphabet=1, k=3, n_nodes=31, n_keys=3000, mean_postings=6, seed=9),
                     reads=dict(n_reads=300, length=(2, 60), seed=10, mutation=0.1, iupac_rate=0.04, n_rate=0.02,
                                gap_rate=0.01, lowercase_rate=0.3)),
    "amino_k6_cfg4_like": dict(db=dict(alphabet=1, k=6, n_nodes=999, n_keys=200000, mean_postings=16, seed=46),
                               reads=dict(n_reads=1000, length=50, seed=1046, iupac_rate=0.01)),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    spec = CASES[request.param]
    db = synth.make_db(**spec["db"])
    rb = synth.make_reads(db, **spec["reads"])
    g, o = both(db)
    yield request.param, db, rb, g, o
    g.close()
    o.close()


def test_kmer_extraction_and_lookup_bit_exact(case):
    _, db, rb, g, o = case
    parity.assert_extract_equal(g.extract(rb), o.extract(rb))


@pytest.mark.parametrize("with_max", [False, True])
def test_node_scores(case, with_max):
    _, db, rb, g, o = case
    cfg = _abi.place_cfg(amb_with_max=with_max)
    So, _ = o.node_scores(rb, cfg, hitcount=False)
    Sg = g.node_scores(rb, cfg)
    oo = o.place(rb, cfg)
    # with --ambwithmax the ambiguity path is pure f32: every read must be bit-exact
    parity.assert_scores_equal(Sg, So, None if with_max else amb_reads(oo))


@pytest.mark.parametrize("kw", [dict(), dict(keep_at_most=1), dict(keep_at_most=3, keep_factor=0.5),
                                dict(keep_at_most=32, keep_factor=0.0), dict(treat_amb=False),
                                dict(amb_with_max=True), dict(ns_bound=-200.0)])
def test_placements(case, kw):
    _, db, rb, g, o = case
    cfg = _abi.place_cfg(**kw)
    oo, gg = o.place(rb, cfg), g.place(rb, cfg)
    amb = amb_reads(oo) if not kw.get("amb_with_max") else None
    # a node that differs from the oracle's must be an exact tie in the oracle's own score vector
    So, _ = o.node_scores(rb, cfg, hitcount=False)
    parity.assert_placements_equal(gg, oo, cfg.keep_at_most, amb, So=So)


def test_edge_reads():
    db = synth.make_db(0, 5, 17, n_keys=600, mean_postings=4, seed=1)
    g, o = both(db)
    reads = ["", "A", "ACGT", "ACGTA", "ACGTAC", "NNNNNNNN", "ACGTZCGTA", "@", "acgtnacgtnacgtRYKM", "-" * 12,
             "A" * 4000, "ACGU" * 30, "ACG.TACGT-ACGTA"]
    rb = reads_from_strings(reads)
    parity.assert_extract_equal(g.extract(rb), o.extract(rb))
    oo, gg = o.place(rb), g.place(rb)
    parity.assert_placements_equal(gg, oo, 7, amb_reads(oo))
    assert list(gg["status"][:4]) == [2, 2, 1, gg["status"][3]]
    assert gg["status"][6] == 3 and gg["status"][7] == 3


def test_empty_batch_and_errors():
    import rappas_b200 as R
    from rappas_b200._lib import RappasError
    db = synth.make_db(0, 5, 17, n_keys=600, mean_postings=4, seed=1)
    g = R.Database.from_synth(db)
    out = g.place(reads_from_strings([]))
    assert out["n_rows"].shape == (0,)
    with pytest.raises(RappasError):
        g.place(reads_from_strings(["ACGTACGT"]), _abi.place_cfg(keep_at_most=0))
    with pytest.raises(RappasError):
        g.place(reads_from_strings(["ACGTACGT"]), _abi.place_cfg(keep_at_most=33))
    bad = synth.make_db(0, 5, 17, n_keys=600, mean_postings=4, seed=1)
    bad.post_node[3] = 17
    with pytest.raises(RappasError):
        R.Database.from_synth(bad)
    dup = synth.make_db(0, 5, 17, n_keys=600, mean_postings=4, seed=1)
    dup.keys[1] = dup.keys[0]
    with pytest.raises(RappasError):
        R.Database.from_synth(dup)
    with pytest.raises(RappasError):
        R.Database.from_arrays(0, 40, 17, db.thr_lin, db.thr_log10, db.keys, db.offsets, db.post_node, db.post_score)


def test_unsorted_postings_and_empty_db():
    import rappas_b200 as R
    db = synth.make_db(0, 6, 200, n_keys=3000, mean_postings=40, seed=2)
    # shuffle postings inside every key: the loader re-sorts by node, results must not change
    rng = np.random.default_rng(0)
    for i in range(db.n_keys):
        lo, hi = int(db.offsets[i]), int(db.offsets[i + 1])
        p = rng.permutation(hi - lo)
        db.post_node[lo:hi] = db.post_node[lo:hi][p]
        db.post_score[lo:hi] = db.post_score[lo:hi][p]
    rb = synth.make_reads(db, 300, 100, seed=3)
    g, o = both(db)
    oo, gg = o.place(rb), g.place(rb)
    parity.assert_placements_equal(gg, oo, 7)
    So, _ = o.node_scores(rb, hitcount=False)
    parity.assert_scores_equal(g.node_scores(rb), So)
    empty = synth.SynthDB(0, 6, 10, db.thr_lin, db.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                          np.zeros(0, np.uint16), np.zeros(0, np.float32))
    ge = R.Database.from_synth(empty)
    out = ge.place(rb.slice(0, 5))
    assert (out["status"] == 1).all() and (out["n_rows"] == 0).all()


def test_rgdb_file_roundtrip(tmp_path):
    import rappas_b200 as R
    from rappas_b200._lib import check, load
    db = synth.make_db(1, 4, 77, n_keys=5000, mean_postings=7, seed=11)
    rb = synth.make_reads(db, 200, 40, seed=12, mutation=0.05)
    path = str(tmp_path / "db.rgdb")
    desc = _abi.RpDbDesc(db.alphabet, db.k, db.n_nodes, float(db.thr_log10), float(db.thr_lin), 0, db.n_keys,
                         db.n_postings)
    check(load()["db_save_file"](path.encode(), C.byref(desc), _abi.ptr(db.keys), _abi.ptr(db.offsets),
                                 _abi.ptr(db.post_node), _abi.ptr(db.post_score)))
    a = R.Database.from_synth(db)
    b = R.Database.from_file(path)
    assert (b.desc.k, b.desc.n_nodes, b.desc.n_keys, b.desc.n_postings) == (db.k, db.n_nodes, db.n_keys, db.n_postings)
    oa, ob = a.place(rb), b.place(rb)
    for key in oa:
        assert np.array_equal(oa[key], ob[key], equal_nan=True)


def test_device_pointer_entry_matches_host_entry():
    import torch
    import rappas_b200 as R
    db = synth.make_db(0, 8, 299, n_keys=49152, mean_postings=16, seed=43)
    rb = synth.make_reads(db, 5000, 150, seed=1043)
    g = R.Database.from_synth(db)
    cfg = _abi.place_cfg()
    host = g.place(rb, cfg)
    n, K = rb.n_reads, cfg.keep_at_most
    dev = torch.device("cuda:0")
    d_seq = torch.from_numpy(rb.seq).to(dev)
    d_off = torch.from_numpy(rb.seq_off.view(np.int64)).to(dev)
    d_n = torch.empty(n, dtype=torch.int32, device=dev)
    d_node = torch.empty((n, K), dtype=torch.int16, device=dev)
    d_score = torch.empty((n, K), dtype=torch.float32, device=dev)
    d_lwr = torch.empty((n, K), dtype=torch.float64, device=dev)
    d_cnt = torch.empty((n, 4), dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream()
    for _ in range(2):  # second launch re-uses scheduler counter + scratch
        g.place_device(cfg, d_seq.data_ptr(), d_off.data_ptr(), n, d_n.data_ptr(), d_node.data_ptr(),
                       d_score.data_ptr(), d_lwr.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_n.cpu().numpy(), host["n_rows"])
    assert np.array_equal(d_node.cpu().numpy().view(np.uint16), host["node"])
    assert np.array_equal(d_score.cpu().numpy(), host["score"])
    assert np.array_equal(d_lwr.cpu().numpy(), host["lwr"])
    assert np.array_equal(d_cnt.cpu().numpy(), host["counts"])
    assert np.array_equal(d_st.cpu().numpy(), host["status"])


def test_batch_split_and_order_invariance():
    db = synth.make_db(0, 8, 299, n_keys=49152, mean_postings=16, seed=43)
    rb = synth.make_reads(db, 3000, (20, 300), seed=5, n_rate=0.003)
    import rappas_b200 as R
    g = R.Database.from_synth(db)
    whole = g.place(rb)
    a, b = g.place(rb.slice(0, 1234)), g.place(rb.slice(1234, 3000))
    for key in whole:
        assert np.array_equal(whole[key], np.concatenate([a[key], b[key]]), equal_nan=True), key
    again = g.place(rb)  # idempotent: no state survives a read / a batch
    for key in whole:
        assert np.array_equal(whole[key], again[key], equal_nan=True)


def test_multi_device_in_process_matches_single_device():
    """DB replicated on every visible GPU, reads split in contiguous slices inside rp_place_batch
    (no collective): identical rows whatever the device count."""
    import rappas_b200 as R
    nd = R.device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    db = synth.make_db(0, 8, 299, n_keys=49152, mean_postings=16, seed=43)
    rb = synth.make_reads(db, 20001, (20, 300), seed=5, n_rate=0.003)
    one = R.Database.from_synth(db, devices=(0,))
    many = R.Database.from_synth(db, devices=tuple(range(nd)))
    a, b = one.place(rb), many.place(rb)
    for key in a:
        assert np.array_equal(a[key], b[key], equal_nan=True), key


@pytest.mark.parametrize("layout", [1, 2], ids=["table+postings", "postings-only"])
@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_hash_partitioned_db_on_one_device(n_parts, layout):
    """Partitioned mode with every partition in the same HBM (how a 1-GPU box exercises it): same rows,
    same k-mer hits and same per-node scores as the replicated DB and the oracle.  layout 1 partitions the
    table with the postings, layout 2 keeps the whole table beside every partition of the postings."""
    import rappas_b200 as R
    db = synth.make_db(0, 10, 1999, n_keys=100000, mean_postings=24, seed=7)
    rb = synth.make_reads(db, 1500, (50, 400), seed=8, iupac_rate=0.005, n_rate=0.002)
    part = R.Database.from_synth(db, devices=(0,) * n_parts, partitioned=layout)
    o = O.OracleDB(db)
    parity.assert_extract_equal(part.extract(rb), o.extract(rb))
    oo = o.place(rb)
    parity.assert_placements_equal(part.place(rb), oo, 7, amb_reads(oo))
    So, _ = o.node_scores(rb, hitcount=False)
    parity.assert_scores_equal(part.node_scores(rb), So, amb_reads(oo))
    tb, bb = part.device_bytes()
    assert bb > 0 and tb >= n_parts * 32 * 32


@pytest.mark.parametrize("layout", [1, 2], ids=["table+postings", "postings-only"])
def test_hash_partitioned_db_over_peer_memory(layout):
    """One partition per GPU; every GPU places its slice of the reads, probing and gathering from the
    owner's HBM over NVLink (peer-mapped pointers, no collective)."""
    import rappas_b200 as R
    nd = R.device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    db = synth.make_db(0, 10, 1999, n_keys=200000, mean_postings=32, seed=7)
    rb = synth.make_reads(db, 30001, 150, seed=8, n_rate=0.002)
    part = R.Database.from_synth(db, devices=tuple(range(nd)), partitioned=layout)
    one = R.Database.from_synth(db, devices=(0,))
    a, b = one.place(rb), part.place(rb)
    for key in a:
        assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_fasta_to_jplace_through_the_gpu(tmp_path):
    """FASTA text -> native ingest (unique sequences) -> CUDA placement -> .jplace, against the same
    pipeline fed by the oracle (rows identical except among exactly tied scores)."""
    import json
    import rappas_b200 as R
    from rappas_b200 import ingest
    db = synth.make_db(0, 8, 299, n_keys=49152, mean_postings=16, seed=43)
    rb = synth.make_reads(db, 400, (20, 300), seed=5, n_rate=0.003)
    lines = []
    for i in range(rb.n_reads):
        lines += [">read%d len=%d" % (i, len(rb.read(i))), rb.read(i)]
    for i in range(0, 100, 7):
        lines += [">again%d" % i, rb.read(i)]
    q = ingest.QueryFile.from_text("\n".join(lines) + "\n")
    assert q.n_records == 415 and q.n_unique <= 400
    cfg = _abi.place_cfg()
    g, o = both(db)
    rg, ro = g.place(q.unique, cfg), o.place(q.unique, cfg)
    edge, bl = np.arange(db.n_nodes, dtype=np.int32), np.linspace(0.01, 1.0, db.n_nodes).astype(np.float32)
    ng = q.write_jplace(tmp_path / "g.jplace", rg, 7, edge, bl, invocation="t")
    no = q.write_jplace(tmp_path / "o.jplace", ro, 7, edge, bl, invocation="t")
    dg, do = json.loads((tmp_path / "g.jplace").read_text()), json.loads((tmp_path / "o.jplace").read_text())
    assert ng == no == len(dg["placements"]) and ng > 300
    for pg, po in zip(dg["placements"], do["placements"]):
        assert pg["nm"] == po["nm"] and abs(len(pg["p"]) - len(po["p"])) <= 1
        # likelihood column: bit-identical floats, except on reads with ambiguity codes (f32 exp10f / log10f on the GPU:
        # an ulp or two of the score, parity.SCORE_RTOL_AMBIG; a row at the keep-factor cut may then come or go)
        m = min(len(pg["p"]), len(po["p"]))
        lg, lo = [r[1] for r in pg["p"][:m]], [r[1] for r in po["p"][:m]]
        if lg != lo or len(pg["p"]) != len(po["p"]):
            assert np.allclose(lg, lo, rtol=parity.SCORE_RTOL_AMBIG, atol=parity.SCORE_ATOL_AMBIG)


@pytest.mark.parametrize("n_nodes,mean", [(40000, 60), (65535, 200)])
def test_very_large_trees_up_to_the_char_limit(n_nodes, mean):
    """Trees up to the reference's limit (node ids are Java chars: 65 535, Pair_16_32_bit.java:26-29,
    PlacementProcess.java:495-496).  S[] of such a read does not fit an SM, so the read is walked in node-range
    passes, S holding one slice at a time; rows and per-node scores stay those of the oracle bit for bit."""
    db = synth.make_db(0, 8, n_nodes, n_keys=20000, mean_postings=mean, seed=3)
    rb = synth.make_reads(db, 300, (30, 400), seed=4, n_rate=0.003)
    g, o = both(db)
    oo = o.place(rb)
    So, _ = o.node_scores(rb, hitcount=False)
    parity.assert_placements_equal(g.place(rb), oo, 7, amb_reads(oo), So=So)
    parity.assert_scores_equal(g.node_scores(rb), So, amb_reads(oo))


@pytest.mark.parametrize("env", [dict(RP_PASSES="1"), dict(RP_PASSES="2"), dict(RP_PASSES="3", RP_STAGE_BYTES="1024"),
                                 dict(RP_PASSES="16"), dict(RP_NO_DIRECT="1")],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in sorted(e.items())))
@pytest.mark.parametrize("name", ["nucl_k8_cfg1_like", "nucl_k10_long_lists", "nucl_k12_big_tree", "amino_k6_cfg4_like"])
def test_node_range_passes_and_table_forms(name, env, monkeypatch):
    """The geometry knobs (read when the DB is loaded) force the same reads through 1, 2, 3 and 16 node-range
    passes (S = one slice of the tree at a time, windows routed by the node range of their list) and through
    the cuckoo table where a direct-address table would be used: rows and scores must not move by a bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    spec = CASES[name]
    db = synth.make_db(**spec["db"])
    rb = synth.make_reads(db, **spec["reads"])
    g, o = both(db)
    try:
        parity.assert_extract_equal(g.extract(rb), o.extract(rb))
        for kw in (dict(), dict(keep_at_most=3, keep_factor=0.5), dict(amb_with_max=True)):
            cfg = _abi.place_cfg(**kw)
            oo = o.place(rb, cfg)
            amb = None if kw.get("amb_with_max") else amb_reads(oo)
            So, _ = o.node_scores(rb, cfg, hitcount=False)
            parity.assert_scores_equal(g.node_scores(rb, cfg), So, amb)
            parity.assert_placements_equal(g.place(rb, cfg), oo, cfg.keep_at_most, amb, So=So)
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("env", [dict(RP_STAGE_BYTES="1024"), dict(RP_STAGE_BYTES="16384"), dict(RP_PASSES="2"),
                                 dict(RP_PASSES="4", RP_STAGE_BYTES="2048"), dict(RP_PASSES="2", RP_AMB_BATCH="0"),
                                 dict(RP_PASSES="3", RP_AMB_BATCH="1", RP_STAGE_BYTES="1536")],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in sorted(e.items())))
@pytest.mark.parametrize("name", ["nucl_k6_ambig", "nucl_k10_long_lists", "nucl_k16_two_ambig", "amino_k3"])
def test_ambiguous_windows_staged_and_from_global_memory(name, env, monkeypatch):
    """An ambiguous window is a group of its own: its alternatives' blocks are staged and S_amb / C_amb is a
    table in the stage's tail -- or, when they do not fit, the consumer walks them in global memory.  The
    geometry knobs (read when the DB is loaded) push the same reads down both branches, and through 2 and 4
    node-range passes (the table then holds the nodes of the pass's slice only).  Sliced trees have two builds of the
    kernel -- ambiguous windows one per group, or up to eight consecutive ones per group -- and a count of the batch's
    ambiguity characters picks one on the device; RP_AMB_BATCH forces either."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    spec = CASES[name]
    db = synth.make_db(**spec["db"])
    rb = synth.make_reads(db, **spec["reads"])
    g, o = both(db)
    try:
        for with_max in (False, True):
            cfg = _abi.place_cfg(amb_with_max=with_max)
            oo = o.place(rb, cfg)
            assert amb_reads(oo).sum() > 20
            amb = None if with_max else amb_reads(oo)  # --ambwithmax is pure f32: bit-exact
            So, _ = o.node_scores(rb, cfg, hitcount=False)
            parity.assert_scores_equal(g.node_scores(rb, cfg), So, amb)
            parity.assert_placements_equal(g.place(rb, cfg), oo, cfg.keep_at_most, amb, So=So)
    finally:
        g.close()
        o.close()


def test_pinned_and_pageable_callers_give_the_same_rows():
    """rp_place_batch pipelines chunks over two streams.  Buffers from rp_host_alloc are copied from / to directly;
    ordinary (pageable) numpy arrays are detected and staged through the library's own pinned ring.  Same bytes out,
    also with a chunk size that makes the batch span many pipeline steps."""
    import rappas_b200 as R
    from rappas_b200._lib import check, load
    fn = load()
    db = synth.make_db(0, 8, 299, n_keys=49152, mean_postings=16, seed=43)
    rb = synth.make_reads(db, 7001, (20, 300), seed=5, n_rate=0.003)
    g = R.Database.from_synth(db)
    cfg = _abi.place_cfg()
    K, n = cfg.keep_at_most, rb.n_reads
    ref = g.place(rb, cfg)  # pageable in, pageable out
    ptrs = []

    def pinned(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(fn["host_alloc"](C.byref(p), max(nbytes, 1)))
        ptrs.append(p)
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    try:
        seq = pinned(rb.seq.shape, np.uint8); seq[:] = rb.seq
        off = pinned(rb.seq_off.shape, np.uint64); off[:] = rb.seq_off
        out = {"n_rows": pinned((n,), np.int32), "node": pinned((n, K), np.uint16), "score": pinned((n, K), np.float32),
               "lwr": pinned((n, K), np.float64), "counts": pinned((n, 4), np.int32), "status": pinned((n,), np.int32)}
        os.environ["RP_CHUNK_READS"] = "500"
        try:
            got = g.place(synth.ReadBatch(seq, off), cfg, out=out)
            again = g.place(rb, cfg)
        finally:
            del os.environ["RP_CHUNK_READS"]
        for key in ref:
            assert np.array_equal(ref[key], got[key], equal_nan=True), key
            assert np.array_equal(ref[key], again[key], equal_nan=True), key
        # mixed: pinned inputs, pageable outputs
        mixed = g.place(synth.ReadBatch(seq, off), cfg)
        for key in ref:
            assert np.array_equal(ref[key], mixed[key], equal_nan=True), key
    finally:
        del seq, off, out
        for p in ptrs:
            fn["host_free"](p)


def test_kmers_and_hits_of_the_placement_kernel_itself(case):
    """K1 + K2 measured ON place_kernel: the codes its producer builds from ballots and funnel shifts and what its
    probes (cuckoo buckets / direct-address table) find, window by window, against the oracle's AmbigSequenceKnife +
    hash lookup (rp_extract_kmers is a separate kernel and is held to the same oracle above)."""
    _, db, rb, g, o = case
    ex = o.extract(rb)
    pw = g.place_windows(rb)
    woff = ex["win_off"].astype(np.int64)
    ok_read = ex["status"] == 0
    visited = pw["hits"] != np.int32(-50529028)  # 0xFCFCFCFC: windows behind an unsupported character
    win_read = np.repeat(np.arange(rb.n_reads), np.diff(woff))
    assert visited[ok_read[win_read]].all()
    plain = (ex["kind"] == _abi.WIN_PLAIN) & ok_read[win_read]
    assert np.array_equal(pw["code"][plain], ex["code"][plain])
    assert np.array_equal(pw["hits"][plain], ex["hits"][plain])
    assert (pw["hits"][(ex["kind"] == _abi.WIN_SKIPPED) & ok_read[win_read]] == -2).all()
    assert (pw["hits"][(ex["kind"] == _abi.WIN_AMBIG) & ok_read[win_read]] == -3).all()


def test_the_full_config3_database():
    """BASELINE.json configs[2] at its full DB size (k = 12, 9 999 nodes, 12.6 M keys, 604 M postings, 4.2 GB in HBM;
    direct-address table, two node-range passes): a sample of its reads against the oracle -- extraction, per-node
    scores and rows.  This is the DB the default bench line is measured on."""
    w = synth.workload(3)
    db = synth.make_db(w.alphabet, w.k, w.n_nodes, w.n_keys, w.mean_postings, seed=42 + w.index, key_mode=w.key_mode)
    rb = synth.make_reads(db, 4000, w.read_len, seed=17, iupac_rate=0.002, n_rate=0.001)
    g, o = both(db)
    try:
        parity.assert_extract_equal(g.extract(rb), o.extract(rb))
        for kw in (dict(), dict(amb_with_max=True)):
            cfg = _abi.place_cfg(**kw)
            oo = o.place(rb, cfg)
            So, _ = o.node_scores(rb, cfg, hitcount=False)
            amb = None if kw else amb_reads(oo)
            parity.assert_placements_equal(g.place(rb, cfg), oo, 7, amb, So=So)
            parity.assert_scores_equal(g.node_scores(rb, cfg), So, amb)
    finally:
        g.close()
        o.close()


def test_k15_long_reads_of_the_config5_shape():
    """BASELINE.json configs[4] in shape: k = 15 (30-bit keys, cuckoo table -- no direct form), 9 999 nodes, reads of
    50-1500 bp with IUPAC codes and N runs, lists of ~48 postings; the replicated DB and, with the same keys split
    over 3 owners, the exchange form."""
    import rappas_b200 as R
    from rappas_b200 import exchange
    db = synth.make_db(0, 15, 9999, n_keys=1_500_000, mean_postings=48, seed=47, key_mode="genome")
    rb = synth.make_reads(db, 500, (50, 1500), seed=5, mutation=0.02, iupac_rate=0.005, n_rate=0.002)
    g, o = both(db)
    parts = [R.Database.partition_of_synth(db, 0, p, 3) for p in range(3)]
    x = exchange.Exchange.local(parts)
    try:
        parity.assert_extract_equal(g.extract(rb), o.extract(rb))
        cfg = _abi.place_cfg()
        oo = o.place(rb, cfg)
        assert oo["counts"][:, _abi.CNT_MATCHED].sum() > 20 * rb.n_reads  # the reads do find their genome
        So, _ = o.node_scores(rb, cfg, hitcount=False)
        gg = g.place(rb, cfg)
        parity.assert_placements_equal(gg, oo, 7, amb_reads(oo), So=So)
        empty = synth.make_reads(db, 0, (50, 60), seed=1)
        xx = x.place([rb, empty, empty], cfg)[0]
        for key in ("n_rows", "node", "score", "lwr", "counts", "status"):
            assert np.array_equal(xx[key], gg[key], equal_nan=True), key
    finally:
        x.close()
        g.close()
        o.close()
