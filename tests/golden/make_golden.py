#!/usr/bin/env python
"""Regenerates tests/golden/*.npz -- small fixed input/output vectors of the placement hot path.

PARITY UNPINNED: the Java reference cannot run here (no JDK/JRE, fastutil jar absent) and holds no
golden vectors of its own, so these vectors are produced by oracle/oracle_py.py, the literal
line-by-line Python transliteration of the Java control flow (PlacementProcess.processQueries,
AmbigSequenceKnife, fillBestScoreList -- each function cites its reference lines).  They pin the C
oracle (tests/test_golden.py, CPU) and the CUDA path (-m gpu) to a committed artefact, so a later
edit of either restatement cannot silently move the expected answers.

    python tests/golden/make_golden.py        # rewrites the .npz files (seeded, deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import oracle_py as P  # noqa: E402
from rappas_b200 import synth  # noqa: E402

K = 7  # keep_at_most of every fixture (ArgumentsParser_v2.java:86)

CASES = {
    # name: (db kwargs, reads kwargs, placement kwargs)
    "nucl_k6": (dict(alphabet=0, k=6, n_nodes=41, n_keys=2500, mean_postings=5, seed=0),
                dict(n_reads=48, length=(4, 90), seed=100, iupac_rate=0.03, n_rate=0.02, gap_rate=0.01, lowercase_rate=0.2),
                dict()),
    "nucl_k8_cfg1_like": (dict(alphabet=0, k=8, n_nodes=299, n_keys=6000, mean_postings=8, seed=43),
                          dict(n_reads=24, length=150, seed=1043), dict()),
    "nucl_k6_ambwithmax": (dict(alphabet=0, k=6, n_nodes=41, n_keys=2500, mean_postings=5, seed=1),
                           dict(n_reads=32, length=(4, 90), seed=101, iupac_rate=0.04, n_rate=0.02),
                           dict(with_max=True)),
    "nucl_k16_two_ambig": (dict(alphabet=0, k=16, n_nodes=23, n_keys=3000, mean_postings=4, seed=5, key_mode="genome"),
                           dict(n_reads=24, length=(10, 120), seed=6, mutation=0.01, iupac_rate=0.03, n_rate=0.01), dict()),
    "amino_k3": (dict(alphabet=1, k=3, n_nodes=31, n_keys=3000, mean_postings=6, seed=9),
                 dict(n_reads=40, length=(2, 60), seed=10, mutation=0.1, iupac_rate=0.04, n_rate=0.02, gap_rate=0.01,
                      lowercase_rate=0.3), dict()),
    "nucl_k4_ties_badchars": (dict(alphabet=0, k=4, n_nodes=9, n_keys=200, mean_postings=3, seed=3),
                              dict(n_reads=60, length=(3, 30), seed=4), dict(keep_factor=0.0)),
}


def build_case(name):
    dbkw, rdkw, plkw = CASES[name]
    db = synth.make_db(**dbkw)
    rb = synth.make_reads(db, **rdkw)
    if name == "nucl_k4_ties_badchars":
        db.post_score[:] = np.round(db.post_score * 2) / 2  # exact f32 ties between nodes: heap/sort tie order
        rb.seq[5] = ord("Z")
        rb.seq[40] = ord("@")
    return db, rb, plkw


def run_python_restatement(db, rb, plkw):
    sess = P.session_from_csr(db.alphabet, db.k, db.n_nodes, db.thr_lin, db.thr_log10, db.keys, db.offsets,
                              db.post_node, db.post_score)
    pp = P.PlacementProcess(sess)
    n = rb.n_reads
    out = {"status": np.zeros(n, np.int32), "counts": np.zeros((n, 4), np.int32), "n_rows": np.zeros(n, np.int32),
           "node": np.full((n, K), 0xFFFF, np.uint16), "score": np.full((n, K), -np.inf, np.float32),
           "lwr": np.zeros((n, K), np.float64), "S": np.full((n, db.n_nodes), np.nan, np.float32)}
    for r in range(n):
        res = pp.place_read(rb.read(r), keep_at_most=K, keep_factor=plkw.get("keep_factor", 0.01),
                            treat_amb=plkw.get("treat_amb", True), with_max=plkw.get("with_max", False))
        out["status"][r] = res.status
        if res.status in (2, 3):
            continue
        out["counts"][r] = res.counts
        out["n_rows"][r] = len(res.rows)
        for i, (node, sc, lwr) in enumerate(res.rows):
            out["node"][r, i], out["score"][r, i], out["lwr"][r, i] = node, sc, lwr
        for x, v in res.S.items():
            out["S"][r, x] = v
    return out


def main():
    for name in CASES:
        db, rb, plkw = build_case(name)
        out = run_python_restatement(db, rb, plkw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            alphabet=db.alphabet, k=db.k, n_nodes=db.n_nodes, thr_lin=db.thr_lin, thr_log10=db.thr_log10,
            keys=db.keys, offsets=db.offsets, post_node=db.post_node, post_score=db.post_score,
            seq=rb.seq, seq_off=rb.seq_off, keep_at_most=K, keep_factor=np.float32(plkw.get("keep_factor", 0.01)),
            treat_amb=plkw.get("treat_amb", True), with_max=plkw.get("with_max", False), **out)
        print(name, "reads", rb.n_reads, "placed", int((out["status"] == 0).sum()), "rows", int(out["n_rows"].sum()))


if __name__ == "__main__":
    main()
