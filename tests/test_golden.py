"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the literal
Python restatement of the Java code) against the C oracle (CPU) and the CUDA path (-m gpu, through the C ABI).

PARITY UNPINNED at the source: the reference has no golden vectors and cannot run here; these files pin
our two restatements and the CUDA library to one committed artefact.
"""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
import parity
from rappas_b200 import _abi, synth

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz"))
                if not os.path.basename(p).startswith("dbbuild_"))  # those belong to tests/test_golden_dbbuild.py


def load(path):
    z = np.load(path)
    db = synth.SynthDB(int(z["alphabet"]), int(z["k"]), int(z["n_nodes"]), np.float32(z["thr_lin"]),
                       np.float32(z["thr_log10"]), z["keys"], z["offsets"], z["post_node"], z["post_score"])
    rb = synth.ReadBatch(np.ascontiguousarray(z["seq"]), np.ascontiguousarray(z["seq_off"]))
    cfg = _abi.place_cfg(keep_at_most=int(z["keep_at_most"]), keep_factor=float(z["keep_factor"]),
                         treat_amb=bool(z["treat_amb"]), amb_with_max=bool(z["with_max"]))
    exp = {k: z[k] for k in ("status", "counts", "n_rows", "node", "score", "lwr", "S")}
    return db, rb, cfg, exp


def test_fixtures_present():
    assert len(GOLDEN) >= 6
    assert sum(os.path.getsize(p) for p in GOLDEN) < 2 << 20, "golden fixtures are meant to stay small"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_c_oracle_matches_golden(path):
    """Exact: same libm under the C oracle and the Python restatement that wrote the vectors."""
    db, rb, cfg, exp = load(path)
    o = O.OracleDB(db)
    out = o.place(rb, cfg)
    assert np.array_equal(out["status"], exp["status"])
    ok = exp["status"] <= 1
    assert np.array_equal(out["counts"][ok], exp["counts"][ok])
    assert np.array_equal(out["n_rows"], exp["n_rows"])
    for r in range(rb.n_reads):
        n = int(exp["n_rows"][r])
        assert np.array_equal(out["node"][r, :n], exp["node"][r, :n]), r
        assert np.array_equal(out["score"][r, :n].view(np.uint32), exp["score"][r, :n].view(np.uint32)), r
        assert np.array_equal(out["lwr"][r, :n], exp["lwr"][r, :n]), r
    S, _ = o.node_scores(rb, cfg, hitcount=False)
    assert np.array_equal(np.isnan(S), np.isnan(exp["S"]))
    m = ~np.isnan(exp["S"])
    assert np.array_equal(S[m].view(np.uint32), exp["S"][m].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_golden(path):
    import rappas_b200 as R
    db, rb, cfg, exp = load(path)
    g = R.Database.from_synth(db)
    out = g.place(rb, cfg)
    amb = exp["counts"][:, _abi.CNT_AMBIG] > 0 if not cfg.amb_with_max else None
    # same bars as the oracle parity tests: statuses / counts exact, plain-window scores bit-exact,
    # ambiguity-path scores 1e-6 (two libms), nodes identical except among exactly tied scores
    parity.assert_placements_equal(out, exp, cfg.keep_at_most, amb)
    parity.assert_scores_equal(g.node_scores(rb, cfg), exp["S"], amb)
    g.close()
