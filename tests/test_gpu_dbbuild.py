"""-m gpu: the GPU phylo-k-mer generation (rp_dbbuild_run through the C ABI) against the recursive CPU oracle.
Bit-exact: the f32 running sum makes the explorers' visiting order part of the result."""
import numpy as np
import pytest

import dbbuild_lib as D
import oracle_lib as O
import parity
from rappas_b200 import _abi, dbbuild, synth

pytestmark = pytest.mark.gpu


def thr_of(alphabet, k, omega=1.5):
    return float(O.threshold(omega, alphabet, k)[1])


@pytest.mark.parametrize("alphabet,k,n_nodes,n_sites,gap_jumps,peak", [
    (0, 5, 12, 60, 0, 0.9), (0, 8, 9, 80, 0, 0.93), (0, 6, 7, 50, 1, 0.9), (0, 6, 7, 50, 2, 0.9), (0, 10, 5, 40, 0, 0.95),
    (1, 3, 5, 30, 0, 0.8), (1, 4, 3, 20, 2, 0.85), (0, 4, 3, 3, 0, 0.9), (0, 1, 2, 5, 0, 0.9), (0, 12, 3, 30, 0, 0.97)])
def test_gpu_build_equals_oracle(alphabet, k, n_nodes, n_sites, gap_jumps, peak):
    pp, states, oid, goff, glen = D.make_inputs(alphabet, k, n_nodes, n_sites, seed=300 + k + gap_jumps, peak=peak,
                                                gap_rate=0.3 if gap_jumps else 0.0)
    thr = thr_of(alphabet, k)
    o = D.oracle_build(alphabet, k, pp, states, oid, thr, goff, glen, gap_jumps)
    g = dbbuild.build_db(alphabet, k, pp, states, oid, thr, goff, glen, gap_jumps)
    assert g["n_tuples"] == o["n_tuples"]
    D.assert_csr_equal({f: g[f] for f in ("keys", "offsets", "post_node", "post_score")},
                       {f: o[f] for f in ("keys", "offsets", "post_node", "post_score")})


def test_nothing_above_the_threshold_and_errors():
    from rappas_b200._lib import RappasError
    pp, states, oid, _, _ = D.make_inputs(0, 6, 3, 20, seed=1, peak=0.9)
    g = dbbuild.build_db(0, 6, pp, states, oid, 1.0)        # no sum of log10 <= 0 can reach +1
    assert g["n_tuples"] == 0 and g["keys"].size == 0 and list(g["offsets"]) == [0]
    with pytest.raises(RappasError):
        dbbuild.build_db(0, 30, pp, states, oid, -5.0)       # 2*30 + 16 bits do not fit the sort key
    with pytest.raises(RappasError):
        dbbuild.build_db(0, 6, pp, states, oid, -5.0, gap_jumps=1)   # jumps without intervals


def test_build_then_place():
    """The pipeline the row completes: posteriors -> GPU build -> rp_db_load -> placement, every stage equal to
    its oracle.  Reads are cut from the most probable ancestral sequences, so they hit the DB."""
    import rappas_b200 as R
    k, n_nodes, n_sites = 8, 20, 120
    pp, states, oid, _, _ = D.make_inputs(0, k, n_nodes, n_sites, seed=77, peak=0.93)
    oid = np.arange(n_nodes, dtype=np.uint16)
    thr_lin, thr = O.threshold(1.5, 0, k)
    g = dbbuild.build_db(0, k, pp, states, oid, float(thr))
    o = D.oracle_build(0, k, pp, states, oid, float(thr))
    D.assert_csr_equal({f: g[f] for f in ("keys", "offsets", "post_node", "post_score")},
                       {f: o[f] for f in ("keys", "offsets", "post_node", "post_score")})
    rng = np.random.default_rng(5)
    letters = "ATCG"  # state 0..3 = A T C G (DNAStatesShifted)
    reads = []
    for _ in range(200):
        nd, a = int(rng.integers(0, n_nodes)), int(rng.integers(0, n_sites - 60))
        reads.append("".join(letters[int(states[nd, i, 0])] for i in range(a, a + 60)))
    rb = synth.reads_from_strings(reads)
    db = synth.SynthDB(alphabet=0, k=k, n_nodes=n_nodes, thr_lin=thr_lin, thr_log10=thr, keys=g["keys"],
                       offsets=g["offsets"], post_node=g["post_node"], post_score=g["post_score"])
    gdb = R.Database.from_synth(db)
    out = gdb.place(rb)
    assert (out["status"] == 0).mean() > 0.9
    oo = O.OracleDB(db).place(rb)
    parity.assert_placements_equal(out, oo, 7, oo["counts"][:, _abi.CNT_AMBIG] > 0)


@pytest.mark.parametrize("cap", [1, 2000, 9000])
def test_build_in_node_batches(cap, monkeypatch):
    """A build whose tuples do not fit one pass runs in batches of consecutive nodes and merges them on the host
    (RP_DBBUILD_MAX_TUPLES forces it on a small input).  Several ancestral nodes share an original id here, so
    the same (k-mer, node) pair comes back from different batches and the maximum has to be taken again."""
    monkeypatch.setenv("RP_DBBUILD_MAX_TUPLES", str(cap))
    pp, states, oid, goff, glen = D.make_inputs(0, 6, 9, 40, seed=21, peak=0.9, gap_rate=0.3)
    oid = np.asarray([3, 7, 3, 1, 7, 3, 2, 1, 9], np.uint16)
    thr = thr_of(0, 6)
    o = D.oracle_build(0, 6, pp, states, oid, thr, goff, glen, 2)
    g = dbbuild.build_db(0, 6, pp, states, oid, thr, goff, glen, 2)
    assert g["n_tuples"] == o["n_tuples"] > 9000
    D.assert_csr_equal({f: g[f] for f in ("keys", "offsets", "post_node", "post_score")},
                       {f: o[f] for f in ("keys", "offsets", "post_node", "post_score")})
