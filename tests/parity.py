"""Comparison helpers: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): k-mer codes, hit lists and match counts bit-exact; per-node scores
within 1e-4 relative -- in fact the CUDA path accumulates every S[x] in the reference's own f32 order,
so plain-window scores are compared for BIT equality and only ambiguity-path scores (f64 pow/log10
from two different libms) get a tolerance; best-placement node identical except at exact ties.
"""
import numpy as np

SCORE_RTOL_AMBIG = 1e-6   # ambiguity path: f32 exp10f / log10f on the GPU vs f64 pow / log10 narrowed to f32 in the
                          # reference (<= ~1e-6 absolute per ambiguous window, i.e. an ulp or two of the f32 score)
SCORE_ATOL_AMBIG = 4e-6   # ... which is an ABSOLUTE error: it only shows against scores of magnitude < 4 (one-window reads)
LWR_RTOL = 1e-9           # f64 pow + a different summation order of <= keep_at_most terms


def lwr_rtol_ambig(best_score):
    """like_weight_ratio = 10^(score - best) / sum: an ulp of an f32 score of magnitude |S| (6e-8 |S|) moves it by
    2.3 * 6e-8 |S| relative; a few ulps are allowed on reads that took the ambiguity path"""
    return 1e-6 * max(1.0, abs(float(best_score)))


def assert_extract_equal(g, o):
    for key in ("status", "kind", "nalt", "code", "hits"):
        assert np.array_equal(g[key], o[key]), key


def assert_scores_equal(Sg, So, ambiguous_reads=None):
    """Sg/So: [n][N] with NaN = untouched.  Bit-exact on reads without treated ambiguous windows."""
    assert np.array_equal(np.isnan(Sg), np.isnan(So)), "touched sets differ"
    n = Sg.shape[0]
    amb = np.zeros(n, bool) if ambiguous_reads is None else np.asarray(ambiguous_reads, bool)
    plain = ~amb
    a, b = Sg[plain].view(np.uint32), So[plain].view(np.uint32)
    nanmask = np.isnan(So[plain])
    assert np.array_equal(a[~nanmask], b[~nanmask]), "plain-read scores are not bit-identical"
    if amb.any():
        x, y = Sg[amb], So[amb]
        m = ~np.isnan(y)
        np.testing.assert_allclose(x[m], y[m], rtol=SCORE_RTOL_AMBIG, atol=SCORE_ATOL_AMBIG)
    return int(amb.sum())


def assert_placements_equal(g, o, K, ambiguous_reads=None, So=None):
    """So: the oracle's per-node score vectors [n][N] (NaN = untouched).  With them a node that differs from the
    oracle's at some row must carry, in the ORACLE's own vector, the score of that row (bit-equal; within
    SCORE_RTOL_AMBIG on reads that took the f64 ambiguity path): i.e. the two only permute tied nodes.  Without
    them a differing node is accepted where the row's score is tied with another listed score, or on the last
    kept row (its tie partner may be the first node that was not kept)."""
    n = o["status"].shape[0]
    assert np.array_equal(g["status"], o["status"])
    if g.get("counts") is not None:
        assert np.array_equal(g["counts"], o["counts"])
    amb = np.zeros(n, bool) if ambiguous_reads is None else np.asarray(ambiguous_reads, bool)
    ties = 0
    for r in range(n):
        nr = int(o["n_rows"][r])
        so, sg = o["score"][r], g["score"][r]
        if not amb[r]:
            assert int(g["n_rows"][r]) == nr, r
            # scores (sorted descending) must be bit-identical even if tied nodes are permuted
            assert np.array_equal(sg[:nr].view(np.uint32), so[:nr].view(np.uint32)), r
        else:
            if int(g["n_rows"][r]) != nr:
                # a row at the keep_factor boundary may flip with a 1-ulp score difference
                assert abs(int(g["n_rows"][r]) - nr) <= 1, r
                nr = min(nr, int(g["n_rows"][r]))
            np.testing.assert_allclose(sg[:nr], so[:nr], rtol=SCORE_RTOL_AMBIG, atol=SCORE_ATOL_AMBIG)
        no, ng = o["node"][r, :nr], g["node"][r, :nr]
        if not np.array_equal(no, ng):
            for i in np.nonzero(no != ng)[0]:
                if So is not None:
                    ref = So[r, int(ng[i])]  # what the oracle itself scored the GPU's node
                    if amb[r]:
                        tied = bool(np.isclose(ref, so[i], rtol=2 * SCORE_RTOL_AMBIG, atol=2 * SCORE_ATOL_AMBIG))
                    else:
                        tied = np.float32(ref).view(np.uint32) == np.float32(so[i]).view(np.uint32)
                else:
                    if amb[r]:
                        others = np.delete(so[:K], i)
                        tied = bool(np.isclose(others, so[i], rtol=2 * SCORE_RTOL_AMBIG, atol=2 * SCORE_ATOL_AMBIG).any()) or (i == nr - 1)
                    else:
                        tied = (np.sum(so[:K] == so[i]) > 1) or (i == nr - 1)
                assert tied, (r, i, no, ng, so[:nr])
            ties += 1
        if not (amb[r] and not np.array_equal(no, ng)):  # (near-tied nodes that swapped places carry each other's ratios)
            np.testing.assert_allclose(g["lwr"][r, :nr], o["lwr"][r, :nr],
                                       rtol=lwr_rtol_ambig(so[0]) if amb[r] and nr else LWR_RTOL)
        # unused slots carry the documented fill values
        nrg = int(g["n_rows"][r])
        assert (g["node"][r, nrg:] == 0xFFFF).all() and np.isneginf(g["score"][r, nrg:]).all()
        assert (g["lwr"][r, nrg:] == 0).all()
    return ties
