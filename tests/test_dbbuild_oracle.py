"""DB-build row, CPU side: the recursive C oracle against the literal Python restatement and
hand-derived answers, and the product's explorer state machine (compiled for the host) against the oracle."""
import numpy as np
import pytest

import dbbuild_lib as D
from oracle import dbbuild_py


def py_csr(alphabet, k, pp, states, oid, thr, gaps=None, gap_jumps=0):
    table, n = dbbuild_py.build(alphabet, k, pp, states, oid, thr, gaps, gap_jumps)
    codes, nodes, scores = [], [], []
    for c in sorted(table):
        for nd in sorted(table[c]):
            codes.append(c); nodes.append(nd); scores.append(table[c][nd])
    return D.csr_from_tuples(np.asarray(codes, np.uint64), np.asarray(nodes, np.uint16), np.asarray(scores, np.float32)), n


def gaps_as_lists(gap_off, gap_len, n_sites):
    if gap_off is None:
        return None
    return [list(map(int, gap_len[int(gap_off[i]):int(gap_off[i + 1])])) or None for i in range(n_sites)]


def test_hand_derived_single_site_chain():
    """k = 2, 2 sites, one node, threshold -1: the words are the pairs whose summed log10 stays >= -1."""
    pp = np.log10(np.array([[[0.7, 0.2, 0.06, 0.04], [0.9, 0.05, 0.03, 0.02]]], np.float64)).astype(np.float32)
    states = np.array([[[3, 0, 2, 1], [1, 2, 0, 3]]], np.uint8)
    oid = np.array([5], np.uint16)
    o = D.oracle_build(0, 2, pp, states, oid, -1.0)
    # 0.7*0.9 = 0.63 and 0.2*0.9 = 0.18 are the only products >= 0.1; code = b0 + 4*b1
    assert list(o["keys"]) == [0 + 4 * 1, 3 + 4 * 1] and list(o["post_node"]) == [5, 5]
    exp = [np.float32(np.float32(pp[0, 0, 1]) + np.float32(pp[0, 1, 0])), np.float32(np.float32(pp[0, 0, 0]) + np.float32(pp[0, 1, 0]))]
    assert [x.view(np.uint32) for x in o["post_score"]] == [x.view(np.uint32) for x in exp]
    assert o["n_tuples"] == 2


@pytest.mark.parametrize("alphabet,k,n_nodes,n_sites,gap_jumps", [(0, 3, 3, 9, 0), (0, 4, 2, 10, 1), (0, 4, 2, 10, 2),
                                                               (1, 2, 2, 6, 0), (1, 3, 1, 5, 2)])
def test_c_oracle_equals_python_restatement(alphabet, k, n_nodes, n_sites, gap_jumps):
    pp, states, oid, goff, glen = D.make_inputs(alphabet, k, n_nodes, n_sites, seed=k * 7 + gap_jumps, peak=0.8,
                                                gap_rate=0.4 if gap_jumps else 0.0)
    thr = float(np.float32(np.log10(np.float32((1.5 / (4 if alphabet == 0 else 20)) ** k))))
    o = D.oracle_build(alphabet, k, pp, states, oid, thr, goff, glen, gap_jumps)
    p, n = py_csr(alphabet, k, pp, states, oid, thr, gaps_as_lists(goff, glen, n_sites), gap_jumps)
    assert o["n_tuples"] == n and n > 0
    D.assert_csr_equal({f: o[f] for f in p}, p)


@pytest.mark.parametrize("alphabet,k,n_nodes,n_sites,gap_jumps,peak", [
    (0, 5, 6, 40, 0, 0.9), (0, 8, 4, 30, 0, 0.93), (0, 6, 5, 30, 1, 0.9), (0, 6, 5, 30, 2, 0.9), (0, 10, 2, 24, 0, 0.95),
    (1, 3, 3, 20, 0, 0.8), (1, 4, 2, 14, 2, 0.85), (0, 4, 3, 3, 0, 0.9), (0, 1, 2, 5, 0, 0.9)])
def test_product_explorer_equals_oracle(alphabet, k, n_nodes, n_sites, gap_jumps, peak):
    """Same tuples (count), same maxima bit for bit: the f32 running sum makes the visiting order part of the
    result, so this pins the state machine to the recursion."""
    pp, states, oid, goff, glen = D.make_inputs(alphabet, k, n_nodes, n_sites, seed=100 + k + gap_jumps, peak=peak,
                                                gap_rate=0.3 if gap_jumps else 0.0)
    thr = float(np.float32(np.log10(np.float32((1.5 / (4 if alphabet == 0 else 20)) ** k))))
    o = D.oracle_build(alphabet, k, pp, states, oid, thr, goff, glen, gap_jumps)
    codes, nodes, scores = D.core_tuples(alphabet, k, pp, states, oid, thr, goff, glen, gap_jumps)
    assert codes.size == o["n_tuples"]
    D.assert_csr_equal(D.csr_from_tuples(codes, nodes, scores), {f: o[f] for f in ("keys", "offsets", "post_node", "post_score")})


def test_pp_prepare_equals_the_restatement():
    """PHYMLWrapper.java:206-229: clamp, (float)Math.log10, Collections.sort (stable, SiteProba.compareTo)."""
    import functools
    from rappas_b200 import dbbuild
    rng = np.random.default_rng(3)
    for ns, soc in ((4, [0, 2, 3, 1]), (20, list(range(20)))):       # PhyML prints A C G T -> states 0 2 3 1
        probs = rng.dirichlet(np.ones(ns) * 0.3, (5, 37)).astype(np.float32)
        probs[0, 0] = 0.25 if ns == 4 else 0.05                        # ties everywhere: column order must survive
        probs[1, 3, :2] = probs[1, 3, 0]                               # a tie for first place
        probs[2, 5, 1] = 0.0                                           # clamped to Float.MIN_VALUE before the log
        pp, st = dbbuild.prepare_posteriors(probs, soc)

        def cmp(a, b):                                                 # SiteProba.compareTo
            d = np.float32(a[0]) - np.float32(b[0])
            return 1 if d < 0.0 else (-1 if d > 0.0 else 0)
        for nd in range(probs.shape[0]):
            for site in range(probs.shape[1]):
                row = []
                for i in range(ns):
                    v = np.float32(max(probs[nd, site, i], np.float32(1.4e-45)))
                    row.append((np.float32(np.log10(np.float64(v))), soc[i]))
                row = sorted(row, key=functools.cmp_to_key(cmp))       # Python's sort is stable, like Collections.sort
                assert [x[0].view(np.uint32) for x in row] == [x.view(np.uint32) for x in pp[nd, site]], (nd, site)
                assert [x[1] for x in row] == list(st[nd, site]), (nd, site)
        assert np.all(np.diff(pp, axis=2) <= 0)


def test_gap_intervals_equal_the_restatement():
    """Alignment.updateGapIntervals (alignement/Alignment.java:231-260), transliterated."""
    from rappas_b200 import dbbuild

    def restatement(charMatrix):
        gapIntervals = [None] * len(charMatrix[0])
        for i in range(len(charMatrix)):
            firstGapIndex, previousChar = -1, 'n'
            for j in range(len(charMatrix[i])):
                c = charMatrix[i][j]
                if c == '-':
                    if previousChar != '-':
                        if firstGapIndex == -1:
                            firstGapIndex = j
                else:
                    if firstGapIndex != -1:
                        if gapIntervals[firstGapIndex] is None:
                            gapIntervals[firstGapIndex] = []
                        length = j - firstGapIndex
                        if length not in gapIntervals[firstGapIndex]:
                            gapIntervals[firstGapIndex].append(length)
                        firstGapIndex = -1
                previousChar = c
        return gapIntervals

    rows = ["AC--GT-A", "AC-TGT--", "--CTG--A", "ACGTGTAA", "AC---T-A", "-C--GT-A"]   # trailing runs are not registered
    off, lens = dbbuild.gap_intervals(rows)
    exp = restatement(rows)
    got = [list(map(int, lens[int(off[j]):int(off[j + 1])])) or None for j in range(len(rows[0]))]
    assert got == exp and exp[2] == [2, 1, 3] and exp[6] == [1] and exp[0] == [2, 1]
    rng = np.random.default_rng(0)
    for _ in range(20):
        rows = ["".join(rng.choice(list("ACGT-"), 40, p=[.2, .2, .2, .2, .2])) for _ in range(int(rng.integers(1, 12)))]
        off, lens = dbbuild.gap_intervals(rows)
        got = [list(map(int, lens[int(off[j]):int(off[j + 1])])) or None for j in range(40)]
        assert got == restatement(rows)


@pytest.mark.parametrize("n_batches", [1, 2, 3, 5, 8])
def test_batch_merge_applies_the_maximum_again(n_batches):
    """The product's host merge of per-batch results (rp_dbbuild_merge.h, compiled for the host): batches of
    sorted distinct (code << 16 | node, score) pairs that share pair keys -> the CSR of the maximum."""
    import ctypes as C
    import os
    import subprocess
    from rappas_b200 import _abi
    so = os.path.join(D.ROOT, "tests", "helpers", "dbbuild_merge_host.so")
    src = os.path.join(D.ROOT, "tests", "helpers", "dbbuild_merge_host.cpp")
    hdr = os.path.join(D.ROOT, "rappas_b200", "csrc", "rp_dbbuild_merge.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    rng = np.random.default_rng(n_batches)
    keys, scores, bounds = [], [], [0]
    for b in range(n_batches):
        n = int(rng.integers(0, 400))
        code = rng.integers(0, 60, n).astype(np.uint64)
        node = rng.integers(0, 12, n).astype(np.uint64)
        pk = np.unique((code << np.uint64(16)) | node)           # sorted, distinct within the batch
        keys.append(pk)
        scores.append((-rng.random(pk.size) * 5).astype(np.float32))
        bounds.append(bounds[-1] + pk.size)
    key = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    score = np.concatenate(scores) if scores else np.zeros(0, np.float32)
    bounds = np.asarray(bounds, np.uint64)
    n = max(1, key.size)
    ok, oo, on, os_ = np.zeros(n, np.uint64), np.zeros(n + 1, np.uint64), np.zeros(n, np.uint16), np.zeros(n, np.float32)
    nk, npost = C.c_uint64(), C.c_uint64()
    lib.merge_host(_abi.ptr(key) if key.size else None, _abi.ptr(score) if key.size else None, _abi.ptr(bounds), n_batches,
                   C.byref(nk), C.byref(npost), _abi.ptr(ok), _abi.ptr(oo), _abi.ptr(on), _abi.ptr(os_))
    got = dict(keys=ok[:nk.value], offsets=oo[:nk.value + 1], post_node=on[:npost.value], post_score=os_[:npost.value])
    exp = D.csr_from_tuples(key >> np.uint64(16), (key & np.uint64(0xFFFF)).astype(np.uint16), score)
    D.assert_csr_equal(got, exp)
