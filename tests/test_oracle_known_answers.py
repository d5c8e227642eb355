"""Hand-derived known answers that pin the CPU oracle (SURVEY.md 8c list).

The reference has no tests or golden vectors; each expectation below is derived by hand from the
cited Java lines.  CPU only.
"""
import math

import numpy as np
import pytest

import oracle_lib as O
from rappas_b200 import _abi, synth
from rappas_b200.synth import SynthDB, reads_from_strings

f32 = np.float32


def tiny_db(alphabet, k, n_nodes, entries, omega=1.5):
    """entries: {kmer string or state tuple: [(node, score), ...]}"""
    lin, lg = synth.threshold(omega, alphabet, k)
    letters = {"A": 0, "T": 1, "C": 2, "G": 3} if alphabet == 0 else {c: i for i, c in enumerate("RHKDESTNQCGPAILMFWYV")}
    keys, offs, nodes, scores = [], [0], [], []
    for kmer, posts in entries.items():
        st = [letters[c] for c in kmer] if isinstance(kmer, str) else list(kmer)
        assert len(st) == k
        keys.append(O.pack_kmer(alphabet, st))
        for x, v in posts:
            nodes.append(x)
            scores.append(v)
        offs.append(len(nodes))
    return SynthDB(alphabet, k, n_nodes, lin, lg, np.array(keys, np.uint64), np.array(offs, np.uint64),
                   np.array(nodes, np.uint16), np.array(scores, np.float32))


def test_threshold_bits():
    # Main_DBBUILD_3.java:165-166 with omega=1.5
    want = {(0, 8): 0xC05A1893, (0, 10): 0xC0884F5C, (0, 12): 0xC0A3926E, (1, 6): 0xC0D7FCFD}
    for (alpha, k), bits in want.items():
        _, lg = O.threshold(1.5, alpha, k)
        assert int(f32(lg).view(np.uint32)) == bits
        assert synth.threshold(1.5, alpha, k)[1] == lg


def test_compress_mer_example():
    # the word of the commented debug block Main_DBBUILD_3.java:1041 -> bytes {0x8D,0x35}
    assert O.pack_kmer(0, [1, 3, 0, 2, 1, 1, 3, 0]) == 0x358D
    # 9 bases: third byte holds base 8 in its low bits (DNAStatesShifted.java:132-139)
    assert O.pack_kmer(0, [1, 3, 0, 2, 1, 1, 3, 0, 2]) == 0x358D | (2 << 16)
    assert O.pack_kmer(1, [19, 0, 5]) == 19 | (0 << 5) | (5 << 10)


def test_max_ambig_per_mer():
    # AmbigSequenceKnife.java:95 : floor(k^(1/|Sigma|))
    L = O.lib()
    for k in range(3, 16):
        assert L["max_ambig_per_mer"](0, k) == 1
    for k in range(16, 32):
        assert L["max_ambig_per_mer"](0, k) == 2
    assert L["max_ambig_per_mer"](0, 2) == 1
    for k in range(2, 13):
        assert L["max_ambig_per_mer"](1, k) == 1


def test_char_classes():
    L = O.lib()
    for ch, st in zip("ATCG", range(4)):
        assert L["char_class"](0, ord(ch)) == st
        assert L["char_class"](0, ord(ch.lower())) == st
    assert L["char_class"](0, ord("U")) == 1 and L["char_class"](0, ord("u")) == 1
    for ch in "RYSWKMBDHVNryswkmbdhvn.-":
        assert L["char_class"](0, ord(ch)) == -1
    for ch in "XZ*!EFIJLOPQ@ 1":
        assert L["char_class"](0, ord(ch)) == -2
    for i, ch in enumerate("RHKDESTNQCGPAILMFWYV"):
        assert L["char_class"](1, ord(ch)) == i
        assert L["char_class"](1, ord(ch.lower())) == i
    for ch in "-*!XxBbZzJj":
        assert L["char_class"](1, ord(ch)) == -1
    for ch in "UuOo.1":
        assert L["char_class"](1, ord(ch)) == -2
    assert L["char_class"](2, ord("U")) == 9 and L["char_class"](2, ord("o")) == 14
    buf = np.zeros(20, np.uint8)
    # DNAStatesShifted.java:62-96 literal order
    want = {"R": [0, 3], "Y": [2, 1], "S": [2, 3], "W": [0, 1], "K": [3, 1], "M": [0, 2], "B": [2, 3, 1],
            "D": [0, 3, 1], "H": [0, 2, 1], "V": [0, 2, 3], "N": [0, 2, 3, 1], "-": [0, 0, 0, 0], ".": [0, 0, 0, 0]}
    for ch, alts in want.items():
        n = L["ambiguity_equivalence"](0, ord(ch), _abi.ptr(buf))
        assert list(buf[:n]) == alts
    assert L["ambiguity_equivalence"](1, ord("B"), _abi.ptr(buf)) == 2 and list(buf[:2]) == [3, 7]
    assert L["ambiguity_equivalence"](1, ord("Z"), _abi.ptr(buf)) == 2 and list(buf[:2]) == [4, 8]
    assert L["ambiguity_equivalence"](1, ord("J"), _abi.ptr(buf)) == 2 and list(buf[:2]) == [13, 14]
    assert L["ambiguity_equivalence"](1, ord("X"), _abi.ptr(buf)) == 20 and list(buf) == list(range(20))


def test_zero_hits_is_unplaced():
    db = tiny_db(0, 4, 5, {"AAAA": [(1, -0.5)]})
    out = O.OracleDB(db).place(reads_from_strings(["CCCCCCCC"]))
    assert out["status"][0] == _abi.STATUS_UNPLACED and out["n_rows"][0] == 0
    assert list(out["counts"][0]) == [5, 0, 0, 0]
    assert out["node"][0, 0] == 0xFFFF and out["score"][0, 0] == -np.inf


def test_length_edge_cases():
    db = tiny_db(0, 4, 5, {"AAAA": [(1, -0.5)]})
    out = O.OracleDB(db).place(reads_from_strings(["AAA", "AA", "", "AAAA", "AAZA", "Z"]))
    # len == k-1 -> Q = 0 -> unplaced ; len < k-1 -> reference throws (AmbigSequenceKnife.java:145)
    assert list(out["status"]) == [1, 2, 2, 0, 3, 3]
    assert list(out["counts"][0]) == [0, 0, 0, 0]
    assert out["n_rows"][3] == 1


def test_single_hit_read():
    # one node, score = v + (Q-1)*T in the f32 order of PlacementProcess.java:728,733 ; lwr = 1 exactly
    db = tiny_db(0, 4, 5, {"ACGT": [(3, -0.25)]})
    T = db.thr_log10
    out = O.OracleDB(db).place(reads_from_strings(["CCACGTCC"]))
    Q = 5
    want = f32(f32(f32(0) + f32(Q) * T) + f32(f32(-0.25) - T))
    assert out["status"][0] == 0 and out["n_rows"][0] == 1
    assert out["node"][0, 0] == 3 and out["score"][0, 0] == want and out["lwr"][0, 0] == 1.0
    assert list(out["counts"][0]) == [5, 1, 0, 0]
    assert abs(float(want) - (-0.25 + (Q - 1) * float(T))) < 1e-5


def test_score_identity_vs_f64():
    # S[x] = sum_hit v + (Q - C[x]) * T  (SURVEY 8c), 1e-4 relative against an f64 evaluation
    w = synth.workload(1, scale=0.2)
    db, rb = synth.build(w, n_reads=50)
    odb = O.OracleDB(db)
    S, Cn = odb.node_scores(rb)
    ex = odb.extract(rb)
    key_index = {int(c): i for i, c in enumerate(db.keys)}
    T = float(db.thr_log10)
    for r in range(rb.n_reads):
        Q = int(ex["win_off"][r + 1] - ex["win_off"][r])
        ref = np.zeros(db.n_nodes)
        cnt = np.zeros(db.n_nodes, np.int64)
        for j in range(Q):
            code = int(ex["code"][int(ex["win_off"][r]) + j])
            if code in key_index:
                i = key_index[code]
                lo, hi = int(db.offsets[i]), int(db.offsets[i + 1])
                ref[db.post_node[lo:hi]] += db.post_score[lo:hi].astype(np.float64)
                cnt[db.post_node[lo:hi]] += 1
        touched = cnt > 0
        assert np.array_equal(touched, ~np.isnan(S[r]))
        assert np.array_equal(cnt, Cn[r])
        want = ref + (Q - cnt) * T
        np.testing.assert_allclose(S[r][touched], want[touched], rtol=1e-4)


def test_gap_behaves_as_A_in_nucleotide_reads():
    # DNAStatesShifted.java:57-58: '-' and '.' -> [A,A,A,A]; the mean path then equals the plain-A
    # contribution up to the f64 round trip 10^v -> log10
    db = tiny_db(0, 4, 6, {"ACGT": [(2, -0.75), (4, -1.5)], "CGTA": [(2, -0.5)]})
    odb = O.OracleDB(db)
    a = odb.place(reads_from_strings(["TTACGTATT"]))
    for gap in "-.":
        g = odb.place(reads_from_strings(["TT" + gap + "CGTATT"]))
        assert g["status"][0] == 0
        assert list(g["node"][0, :2]) == list(a["node"][0, :2])
        np.testing.assert_allclose(g["score"][0, :2], a["score"][0, :2], rtol=1e-6)
        # windows 0..2 contain the gap -> 3 ambiguous windows treated, only the plain CGTA match counted
        assert list(g["counts"][0]) == [6, 1, 3, 0]
    assert list(a["counts"][0]) == [6, 2, 0, 0]


def test_ambiguity_mean_and_max_by_hand():
    # window "ACRT": R -> [A, G] (DNAStatesShifted.java:62-63) -> alternatives ACAT, ACGT
    db = tiny_db(0, 4, 4, {"ACAT": [(1, -1.0)], "ACGT": [(1, -2.0), (2, -0.5)]})
    T, Tlin = db.thr_log10, db.thr_lin
    odb = O.OracleDB(db)
    rd = reads_from_strings(["ACRT"])
    Q = 1
    out = odb.place(rd)
    assert list(out["counts"][0]) == [1, 0, 1, 0]  # matched NOT counted in the ambiguity path (quirk 7)
    S, _ = odb.node_scores(rd)
    # node 1: both alternatives hit
    samb = f32(float(f32(0)) + math.pow(10.0, -1.0))
    samb = f32(float(samb) + math.pow(10.0, -2.0))
    avg = f32(f32(samb + f32(f32(0) * Tlin)) / f32(2))
    s1 = f32(float(f32(Q) * T) + (math.log10(float(avg)) - float(T)))
    # node 2: one of two alternatives, the other padded with T_lin
    samb2 = f32(0.0 + math.pow(10.0, -0.5))
    avg2 = f32(f32(samb2 + f32(f32(1) * Tlin)) / f32(2))
    s2 = f32(float(f32(Q) * T) + (math.log10(float(avg2)) - float(T)))
    assert S[0, 1] == s1 and S[0, 2] == s2
    assert np.isnan(S[0, 0]) and np.isnan(S[0, 3])
    Smax, _ = odb.node_scores(rd, _abi.place_cfg(amb_with_max=True))
    assert Smax[0, 1] == f32(f32(Q) * T + f32(f32(-1.0) - T))
    assert Smax[0, 2] == f32(f32(Q) * T + f32(f32(-0.5) - T))
    no = odb.place(rd, _abi.place_cfg(treat_amb=False))
    assert no["status"][0] == 1 and list(no["counts"][0]) == [1, 0, 0, 1]


def test_too_many_ambiguities_skips_window():
    db = tiny_db(0, 4, 4, {"ACGT": [(1, -1.0)]})
    out = O.OracleDB(db).place(reads_from_strings(["NNCGTACGT"]))
    # windows 0,1 hold 2 and ... ambiguities: w0 "NNCG" (2) skipped, w1 "NCGT" (1) treated, rest plain
    assert list(out["counts"][0]) == [6, 1, 1, 1]


def test_two_ambiguities_non_cartesian_quirk():
    # k=16 -> maxAmbigPerMer = 2; two 2-way codes give alternatives (a0,b0),(a1,b1),(a0,b0),(a1,b1)
    # (AmbigSequenceKnife.java:249-256)
    k = 16
    base = "ACGTACGTACGTAC"
    word = "R" + base + "Y"  # R=[A,G] Y=[C,T]
    e = {"A" + base + "C": [(0, -1.0)], "G" + base + "T": [(1, -1.0)], "A" + base + "T": [(2, -1.0)],
         "G" + base + "C": [(3, -1.0)]}
    db = tiny_db(0, k, 4, e)
    odb = O.OracleDB(db)
    ex = odb.extract(reads_from_strings([word]))
    assert ex["kind"][0] == _abi.WIN_AMBIG and ex["nalt"][0] == 4
    S, Cn = odb.node_scores(reads_from_strings([word]))
    assert not np.isnan(S[0, 0]) and not np.isnan(S[0, 1])
    assert np.isnan(S[0, 2]) and np.isnan(S[0, 3])  # the cross combinations are never enumerated


def test_lwr_shift_and_keep_factor():
    # two nodes, scores far below -308 -> shift branch; lwr = 10^(s-best)/sum
    posts = {"ACGT": [(1, -0.5), (2, -1.5)]}
    db = tiny_db(0, 4, 4, posts)
    read = "ACGT" + "C" * 200  # Q = 201, Q*T ~ -342
    out = O.OracleDB(db).place(reads_from_strings([read]), _abi.place_cfg(keep_factor=0.01))
    assert out["n_rows"][0] == 2 and list(out["node"][0, :2]) == [1, 2]
    s = out["score"][0]
    assert s[0] < -308
    d = float(f32(s[1] - s[0]))
    tot = 1.0 + math.pow(10.0, d)
    assert out["lwr"][0, 0] == pytest.approx(1.0 / tot, rel=1e-12)
    assert out["lwr"][0, 1] == pytest.approx(math.pow(10.0, float(s[1]) - float(s[0])) / tot, rel=1e-12)
    # keep_factor 0.5: second row (ratio 0.1 of best) is cut, but still counted in the sum
    out2 = O.OracleDB(db).place(reads_from_strings([read]), _abi.place_cfg(keep_factor=0.5))
    assert out2["n_rows"][0] == 1 and out2["lwr"][0, 0] == out["lwr"][0, 0]
    # ns_bound above the best score suppresses all rows but the read still counts as placed
    out3 = O.OracleDB(db).place(reads_from_strings([read]), _abi.place_cfg(ns_bound=-10.0))
    assert out3["n_rows"][0] == 0 and out3["status"][0] == 0


def test_keep_at_most_limits_rows_and_sum():
    posts = {"ACGT": [(i, -0.1 * (i + 1)) for i in range(10)]}
    db = tiny_db(0, 4, 12, posts)
    for K in (1, 3, 7, 12):
        out = O.OracleDB(db).place(reads_from_strings(["ACGTT"]), _abi.place_cfg(keep_at_most=K, keep_factor=0.0))
        nb = min(K, 10)
        assert out["n_rows"][0] == nb
        assert list(out["node"][0, :nb]) == list(range(nb))
        assert out["lwr"][0, :nb].sum() == pytest.approx(1.0, rel=1e-12)


def test_threaded_oracle_equals_serial():
    db, rb = synth.build(synth.workload(1, scale=0.1), n_reads=300)
    odb = O.OracleDB(db)
    a, b = odb.place(rb), odb.place(rb, threads=4)
    for key in a:
        assert np.array_equal(a[key], b[key], equal_nan=True)
