"""FASTA ingest + duplicate structure + .jplace writer (C++ in librappas_b200.so, no GPU work) against
the literal Python restatement in tests/ref_host.py; placement rows come from the CPU oracle."""
import json
import re

import numpy as np
import pytest

import oracle_lib as O
import ref_host
from rappas_b200 import _abi, ingest, synth

TRICKY = (
    "# a comment line\n\n>r1 first read\nACGT\nACGTAC\r\n\n>r2\n  ACGTTT  \n>r3 dup of r1 with gaps\nAC-GT\nACG-TAC\n"
    ">r4 exact dup of r1\nACGTACGTAC\n>r5\n#not a sequence line\nNNNN\rACGT\n>r6 empty\n>r7 last, no newline\nacgtRYK")


def test_fasta_records_match_the_restatement():
    q = ingest.QueryFile.from_text(TRICKY)
    exp = ref_host.read_fasta(TRICKY)
    assert q.n_records == len(exp) == 7
    assert q.headers == [h for h, _ in exp]
    seqs = [q.unique.read(int(u)) for u in q.unique_of]
    assert seqs == [s for _, s in exp]
    assert seqs[0] == "ACGTACGTAC" and seqs[1] == "ACGTTT" and seqs[5] == "" and seqs[6] == "acgtRYK"
    # exact duplicates share a unique sequence; gapped variants share only the duplicate group
    assert q.unique_of[0] == q.unique_of[3] != q.unique_of[2]
    assert q.group_of[0] == q.group_of[2] == q.group_of[3]
    assert q.n_unique == 6 and q.n_groups == 5


def test_fasta_errors():
    from rappas_b200._lib import RappasError
    with pytest.raises(RappasError):
        ingest.QueryFile.from_text("ACGT\n>late header\nACGT\n")
    with pytest.raises(RappasError):
        ingest.QueryFile.from_text("\n# nothing\n")
    with pytest.raises(RappasError):
        ingest.QueryFile.from_file("/nonexistent/q.fa")


def test_java_number_layout():
    known = {(1e-5, False): "1.0E-5", (100.0, False): "100.0", (1e7, False): "1.0E7", (0.001, False): "0.001",
             (9999999.0, False): "9999999.0", (0.5, True): "0.5", (-487.12344, True): "-487.12344",
             (1.0, False): "1.0", (123456.789, False): "123456.789", (1.0 / 3, False): "0.3333333333333333",
             (0.0, False): "0.0", (float("inf"), False): "null", (3.4e38, True): "3.4E38", (1.17549435e-38, True): "1.1754944E-38"}
    for (v, f), s in known.items():
        assert ingest.java_number(v, f) == s == ref_host.java_number(v, f), (v, f)
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 1, 300), 10.0 ** rng.uniform(-12, 12, 300) * rng.choice([-1, 1], 300),
                           -rng.uniform(0, 900, 300)])
    for v in vals:
        assert ingest.java_number(v) == ref_host.java_number(v)
        assert ingest.java_number(v, True) == ref_host.java_number(v, True)
        assert float(ingest.java_number(v)) == v and np.float32(ingest.java_number(v, True)) == np.float32(v)


@pytest.mark.parametrize("guppy", [False, True])
def test_jplace_equals_restatement(tmp_path, guppy):
    db = synth.make_db(0, 6, 41, n_keys=2500, mean_postings=5, seed=0)
    rb = synth.make_reads(db, 60, (5, 90), seed=100, n_rate=0.01)
    rng = np.random.default_rng(1)
    reads = [rb.read(i) for i in range(rb.n_reads)]
    reads += ["TTTTTTTT" * 2, "T" * 5]  # probably unplaced; too short would abort the reference, so len >= k-1
    lines = []
    for i, s in enumerate(reads):
        lines.append(">q%d some description %d" % (i, i))
        lines.append(s)
    for j in range(25):  # exact duplicates, gapped duplicates (gap = extra A for the placement, same checksum)
        i = int(rng.integers(0, len(reads)))
        s = reads[i]
        if j % 3 == 0 and len(s) > 8:
            cut = int(rng.integers(1, len(s) - 1))
            s = s[:cut] + "-" + s[cut:]
        lines.append(">dup%d of q%d" % (j, i))
        lines.append(s)
    text = "\n".join(lines) + "\n"
    q = ingest.QueryFile.from_text(text)
    odb = O.OracleDB(db)
    cfg = _abi.place_cfg()
    res = odb.place(q.unique, cfg)
    edge_id = rng.permutation(db.n_nodes).astype(np.int32)
    branch = rng.uniform(1e-4, 2.0, db.n_nodes).astype(np.float32)
    out, npl = tmp_path / "out.jplace", tmp_path / "not_placed.txt"
    n = q.write_jplace(out, res, cfg.keep_at_most, edge_id, branch, tree_newick="(A:1{0},B:2{1}){2};",
                       invocation="test", guppy_compat=guppy, not_placed_path=npl)
    raw = out.read_text()
    doc = json.loads(raw)

    def place_one(seq):
        one = odb.place(synth.reads_from_strings([seq]), cfg)
        k = int(one["n_rows"][0])
        return int(one["status"][0]), [(int(one["node"][0, i]), one["score"][0, i], one["lwr"][0, i]) for i in range(k)]

    exp, exp_np = ref_host.build_jplace(ref_host.read_fasta(text), place_one, edge_id, branch, guppy)
    assert n == len(exp) == len(doc["placements"]) > 30
    assert doc["version"] == 3 and doc["metadata"] == {"invocation": "test"} and doc["tree"] == "(A:1{0},B:2{1}){2};"
    assert doc["fields"] == (["distal_length", "edge_num", "like_weight_ratio", "likelihood", "pendant_length"] if guppy
                             else ["edge_num", "likelihood", "like_weight_ratio", "distal_length", "pendant_length"])
    # numbers: compare the printed tokens (Java formatting) and the parsed values
    got_rows = re.findall(r"\[([^\[\]\"]+)\]", raw.split('"placements"')[1].split('"version"')[0])
    exp_rows = [",".join(str(c) for c in row) for pl in exp for row in pl["p"]]
    assert got_rows == exp_rows
    for g, e in zip(doc["placements"], exp):
        assert g["nm"] == e["nm"]
        assert len(g["p"]) == len(e["p"])
    assert any(len(pl["nm"]) > 1 for pl in doc["placements"])
    assert npl.read_text().splitlines() == exp_np


def test_writer_refuses_reads_the_reference_aborts_on(tmp_path):
    from rappas_b200._lib import RappasError
    db = synth.make_db(0, 6, 41, n_keys=2500, mean_postings=5, seed=0)
    q = ingest.QueryFile.from_text(">ok\nACGTACGTACGT\n>bad\nACGTZZACGTAA\n")
    res = O.OracleDB(db).place(q.unique, _abi.place_cfg())
    with pytest.raises(RappasError):
        q.write_jplace(tmp_path / "x.jplace", res, 7, np.arange(41), np.ones(41))


@pytest.mark.parametrize("threads", [1, 2, 3, 7])
def test_parallel_ingest_equals_the_restatement(threads, monkeypatch):
    """The text is cut at header lines into one piece per thread (RP_HOST_THREADS forces the count on a small
    input): records, unique sequences and duplicate groups must not depend on where the cuts fall."""
    monkeypatch.setenv("RP_HOST_THREADS", str(threads))
    rng = np.random.default_rng(threads)
    base = ["ACGT" * int(rng.integers(1, 12)) + "ACGTN"[int(rng.integers(0, 5))] for _ in range(40)]
    parts = ["# leading comment\r\n"]
    for i in range(400):
        s = base[int(rng.integers(0, len(base)))]
        if i % 7 == 3:
            cut = int(rng.integers(1, len(s)))
            s = s[:cut] + "-" + s[cut:]                      # gapped duplicate: own unique, shared group
        eol = ["\n", "\r\n", "\r"][int(rng.integers(0, 3))]
        hdr = ">r%d some > text %d" % (i, i)                   # '>' inside a header is not a record start
        if i % 11 == 5:
            half = len(s) // 2
            body = s[:half] + eol + eol + "#comment > here" + eol + s[half:]   # wrapped, blank and '#' lines inside
        else:
            body = s
        parts.append(hdr + eol + body + eol)
    text = "".join(parts)
    q = ingest.QueryFile.from_text(text)
    exp = ref_host.read_fasta(text)
    assert q.n_records == len(exp) == 400
    assert q.headers == [h for h, _ in exp]
    assert [q.unique.read(int(u)) for u in q.unique_of] == [s for _, s in exp]
    # unique ids in order of first appearance of the exact sequence; groups by the gap-stripped sequence
    first, groups = {}, {}
    for r, (_, s) in enumerate(exp):
        assert q.unique_of[r] == first.setdefault(s, len(first))
        assert q.group_of[r] == groups.setdefault(s.replace("-", ""), len(groups))
    assert q.n_unique == len(first) and q.n_groups == len(groups)


@pytest.mark.parametrize("threads", [1, 3])
def test_block_parallel_writer_equals_the_restatement(tmp_path, threads, monkeypatch):
    """More placements than one formatting block (8 192), forced onto 1 and 3 host threads: the document must
    be the restatement's, placement by placement, whatever thread formatted which block."""
    monkeypatch.setenv("RP_HOST_THREADS", str(threads))
    rng = np.random.default_rng(7)
    n_unique, n_dup, n_nodes, K = 20000, 3000, 101, 7
    alphabet = np.frombuffer(b"ACGT", np.uint8)
    seqs = ["".join(map(chr, alphabet[rng.integers(0, 4, 24)])) + "%05d" % i for i in range(n_unique)]  # distinct
    order = list(range(n_unique)) + [int(x) for x in rng.integers(0, n_unique, n_dup)]
    rng.shuffle(order)
    text = "".join(">q%d d%d\n%s\n" % (j, j, seqs[i]) for j, i in enumerate(order))
    q = ingest.QueryFile.from_text(text)
    assert q.n_unique == n_unique
    n_rows = rng.integers(0, 4, n_unique).astype(np.int32)      # 0 rows = below nsBound: not written, not registered
    status = np.where(rng.random(n_unique) < 0.02, 1, 0).astype(np.int32)  # a few reads without any hit
    n_rows[status == 1] = 0
    res = dict(n_rows=n_rows, node=rng.integers(0, n_nodes, (n_unique, K)).astype(np.uint16),
               score=(-rng.random((n_unique, K)) * 700).astype(np.float32), lwr=rng.random((n_unique, K)), status=status)
    edge_id = rng.permutation(n_nodes).astype(np.int32)
    branch = rng.uniform(1e-4, 2.0, n_nodes).astype(np.float32)
    out, npl = tmp_path / "o.jplace", tmp_path / "np.txt"
    n = q.write_jplace(out, res, K, edge_id, branch, tree_newick="(a,b);", invocation="t", not_placed_path=npl)
    by_seq = {}
    for u in range(n_unique):
        by_seq[q.unique.read(u)] = u

    def place_one(seq):
        u = by_seq[seq]
        return int(status[u]), [(int(res["node"][u, i]), res["score"][u, i], res["lwr"][u, i]) for i in range(int(n_rows[u]))]

    exp, exp_np = ref_host.build_jplace(ref_host.read_fasta(text), place_one, edge_id, branch, False)
    doc = json.loads(out.read_text())
    assert n == len(exp) == len(doc["placements"]) > 8192
    raw = out.read_text()
    got_rows = re.findall(r"\[([^\[\]\"]+)\]", raw.split('"placements"')[1].split('"version"')[0])
    assert got_rows == [",".join(str(c) for c in row) for pl in exp for row in pl["p"]]
    assert [pl["nm"] for pl in doc["placements"]] == [pl["nm"] for pl in exp]
    assert npl.read_text().splitlines() == exp_np


def test_ingest_fuzz_against_the_restatement(monkeypatch):
    """Random FASTA-ish text over a small alphabet of troublesome bytes, cut into 1-5 pieces: the native parser
    must agree with the restatement on every record, or fail where the restatement fails."""
    from rappas_b200._lib import RappasError
    rng = np.random.default_rng(12345)
    alphabet = [">", "#", "\n", "\r", " ", "-", "A", "C", "G", "T", "N", "a", "\t", ">r", "\r\n", "ACGT", ">x y\n"]
    weights = np.array([2, 1, 6, 2, 2, 2, 6, 6, 6, 6, 1, 1, 1, 3, 2, 8, 4], float)
    weights /= weights.sum()
    n_ok = n_err = 0
    for it in range(300):
        text = ">h0\n" * int(rng.random() < 0.8) + "".join(rng.choice(alphabet, int(rng.integers(0, 120)), p=weights))
        monkeypatch.setenv("RP_HOST_THREADS", str(1 + it % 5))
        try:
            exp = ref_host.read_fasta(text)
        except Exception:
            exp = None
        if exp is not None and len(exp) == 0:
            exp = None  # "No valid fasta sequences were found"
        if exp is None:
            with pytest.raises(RappasError):
                ingest.QueryFile.from_text(text)
            n_err += 1
            continue
        q = ingest.QueryFile.from_text(text)
        assert q.headers == [h for h, _ in exp], repr(text)
        assert [q.unique.read(int(u)) for u in q.unique_of] == [s for _, s in exp], repr(text)
        groups = {}
        for r, (_, s) in enumerate(exp):
            assert q.group_of[r] == groups.setdefault(s.replace("-", ""), len(groups)), repr(text)
        n_ok += 1
    assert n_ok > 150 and n_err > 5
