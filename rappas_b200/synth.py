"""Seeded synthetic phylo-kmer DBs and query reads of the shapes named in BASELINE.json.configs.

The reference ships no DB, no reads and publishes no posting-length distribution
(SURVEY.md 8d), so everything here is a *declared assumption*:

  keys      nucleotide: round(occupancy * 4^k) distinct codes drawn without replacement
            ("with k<14 we generally get at least 75% of the possible k-mers",
            core/hash/CustomHash_v4_FastUtil81.java:49); amino: the distinct k-mers of a random
            "ancestral" sequence of n_keys residues (20^k/10 = the reference's initial capacity, :52)
  postings  P = min(N, Geometric(1/mean_postings)) per key; node ids = a contiguous run
            (neighbouring edges share k-mers) starting at a random node, wrapped mod N, so a
            key never lists a node twice (CustomHash_v4_FastUtil81.addTuple keeps one value per
            (k-mer,node), :76-89); scores v = T * U(0,1)^2, i.e. T <= v <= 0
            (WordExplorer_v3.java:119-121 prunes below T)
  reads     uniform random residues (hit rate = key-space occupancy) or mutated substrings of
            the ancestral sequence (amino); optional IUPAC / N / gap injection and U[lo,hi] lengths

Seeds follow SURVEY.md 8d: DB = 42 + config index, reads = 1042 + config index.
Host-side numpy only; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

ALPHA_NUCL, ALPHA_AMINO, ALPHA_AMINO_UO = 0, 1, 2

NUCL_LETTERS = np.frombuffer(b"ATCG", dtype=np.uint8)                  # state -> letter, DNAStatesShifted.java:33
AMINO_LETTERS = np.frombuffer(b"RHKDESTNQCGPAILMFWYV", dtype=np.uint8)  # AAStates.java:23-28
NUCL_IUPAC = np.frombuffer(b"RYSWKMBDHV", dtype=np.uint8)
AMINO_AMBIG = np.frombuffer(b"BZJX", dtype=np.uint8)


def threshold(omega: float, alphabet: int, k: int):
    """(T_lin, T_log10) with the float/double sequence of Main_DBBUILD_3.java:165-166."""
    n = 4 if alphabet == ALPHA_NUCL else 20
    ratio = np.float32(omega) / np.float32(n)
    lin = np.float32(math.pow(0.0 + float(ratio), k))
    lg = np.float32(math.log10(float(lin)))
    return lin, lg


@dataclass
class SynthDB:
    alphabet: int
    k: int
    n_nodes: int
    thr_lin: np.float32
    thr_log10: np.float32
    keys: np.ndarray        # uint64 [n_keys]
    offsets: np.ndarray     # uint64 [n_keys+1]
    post_node: np.ndarray   # uint16 [n_postings]
    post_score: np.ndarray  # float32 [n_postings]
    genome: np.ndarray | None = None  # state bytes of the ancestral sequence (amino / genome mode)

    @property
    def n_keys(self):
        return int(self.keys.shape[0])

    @property
    def n_postings(self):
        return int(self.post_node.shape[0])

    @property
    def bits(self):
        return 2 if self.alphabet == ALPHA_NUCL else 5


def _window_codes(states: np.ndarray, k: int, bits: int) -> np.ndarray:
    """codes of all len-k+1 windows: sum b_i << (bits*i), first residue least significant"""
    n = states.shape[0] - k + 1
    code = np.zeros(n, dtype=np.uint64)
    for i in range(k):
        code |= states[i:i + n].astype(np.uint64) << np.uint64(bits * i)
    return code


def make_db(alphabet: int, k: int, n_nodes: int, n_keys: int, mean_postings: float, seed: int,
            omega: float = 1.5, key_mode: str | None = None) -> SynthDB:
    rng = np.random.default_rng(seed)
    thr_lin, thr_log10 = threshold(omega, alphabet, k)
    bits = 2 if alphabet == ALPHA_NUCL else 5
    nstates = 4 if alphabet == ALPHA_NUCL else 20
    if key_mode is None:
        key_mode = "random" if alphabet == ALPHA_NUCL else "genome"
    genome = None
    if key_mode == "random":
        space = nstates ** k
        if alphabet != ALPHA_NUCL:
            raise ValueError("random key mode is nucleotide-only (dense code space)")
        if n_keys > space:
            raise ValueError("n_keys exceeds the key space")
        if space <= (1 << 27):
            keys = rng.permutation(space)[:n_keys].astype(np.uint64)
        else:  # sparse draw for a large space
            keys = np.unique(rng.integers(0, space, size=int(n_keys * 1.05) + 16, dtype=np.uint64))
            keys = rng.permutation(keys)[:n_keys]
    elif key_mode == "genome":
        genome = rng.integers(0, nstates, size=n_keys + k - 1, dtype=np.uint8)
        keys = np.unique(_window_codes(genome, k, bits))
        keys = rng.permutation(keys)
    else:
        raise ValueError(key_mode)
    nk = keys.shape[0]
    plen = np.minimum(rng.geometric(1.0 / mean_postings, size=nk), n_nodes).astype(np.int64)
    offsets = np.zeros(nk + 1, dtype=np.uint64)
    np.cumsum(plen, out=offsets[1:].view(np.int64))
    total = int(offsets[-1])
    start = rng.integers(0, n_nodes, size=nk, dtype=np.int64)
    post_node = np.empty(total, dtype=np.uint16)
    post_score = np.empty(total, dtype=np.float32)
    CH = 1 << 22  # keys per chunk: bounds the temporaries for the 600 M-posting config
    for c0 in range(0, nk, CH):
        c1 = min(nk, c0 + CH)
        p0, p1 = int(offsets[c0]), int(offsets[c1])
        pl = plen[c0:c1]
        key_of = np.repeat(np.arange(c1 - c0, dtype=np.int64), pl)
        pos = np.arange(p1 - p0, dtype=np.int64) - (offsets[c0:c1].astype(np.int64) - p0)[key_of]
        post_node[p0:p1] = ((start[c0:c1][key_of] + pos) % n_nodes).astype(np.uint16)
        u = rng.random(p1 - p0, dtype=np.float32)
        post_score[p0:p1] = (thr_log10 * (u * u)).astype(np.float32)
    return SynthDB(alphabet, k, n_nodes, thr_lin, thr_log10, keys, offsets, post_node, post_score, genome)


@dataclass
class ReadBatch:
    seq: np.ndarray      # uint8, concatenated ASCII
    seq_off: np.ndarray  # uint64 [n+1]

    @property
    def n_reads(self):
        return int(self.seq_off.shape[0] - 1)

    def read(self, i: int) -> str:
        return self.seq[int(self.seq_off[i]):int(self.seq_off[i + 1])].tobytes().decode("latin-1")

    def window_offsets(self, k: int) -> np.ndarray:
        lens = np.diff(self.seq_off.astype(np.int64))
        q = np.maximum(lens - k + 1, 0)
        out = np.zeros(self.n_reads + 1, dtype=np.uint64)
        np.cumsum(q, out=out[1:].view(np.int64))
        return out

    def slice(self, lo: int, hi: int) -> "ReadBatch":
        b0, b1 = int(self.seq_off[lo]), int(self.seq_off[hi])
        return ReadBatch(self.seq[b0:b1].copy(), (self.seq_off[lo:hi + 1] - self.seq_off[lo]).astype(np.uint64))


def reads_from_strings(reads) -> ReadBatch:
    bs = [r.encode("latin-1") if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:].view(np.int64))
    seq = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return ReadBatch(seq, off)


def make_reads(db: SynthDB, n_reads: int, length, seed: int, mode: str | None = None,
               mutation: float = 0.04, iupac_rate: float = 0.0, n_rate: float = 0.0,
               gap_rate: float = 0.0, lowercase_rate: float = 0.0) -> ReadBatch:
    """length: int or (lo, hi) inclusive for U{lo..hi}."""
    rng = np.random.default_rng(seed)
    nucl = db.alphabet == ALPHA_NUCL
    letters = NUCL_LETTERS if nucl else AMINO_LETTERS
    nstates = 4 if nucl else 20
    if mode is None:
        mode = "genome" if db.genome is not None else "uniform"
    if isinstance(length, (tuple, list)):
        lens = rng.integers(length[0], length[1] + 1, size=n_reads, dtype=np.int64)
    else:
        lens = np.full(n_reads, int(length), dtype=np.int64)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:].view(np.int64))
    total = int(off[-1])
    if mode == "uniform":
        states = rng.integers(0, nstates, size=total, dtype=np.uint8)
    elif mode == "genome":
        g = db.genome
        if g is None:
            raise ValueError("genome mode needs a DB generated with key_mode='genome'")
        maxlen = int(lens.max()) if n_reads else 0
        if g.shape[0] <= maxlen:
            raise ValueError("ancestral sequence shorter than the reads")
        st = rng.integers(0, g.shape[0] - maxlen, size=n_reads, dtype=np.int64)
        read_of = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
        pos = np.arange(total, dtype=np.int64) - off[:-1].astype(np.int64)[read_of]
        states = g[st[read_of] + pos].copy()
        mut = rng.random(total) < mutation
        states[mut] = rng.integers(0, nstates, size=int(mut.sum()), dtype=np.uint8)
    else:
        raise ValueError(mode)
    seq = letters[states]
    if iupac_rate > 0:
        m = rng.random(total) < iupac_rate
        pool = NUCL_IUPAC if nucl else AMINO_AMBIG
        seq[m] = pool[rng.integers(0, pool.shape[0], size=int(m.sum()))]
    if n_rate > 0:
        m = rng.random(total) < n_rate
        seq[m] = ord("N") if nucl else ord("X")
    if gap_rate > 0:
        m = rng.random(total) < gap_rate
        seq[m] = ord("-")
    if lowercase_rate > 0:
        m = (rng.random(total) < lowercase_rate) & (seq >= 65) & (seq <= 90)
        seq[m] += 32
    return ReadBatch(np.ascontiguousarray(seq), off)


# --------------------------------------------------------------------------- BASELINE.json configs
@dataclass
class Workload:
    name: str
    index: int
    alphabet: int
    k: int
    n_taxa: int
    n_keys: int
    mean_postings: float
    n_reads: int
    read_len: object
    iupac_rate: float = 0.0
    n_rate: float = 0.0
    key_mode: str | None = None  # None = make_db's default for the alphabet

    @property
    def n_nodes(self):
        return 2 * self.n_taxa - 1  # rooted binary tree, SURVEY.md 8 notation


def workload(index: int, scale: float = 1.0) -> Workload:
    """The five BASELINE.json configs; `scale` shrinks reads AND keys for tests."""
    if index == 1:
        w = Workload("cfg1_nucl_k8_150taxa_10k", 1, ALPHA_NUCL, 8, 150, round(0.75 * 4 ** 8), 16, 10_000, 150)
    elif index == 2:
        w = Workload("cfg2_nucl_k10_1000taxa_1M", 2, ALPHA_NUCL, 10, 1000, round(0.75 * 4 ** 10), 32, 1_000_000, 150)
    elif index == 3:
        w = Workload("cfg3_nucl_k12_5000taxa_10M", 3, ALPHA_NUCL, 12, 5000, round(0.75 * 4 ** 12), 48, 10_000_000, 150)
    elif index == 4:
        w = Workload("cfg4_amino_k6_500taxa_1M", 4, ALPHA_AMINO, 6, 500, 6_400_000, 16, 1_000_000, 50)
    elif index == 5:
        # stress shape, host-materialisable stand-in: k=15 keys drawn sparsely; the >1-GPU DB of
        # SURVEY 8d is generated per partition on device, not here
        # keys = the k-mers of a random 8 Mbp "genome" and reads = mutated substrings of it: with uniform
        # random reads a k=15 DB of 8 M keys (0.7 % of the key space) would hardly ever be hit
        w = Workload("cfg5_stress_k15_var_len", 5, ALPHA_NUCL, 15, 5000, 8_000_000, 48, 1_000_000, (50, 1500),
                     iupac_rate=0.005, n_rate=0.002, key_mode="genome")
    else:
        raise ValueError(index)
    if scale != 1.0:
        w.n_reads = max(1, int(w.n_reads * scale))
        w.n_keys = max(16, int(w.n_keys * scale))
    return w


def build(w: Workload, reads: bool = True, n_reads: int | None = None):
    db = make_db(w.alphabet, w.k, w.n_nodes, w.n_keys, w.mean_postings, seed=42 + w.index, key_mode=w.key_mode)
    if not reads:
        return db, None
    rb = make_reads(db, n_reads if n_reads is not None else w.n_reads, w.read_len, seed=1042 + w.index,
                    iupac_rate=w.iupac_rate, n_rate=w.n_rate)
    return db, rb
