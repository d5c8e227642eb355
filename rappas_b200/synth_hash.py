"""Host restatement (numpy) of the hash-defined synthetic DB that the GPU generates per partition on the device
(rappas_b200/csrc/rp_synth.h, rp_synthdb.cu).  Every property of a key is a pure function
of (seed, code), so the keys a read sample probes can be regenerated here and handed to the CPU oracle as an
ordinary CSR DB -- the > 1-HBM DB itself never exists on the host."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import synth

U = np.uint64
_M1, _M2 = U(0xBF58476D1CE4E5B9), U(0x94D049BB133111EB)


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U(30))) * _M1
        z = (z ^ (z >> U(27))) * _M2
    return z ^ (z >> U(31))


def plen_table(mean_postings: float, n_nodes: int) -> np.ndarray:
    """inverse CDF of P = min(N, 1 + Geometric(1/mean)) at 65 536 quantiles (what rp_db_synth_partition takes)"""
    q = (np.arange(65536, dtype=np.float64) + 0.5) / 65536.0
    if mean_postings <= 1.0:
        return np.ones(65536, np.uint16)
    p = 1.0 / float(mean_postings)
    P = 1 + np.floor(np.log1p(-q) / math.log1p(-p)).astype(np.int64)
    return np.clip(P, 1, n_nodes).astype(np.uint16)


@dataclass
class HashDB:
    k: int
    n_nodes: int
    seed: int
    occupancy: float
    mean_postings: float
    omega: float = 1.5

    def __post_init__(self):
        self.alphabet = synth.ALPHA_NUCL
        self.thr_lin, self.thr_log10 = synth.threshold(self.omega, self.alphabet, self.k)
        self.plen = plen_table(self.mean_postings, self.n_nodes)
        self.occ32 = 0xFFFFFFFF if self.occupancy >= 1.0 else int(self.occupancy * 4294967296.0)

    # ---- rp_synth.h, line by line
    def key_hash(self, codes):
        codes = np.asarray(codes, dtype=np.uint64)
        with np.errstate(over="ignore"):
            return _mix(U(self.seed) * U(0x9E3779B97F4A7C15) + codes + U(1))

    def present(self, h):
        return (h >> U(32)).astype(np.uint64) < U(self.occ32)

    def lengths(self, h):
        return self.plen[((h >> U(16)) & U(0xFFFF)).astype(np.int64)].astype(np.int64)

    def starts(self, h):
        return (_mix(h ^ U(0xA5A5A5A5A5A5A5A5)) % U(self.n_nodes)).astype(np.int64)

    def scores(self, h_rep, i):
        """h_rep: the key hash repeated per posting; i: index of the posting inside its key"""
        with np.errstate(over="ignore"):
            hp = _mix(h_rep + (i.astype(np.uint64) + U(1)) * U(0xD1B54A32D192ED03))
        u = (hp >> U(40)).astype(np.float32) * np.float32(5.9604644775390625e-8)
        return (np.float32(self.thr_log10) * (u * u).astype(np.float32)).astype(np.float32)

    # ---- the sub-DB of the given codes (those that are keys), as the CSR arrays the oracle / rp_db_load take
    def sub_db(self, codes) -> synth.SynthDB:
        codes = np.unique(np.asarray(codes, dtype=np.uint64))
        h = self.key_hash(codes)
        keep = self.present(h)
        codes, h = codes[keep], h[keep]
        plen = self.lengths(h)
        off = np.zeros(codes.shape[0] + 1, np.uint64)
        np.cumsum(plen, out=off[1:].view(np.int64))
        key_of = np.repeat(np.arange(codes.shape[0], dtype=np.int64), plen)
        i = np.arange(int(off[-1]), dtype=np.int64) - off[:-1].astype(np.int64)[key_of]
        node = ((self.starts(h)[key_of] + i) % self.n_nodes).astype(np.uint16)
        score = self.scores(h[key_of], i)
        return synth.SynthDB(self.alphabet, self.k, self.n_nodes, self.thr_lin, self.thr_log10, codes, off, node, score)

    def expected_keys(self) -> float:
        return self.occupancy * 4.0 ** self.k


def probed_codes(reads: synth.ReadBatch, k: int) -> np.ndarray:
    """every k-mer code a batch of nucleotide reads can probe: the plain windows and all alternatives of windows
    with exactly one IUPAC / N / gap character (maxAmbigPerMer = 1 for k <= 15).  Windows with other characters
    (or more ambiguities) probe nothing."""
    lut = np.full(256, 255, np.uint8)
    for ch, st in ((b"A", 0), (b"T", 1), (b"U", 1), (b"C", 2), (b"G", 3)):
        lut[ch[0]] = st
        lut[ch.lower()[0]] = st
    alts = {"R": (0, 3), "Y": (2, 1), "S": (2, 3), "W": (0, 1), "K": (3, 1), "M": (0, 2), "B": (2, 3, 1), "D": (0, 3, 1),
            "H": (0, 2, 1), "V": (0, 2, 3), "N": (0, 2, 3, 1), ".": (0,), "-": (0,)}
    st = lut[reads.seq]
    amb = np.zeros(reads.seq.shape[0], bool)
    for ch in alts:
        amb |= (reads.seq == ord(ch)) | (reads.seq == ord(ch.lower()))
    out = []
    stz = np.where(st == 255, 0, st).astype(np.uint64)
    pow4 = (U(1) << (U(2) * np.arange(k, dtype=np.uint64)))
    for r in range(reads.n_reads):
        b0, b1 = int(reads.seq_off[r]), int(reads.seq_off[r + 1])
        n = b1 - b0 - k + 1
        if n <= 0:
            continue
        s, a, raw = stz[b0:b1], amb[b0:b1], reads.seq[b0:b1]
        bad = (st[b0:b1] == 255) & ~a
        win = np.lib.stride_tricks.sliding_window_view
        code = (win(s, k) * pow4).sum(axis=1).astype(np.uint64)
        na = win(a.astype(np.int32), k).sum(axis=1)
        nb = win(bad.astype(np.int32), k).sum(axis=1)
        out.append(code[(na == 0) & (nb == 0)])
        for j in np.nonzero((na == 1) & (nb == 0))[0]:
            o = int(np.argmax(a[j:j + k]))
            base = int(code[j])  # the ambiguous position contributed state 0
            for alt in alts[chr(raw[j + o]).upper()]:
                out.append(np.array([base + (alt << (2 * o))], dtype=np.uint64))
    return np.unique(np.concatenate(out)) if out else np.zeros(0, np.uint64)
