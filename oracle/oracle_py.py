"""Second, independent restatement of the RAPPAS placement hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/rappas_oracle.h): the Java reference cannot run here.  This file is a
deliberately *literal* transliteration of the Java control flow (byte-array words of length
k*n, dict-based hash of dicts, an array-backed PriorityQueue), written separately from
oracle/rappas_oracle.c so that a transcription slip in either shows up as a disagreement
(tests/test_oracle_cross.py).  Pure-Python loops: small cases only.

Citations are relative to /root/reference/src.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32
NEG_INF32 = f32(-np.inf)


# ----------------------------------------------------------------------------- States
class DNAStatesShifted:
    """core/DNAStatesShifted.java"""

    non_ambiguous = 4

    def __init__(self):
        A, T, C, G = 0, 1, 2, 3  # :33-34, :184-204
        self.states = {}
        for ch, b in (("A", A), ("T", T), ("U", T), ("C", C), ("G", G)):
            self.states[ch] = b
            self.states[ch.lower()] = b
        amb = {  # :62-96, literal array order
            "R": [A, G], "Y": [C, T], "S": [C, G], "W": [A, T], "K": [G, T], "M": [A, C],
            "B": [C, G, T], "D": [A, G, T], "H": [A, C, T], "V": [A, C, G], "N": [A, C, G, T],
        }
        self.ambiguous = {}
        for ch, alts in amb.items():
            self.ambiguous[ch] = alts
            self.ambiguous[ch.lower()] = alts
        # :57-58 `new byte[4]` never filled => four zeros (= four times 'A')
        self.ambiguous["."] = [0, 0, 0, 0]
        self.ambiguous["-"] = [0, 0, 0, 0]

    def is_ambiguous(self, c):  # :212-214
        return c in self.ambiguous

    def state_to_byte(self, c):  # :236-242
        if c not in self.states:
            raise ValueError(c)
        return self.states[c]

    def ambiguity_equivalence(self, c):  # :223-228
        return self.ambiguous[c]

    def compress_mer(self, word):  # :115-143
        byte_count = int(math.ceil((0.0 + len(word)) / 4))
        kmer = [0] * byte_count
        four = 0
        for i in range(len(word)):
            if i > 0 and i % 4 == 0:
                kmer[(i // 4) - 1] = four
                four = 0
            four = (four | (word[i] << (2 * (i % 4)))) & 0xFF
        kmer[-1] = four
        return bytes(kmer)


class AAStates:
    """core/AAStates.java"""

    non_ambiguous = 20

    def __init__(self, convert_uo=False):
        order = "RHKDESTNQCGPAILMFWYV"  # :23-28
        self.states = {}
        for i, ch in enumerate(order):  # :74-93
            self.states[ch] = i
            self.states[ch.lower()] = i
        all20 = list(range(20))
        self.ambiguous = {"-": all20, "*": all20, "!": all20, "X": all20, "x": all20}  # :97-101
        for ch, alts in (("B", [3, 7]), ("Z", [4, 8]), ("J", [13, 14])):  # :102-107
            self.ambiguous[ch] = alts
            self.ambiguous[ch.lower()] = alts
        if convert_uo:  # :118-123
            for ch, b in (("U", 9), ("u", 9), ("O", 14), ("o", 14)):
                self.states[ch] = b

    def is_ambiguous(self, c):
        return c in self.ambiguous

    def state_to_byte(self, c):
        if c not in self.states:
            raise ValueError(c)
        return self.states[c]

    def ambiguity_equivalence(self, c):
        return self.ambiguous[c]

    def compress_mer(self, word):  # :195-197 identity
        return bytes(word)


def make_states(alphabet: int):
    return DNAStatesShifted() if alphabet == 0 else AAStates(convert_uo=(alphabet == 2))


def threshold(omega: float, alphabet: int, k: int):
    """Main_DBBUILD_3.java:165-166"""
    n = 4 if alphabet == 0 else 20
    ratio = f32(omega) / f32(n)  # float / int -> float
    lin = f32(math.pow(0.0 + float(ratio), k))
    lg = f32(math.log10(float(lin)))
    return lin, lg


# ----------------------------------------------------------------------------- Knife
class UnsupportedState(Exception):
    pass


class NegativeArraySize(Exception):
    pass


class AmbigSequenceKnife:
    """core/algos/AmbigSequenceKnife.java, SAMPLING_LINEAR only (Main_PLACEMENT_v07.java:112)"""

    def __init__(self, k, states):
        self.k = k
        self.s = states
        self.max_ambig_per_mer = int(math.floor(math.pow(k, 1.0 / states.non_ambiguous)))  # :95

    def init(self, seq: str):  # initTables :98-174
        k = self.k
        self.mer_iterator = 0
        self.sequence = [0] * len(seq)
        self.ambiguity_count_per_mer = [0] * len(seq)
        self.ambiguity_offsets = {}
        for i, c in enumerate(seq):
            if self.s.is_ambiguous(c):
                for j in range(i - k + 1, i + 1):
                    if j > -1 and j < len(seq):
                        self.ambiguity_count_per_mer[j] += 1
                        self.ambiguity_offsets.setdefault(j, {})[i - j] = self.s.ambiguity_equivalence(c)
                self.sequence[i] = -1
            else:
                try:
                    self.sequence[i] = self.s.state_to_byte(c)
                except ValueError:
                    raise UnsupportedState(c)  # System.exit(1) in the reference, :124-128
        if len(seq) - k + 1 < 0:
            raise NegativeArraySize()  # new int[negative], :145
        self.mer_order = list(range(len(seq) - k + 1))

    def get_mer_count(self):
        return len(self.mer_order)

    def get_next_byte_word(self):  # :209-272 (minK == k: every linear window is full length)
        k = self.k
        if self.mer_iterator > len(self.mer_order) - 1:
            return None
        cur = self.mer_order[self.mer_iterator]
        word = self.sequence[cur:cur + k]
        if self.ambiguity_count_per_mer[self.mer_iterator] < 1:
            self.mer_iterator += 1
            return word
        if self.ambiguity_count_per_mer[self.mer_iterator] > self.max_ambig_per_mer:
            self.mer_iterator += 1
            return [0]  # new byte[1]
        alt_product = 1
        for off in self.ambiguity_offsets[cur]:
            alt_product *= len(self.ambiguity_offsets[cur][off])
        words = [0] * (alt_product * k)
        for i in range(k):
            if self.sequence[cur + i] != -1:
                for j in range(alt_product):
                    words[i + j * k] = self.sequence[cur + i]
            else:
                alts = self.ambiguity_offsets[cur][i]
                jump = 0
                for _step in range(alt_product // len(alts)):
                    for j in range(len(alts)):
                        words[i + jump * k] = alts[j]
                        jump += 1
        self.mer_iterator += 1
        return words


# ----------------------------------------------------------------------------- Hash
class CustomHash:
    """core/hash/CustomHash_v4_FastUtil81.java: map packed k-mer bytes -> ordered {node: score}"""

    def __init__(self, states):
        self.states = states
        self.hash = {}

    def add_postings(self, word_states, postings):
        """postings: list of (node, score) in char2FloatEntrySet() iteration order"""
        self.hash[self.states.compress_mer(list(word_states))] = [(int(x), f32(v)) for x, v in postings]

    def get_pairs_of_top_position2(self, key: bytes):  # :146-153
        return self.hash.get(key)


# ----------------------------------------------------------------------------- PriorityQueue
def float_compare(a, b):
    """java.lang.Float.compare"""
    if a < b:
        return -1
    if a > b:
        return 1

    def bits(x):
        if x != x:
            return 0x7FC00000
        v = int(np.array(x, dtype=np.float32).view(np.int32))
        return v

    ia, ib = bits(a), bits(b)
    return 0 if ia == ib else (-1 if ia < ib else 1)


class JavaPriorityQueue:
    """java.util.PriorityQueue<Score> (JDK 8): array heap, siftUp/siftDownComparable"""

    def __init__(self):
        self.q = []

    def add(self, e):
        k = len(self.q)
        self.q.append(e)
        while k > 0:
            parent = (k - 1) >> 1
            if float_compare(e[1], self.q[parent][1]) >= 0:
                break
            self.q[k] = self.q[parent]
            k = parent
        self.q[k] = e

    def remove(self):
        s = len(self.q) - 1
        result = self.q[0]
        x = self.q.pop()
        if s != 0:
            k, half = 0, s >> 1
            while k < half:
                child = 2 * k + 1
                c = self.q[child]
                right = child + 1
                if right < s and float_compare(c[1], self.q[right][1]) > 0:
                    child = right
                    c = self.q[child]
                if float_compare(x[1], c[1]) <= 0:
                    break
                self.q[k] = c
                k = child
            self.q[k] = x
        return result

    def __iter__(self):
        return iter(list(self.q))

    def __len__(self):
        return len(self.q)


# ----------------------------------------------------------------------------- Placement
@dataclass
class Session:
    alphabet: int
    k: int
    n_nodes: int
    thr_lin: np.float32
    thr_log10: np.float32
    hash: CustomHash = None
    states: object = None


@dataclass
class ReadResult:
    status: int
    rows: list = field(default_factory=list)  # (node, score f32, lwr f64) best first
    counts: tuple = (0, 0, 0, 0)              # windows, matched, ambiguous treated, skipped
    S: dict = field(default_factory=dict)     # node -> final f32 score
    C: dict = field(default_factory=dict)


class PlacementProcess:
    """core/algos/PlacementProcess.java, processQueries :471-1118 (per-read body only)"""

    def __init__(self, session: Session, ns_bound=-np.inf):
        self.session = session
        self.ns_bound = f32(ns_bound)

    def _lookup(self, word):
        s = self.session
        return s.hash.get_pairs_of_top_position2(s.states.compress_mer(word))

    def _treat_amb(self, w, Q, L, C, S, with_max):  # :1129-1174 / :1185-1236
        s = self.session
        T, Tlin = s.thr_log10, s.thr_lin
        S_amb, C_amb, L_amb = {}, {}, []
        W_size = len(w) // s.k
        for i in range(W_size):
            w_prime = w[i * s.k:(i + 1) * s.k]
            pairs = self._lookup(w_prime)
            if pairs is None:
                continue
            for x, v in pairs:
                if C_amb.get(x, 0) == 0:
                    L_amb.append(x)
                    if with_max:
                        S_amb[x] = v
                C_amb[x] = C_amb.get(x, 0) + 1
                if with_max:
                    if v > S_amb[x]:
                        S_amb[x] = v
                else:
                    S_amb[x] = f32(float(S_amb.get(x, f32(0))) + math.pow(10.0, float(v)))
        for x in L_amb:
            if C[x] == 0:
                L.append(x)
                S[x] = f32(Q) * T
            C[x] += 1
            if with_max:
                S[x] = S[x] + (S_amb[x] - T)
            else:
                avg = (S_amb[x] + f32(W_size - C_amb[x]) * Tlin) / f32(W_size)
                S[x] = f32(float(S[x]) + (math.log10(float(avg)) - float(T)))

    def fill_best_score_list(self, S, L, best_list, num_best):  # :396-451
        heap = JavaPriorityQueue()
        for node in L:
            heap.add((node, S[node]))
            if len(heap) > num_best:
                heap.remove()
        total = 0.0
        lowest = f32(0.0)
        best = f32(-3.4028234663852886e38)
        for node, sc in heap:
            total += math.pow(10.0, float(sc))
            if sc < lowest:
                lowest = sc
            if sc > best:
                best = sc
        for i, e in enumerate(heap):
            best_list[i] = e
        # Arrays.sort(Object[]): stable ascending
        import functools
        best_list.sort(key=functools.cmp_to_key(lambda a, b: float_compare(a[1], b[1])))
        shift = best if f32(-308.0) >= lowest else f32(0.0)
        if shift != f32(0.0):
            total = 0.0
            for ii in range(len(best_list) - num_best, len(best_list)):
                total += math.pow(10.0, float(f32(best_list[ii][1] - shift)))
        return total

    def place_read(self, seq: str, keep_at_most=7, keep_factor=0.01, treat_amb=True, with_max=False) -> ReadResult:
        s = self.session
        T = s.thr_log10
        keep_factor = f32(keep_factor)
        sk = AmbigSequenceKnife(s.k, s.states)
        try:
            sk.init(seq)
        except UnsupportedState:
            return ReadResult(status=3)
        except NegativeArraySize:
            return ReadResult(status=2)
        N = s.n_nodes
        C = [0] * N
        S = [f32(0.0)] * N
        L = []
        q_count = amb_treated = matching = skipped = 0
        Q = sk.get_mer_count()
        while True:  # :687-764
            qw = sk.get_next_byte_word()
            if qw is None:
                break
            if len(qw) == 1:
                q_count += 1
                skipped += 1
                continue
            if len(qw) == s.k:
                pairs = self._lookup(qw)
                if pairs is None:
                    q_count += 1
                    continue
                matching += 1
                for x, v in pairs:
                    if C[x] == 0:
                        L.append(x)
                        S[x] = S[x] + f32(Q) * T
                    C[x] += 1
                    S[x] = S[x] + (v - T)
            else:
                if treat_amb:
                    amb_treated += 1
                    self._treat_amb(qw, Q, L, C, S, with_max)
                else:
                    q_count += 1
                    skipped += 1
                    continue
            q_count += 1
        counts = (q_count, matching, amb_treated, skipped)
        if len(L) < 1:  # :797-806
            return ReadResult(status=1, counts=counts)
        nb = keep_at_most if len(L) >= keep_at_most else len(L)
        best_list = [(-1, NEG_INF32)] * keep_at_most
        total = self.fill_best_score_list(S, L, best_list, nb)
        rows = []
        K = keep_at_most
        if best_list[K - 1][1] >= self.ns_bound:  # :974
            best = best_list[K - 1][1]
            lowest = best_list[K - nb][1]
            shift = best if f32(-308.0) >= lowest else f32(0.0)
            best_ratio = -1.0
            i = K - 1
            while i > K - nb - 1:
                lwr = math.pow(10.0, float(best_list[i][1]) - float(shift)) / total  # :392-393
                if i == K - 1:
                    best_ratio = lwr
                if i < K - 1 and lwr < best_ratio * float(keep_factor):
                    break
                rows.append((best_list[i][0], best_list[i][1], lwr))
                i -= 1
        return ReadResult(status=0, rows=rows, counts=counts,
                          S={x: S[x] for x in L}, C={x: C[x] for x in L})


def session_from_csr(alphabet, k, n_nodes, thr_lin, thr_log10, keys, offsets, post_node, post_score) -> Session:
    """Build the dict-of-lists hash from the flat CSR export (codes unpacked back to state words)."""
    states = make_states(alphabet)
    h = CustomHash(states)
    bits = 2 if alphabet == 0 else 5
    mask = (1 << bits) - 1
    for i, code in enumerate(keys):
        code = int(code)
        word = [(code >> (bits * j)) & mask for j in range(k)]
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        h.add_postings(word, [(int(post_node[p]), f32(post_score[p])) for p in range(lo, hi)])
    return Session(alphabet=alphabet, k=k, n_nodes=n_nodes, thr_lin=f32(thr_lin), thr_log10=f32(thr_log10),
                   hash=h, states=states)
