"""Literal pure-Python restatement of the RAPPAS phylo-k-mer generation.  TEST INFRASTRUCTURE ONLY.

A second, independent transliteration (class with the Java field names, Java recursion, a dict of dicts
for the hash) used to cross-check oracle/dbbuild_oracle.c on small inputs.  PARITY UNPINNED (see the C file).

  WordExplorer_v3           core/algos/WordExplorer_v3.java:33-199
  driver loop               main_v2/Main_DBBUILD_3.java:648-716
  addTuple                  core/hash/CustomHash_v4_FastUtil81.java:73-90
"""
import numpy as np

f32 = np.float32


class WordExplorer_v3:
    def __init__(self, session, refPosition, nodeId, doGapJumps, limitTo1Jump):
        self.session = session
        self.gapIntervals = session["gapIntervals"]
        self.refPosition = refPosition
        self.nodeId = nodeId
        self.doGapJumps = doGapJumps
        self.limitTo1Jump = limitTo1Jump
        self.word = [0] * session["k"]
        self.current_k = 0
        self.currentLogSum = f32(0.0)
        self.boundReached = False
        self.boundReachingK = -1
        self.idxOfFirstJump = -1
        self.generateTupleCount = 0
        self.originalId = int(session["originalId"][nodeId])

    def exploreWords(self, i, j):
        s = self.session
        pp, states = s["pp"], s["states"]
        siteCount, stateCount = pp.shape[1], pp.shape[2]
        if i > siteCount - 1:
            return
        if self.current_k == 0:
            self.idxOfFirstJump = -1
        self.word[self.current_k] = int(states[self.nodeId, i, j])
        # currentLogSum += getPP(...)  with getPP returning double: float = (float)((double)float + double)
        self.currentLogSum = f32(np.float64(self.currentLogSum) + np.float64(pp[self.nodeId, i, j]))
        self.boundReached = bool(self.currentLogSum < s["PPStarThresholdAsLog10"])
        if self.boundReached:
            self.boundReachingK = self.current_k
        if self.current_k == s["k"] - 1:
            if not self.boundReached:
                s["addTuple"](tuple(self.word), self.currentLogSum, self.originalId)
                self.generateTupleCount += 1
            self.currentLogSum = f32(np.float64(self.currentLogSum) - np.float64(pp[self.nodeId, i, j]))
            return
        else:
            for j2 in range(stateCount):
                if self.boundReached and self.boundReachingK == self.current_k + 1:
                    break
                self.current_k += 1
                self.exploreWords(i + 1, j2)
                self.current_k -= 1
                if self.doGapJumps and i < siteCount - 1:
                    if self.gapIntervals[i + 1] is not None:
                        if not self.limitTo1Jump:
                            for length in self.gapIntervals[i + 1]:
                                self.current_k += 1
                                self.exploreWords((i + 1) + length, j2)
                                self.current_k -= 1
                        else:
                            if self.idxOfFirstJump == -1:
                                self.idxOfFirstJump = i
                                for length in self.gapIntervals[i + 1]:
                                    self.current_k += 1
                                    self.exploreWords((i + 1) + length, j2)
                                    self.current_k -= 1
        self.currentLogSum = f32(np.float64(self.currentLogSum) - np.float64(pp[self.nodeId, i, j]))


def build(alphabet, k, pp, states, original_id, thr_log10, gap_intervals=None, gap_jumps=0):
    """-> (hash: {code: {node: f32}}, n_tuples).  gap_intervals: list (per site) of None or list of lengths."""
    pp = np.asarray(pp, dtype=np.float32)
    states = np.asarray(states, dtype=np.uint8)
    n_nodes, n_sites, n_states = pp.shape
    bits = 2 if alphabet == 0 else 5
    table = {}

    def addTuple(word, PPStar, nodeId):
        code = 0
        for i, b in enumerate(word):
            code |= int(b) << (bits * i)
        m = table.get(code)
        if m is not None:
            old = m.get(nodeId)
            if old is None:
                m[nodeId] = PPStar        # putIfAbsent inserted it; returned the default 10.0 -> no replacement
            elif PPStar > old:
                m[nodeId] = PPStar
        else:
            table[code] = {nodeId: PPStar}

    session = dict(k=k, pp=pp, states=states, originalId=original_id, PPStarThresholdAsLog10=f32(thr_log10),
                   gapIntervals=gap_intervals if gap_intervals is not None else [None] * n_sites, addTuple=addTuple)
    total = 0
    for nodeId in range(n_nodes):
        for pos in range(0, n_sites - k + 2):
            wd = WordExplorer_v3(session, pos, nodeId, gap_jumps != 0, gap_jumps == 2)
            for j in range(n_states):
                wd.exploreWords(pos, j)
            total += wd.generateTupleCount
    return table, total
