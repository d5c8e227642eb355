#!/usr/bin/env python
"""Regenerates tests/golden/dbbuild_*.npz -- small fixed input/output vectors of the phylo-k-mer generation.

PARITY UNPINNED (the Java reference cannot run here): the vectors are produced by oracle/dbbuild_py.py, the
literal Python transliteration of WordExplorer_v3.exploreWords + the Main_DBBUILD_3 driver loop + addTuple.
They pin the recursive C oracle (tests/test_golden_dbbuild.py, CPU), the product's explorer state machine
(compiled for the host) and the CUDA path (-m gpu) to a committed artefact.

    python tests/golden/make_golden_dbbuild.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import dbbuild_lib as D  # noqa: E402
from oracle import dbbuild_py  # noqa: E402

CASES = {
    # name: (alphabet, k, n_nodes, n_sites, seed, peak, gap_rate, gap_jumps, omega)
    "dbbuild_nucl_k5": (0, 5, 3, 16, 11, 0.85, 0.0, 0, 1.5),
    "dbbuild_nucl_k4_jumps_all": (0, 4, 2, 12, 12, 0.85, 0.4, 1, 1.5),
    "dbbuild_nucl_k4_one_jump": (0, 4, 2, 12, 12, 0.85, 0.4, 2, 1.5),
    "dbbuild_amino_k2": (1, 2, 2, 8, 13, 0.7, 0.0, 0, 1.5),
}


def threshold(alphabet, k, omega):
    # Main_DBBUILD_3.java:165-166
    lin = np.float32(np.power(np.float64(np.float32(omega) / np.float32(4 if alphabet == 0 else 20)), k))
    return np.float32(np.log10(np.float64(lin)))


def main():
    for name, (alphabet, k, n_nodes, n_sites, seed, peak, gap_rate, gap_jumps, omega) in CASES.items():
        pp, states, oid, goff, glen = D.make_inputs(alphabet, k, n_nodes, n_sites, seed, peak, gap_rate)
        thr = threshold(alphabet, k, omega)
        gaps = None
        if goff is not None:
            gaps = [list(map(int, glen[int(goff[i]):int(goff[i + 1])])) or None for i in range(n_sites)]
        table, n_tuples = dbbuild_py.build(alphabet, k, pp, states, oid, thr, gaps, gap_jumps)
        codes, nodes, scores = [], [], []
        for c in sorted(table):
            for nd in sorted(table[c]):
                codes.append(c); nodes.append(nd); scores.append(table[c][nd])
        csr = D.csr_from_tuples(np.asarray(codes, np.uint64), np.asarray(nodes, np.uint16), np.asarray(scores, np.float32))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), alphabet=alphabet, k=k, thr_log10=thr, gap_jumps=gap_jumps,
                            pp=pp, states=states, original_id=oid,
                            gap_off=goff if goff is not None else np.zeros(0, np.uint64),
                            gap_len=glen if glen is not None else np.zeros(0, np.int32),
                            n_tuples=n_tuples, **csr)
        print(name, "tuples", n_tuples, "keys", csr["keys"].size, "postings", csr["post_node"].size)


if __name__ == "__main__":
    main()
