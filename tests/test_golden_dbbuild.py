"""Committed vectors of the phylo-k-mer generation (tests/golden/dbbuild_*.npz, made by the literal Python
restatement): the recursive C oracle and the product's explorer state machine on the CPU, the CUDA path with -m gpu."""
import glob
import os

import numpy as np
import pytest

import dbbuild_lib as D

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "dbbuild_*.npz")))
FIELDS = ("keys", "offsets", "post_node", "post_score")


def load(path):
    z = np.load(path)
    goff = z["gap_off"] if z["gap_off"].size else None
    glen = z["gap_len"] if z["gap_off"].size else None
    args = (int(z["alphabet"]), int(z["k"]), z["pp"], z["states"], z["original_id"], float(z["thr_log10"]), goff, glen,
            int(z["gap_jumps"]))
    return args, {f: z[f] for f in FIELDS}, int(z["n_tuples"])


def test_fixtures_exist():
    assert len(FIXTURES) == 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_c_oracle_and_host_state_machine_reproduce_the_vectors(path):
    args, exp, n_tuples = load(path)
    o = D.oracle_build(*args)
    assert o["n_tuples"] == n_tuples
    D.assert_csr_equal({f: o[f] for f in FIELDS}, exp)
    codes, nodes, scores = D.core_tuples(*args)
    assert codes.size == n_tuples
    D.assert_csr_equal(D.csr_from_tuples(codes, nodes, scores), exp)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_build_reproduces_the_vectors(path):
    from rappas_b200 import dbbuild
    args, exp, n_tuples = load(path)
    g = dbbuild.build_db(*args)
    assert g["n_tuples"] == n_tuples
    D.assert_csr_equal({f: g[f] for f in FIELDS}, exp)
