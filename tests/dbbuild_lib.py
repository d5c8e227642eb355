"""Synthetic inputs and oracle bindings for the DB-build row.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O
from rappas_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_inputs(alphabet, k, n_nodes, n_sites, seed, peak=0.9, gap_rate=0.0):
    """PProbasSorted-like arrays: per (node, site) a probability vector peaked on one state (mass `peak` +- jitter),
    sorted descending, as log10 f32; the states in that order; original ids; optional gap intervals (CSR)."""
    rng = np.random.default_rng(seed)
    ns = 4 if alphabet == 0 else 20
    conc = rng.uniform(peak - 0.25, min(0.999, peak + 0.09), (n_nodes, n_sites, 1))
    rest = rng.dirichlet(np.ones(ns - 1), (n_nodes, n_sites)) * (1.0 - conc)
    probs = np.concatenate([conc, rest], axis=2)
    order = np.argsort(-probs, axis=2, kind="stable")
    probs = np.take_along_axis(probs, order, axis=2)
    perm = np.stack([rng.permutation(ns) for _ in range(n_nodes * n_sites)]).reshape(n_nodes, n_sites, ns)
    states = np.take_along_axis(perm, order, axis=2).astype(np.uint8)
    pp = np.log10(np.maximum(probs, 1e-30)).astype(np.float32)
    original_id = rng.permutation(4 * n_nodes)[:n_nodes].astype(np.uint16)
    gap_off = gap_len = None
    if gap_rate > 0:
        lens, off = [], [0]
        for _ in range(n_sites):
            if rng.random() < gap_rate:
                lens += sorted(set(int(x) for x in rng.integers(1, 4, int(rng.integers(1, 3)))))
            off.append(len(lens))
        gap_off = np.asarray(off, np.uint64)
        gap_len = np.asarray(lens if lens else [0], np.int32)
    return pp, states, original_id, gap_off, gap_len


def desc(alphabet, k, pp, thr_log10, gap_jumps):
    return _abi.RpDbBuildDesc(int(alphabet), int(k), pp.shape[0], pp.shape[1], pp.shape[2], float(thr_log10),
                              int(gap_jumps), 0)


def oracle_build(alphabet, k, pp, states, original_id, thr_log10, gap_off=None, gap_len=None, gap_jumps=0):
    """-> dict(keys, offsets, post_node, post_score, n_tuples) from oracle/dbbuild_oracle.c"""
    O.lib()
    lib = C.CDLL(O._SO)
    d = desc(alphabet, k, pp, thr_log10, gap_jumps)
    nk, npost, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    keys, offs, nodes, scores = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    lib.rpo_dbbuild.restype = C.c_int
    rc = lib.rpo_dbbuild(C.byref(d), _abi.ptr(pp), _abi.ptr(states), _abi.ptr(original_id),
                         _abi.ptr(gap_off) if gap_off is not None else None,
                         _abi.ptr(gap_len) if gap_len is not None else None,
                         C.byref(nk), C.byref(npost), C.byref(nt), C.byref(keys), C.byref(offs), C.byref(nodes),
                         C.byref(scores))
    assert rc == 0, rc

    def view(p, n, dt):
        ct = np.ctypeslib.as_ctypes_type(dt)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
    out = dict(keys=view(keys, nk.value, np.uint64), offsets=view(offs, nk.value + 1, np.uint64),
               post_node=view(nodes, npost.value, np.uint16), post_score=view(scores, npost.value, np.float32),
               n_tuples=nt.value)
    lib.rpo_dbbuild_free_arrays.restype = None
    lib.rpo_dbbuild_free_arrays(keys, offs, nodes, scores)
    return out


_core = None


def core_tuples(alphabet, k, pp, states, original_id, thr_log10, gap_off=None, gap_len=None, gap_jumps=0, cap=1 << 22):
    """The product's explorer state machine compiled for the host (tests/helpers/dbbuild_core_host.cpp):
    -> (codes, nodes, scores) of every addTuple, in order."""
    global _core
    so = os.path.join(ROOT, "tests", "helpers", "dbbuild_core_host.so")
    src = os.path.join(ROOT, "tests", "helpers", "dbbuild_core_host.cpp")
    hdr = os.path.join(ROOT, "rappas_b200", "csrc", "rp_dbbuild_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
        _core = None
    if _core is None:
        _core = C.CDLL(so)
        _core.core_dbbuild_tuples.restype = C.c_int
    codes, nodes, scores = np.zeros(cap, np.uint64), np.zeros(cap, np.uint16), np.zeros(cap, np.float32)
    n = C.c_uint64()
    rc = _core.core_dbbuild_tuples(int(alphabet), int(k), pp.shape[0], pp.shape[1], pp.shape[2], C.c_float(thr_log10),
                                   int(gap_jumps), _abi.ptr(pp), _abi.ptr(states), _abi.ptr(original_id),
                                   _abi.ptr(gap_off) if gap_off is not None else None,
                                   _abi.ptr(gap_len) if gap_len is not None else None, C.c_uint64(cap), C.byref(n),
                                   _abi.ptr(codes), _abi.ptr(nodes), _abi.ptr(scores))
    assert rc == 0, "tuple capacity too small"
    return codes[:n.value], nodes[:n.value], scores[:n.value]


def csr_from_tuples(codes, nodes, scores):
    """addTuple semantics on a tuple list: max per (k-mer, node); keys ascending, postings by ascending node."""
    if codes.size == 0:
        return dict(keys=np.zeros(0, np.uint64), offsets=np.zeros(1, np.uint64), post_node=np.zeros(0, np.uint16),
                    post_score=np.zeros(0, np.float32))
    order = np.lexsort((nodes, codes))
    c, n, s = codes[order], nodes[order], scores[order]
    new = np.ones(c.size, bool)
    new[1:] = (c[1:] != c[:-1]) | (n[1:] != n[:-1])
    grp = np.cumsum(new) - 1
    best = np.full(int(grp[-1]) + 1, -np.inf, np.float32)
    np.maximum.at(best, grp, s)
    pc, pn = c[new], n[new]
    knew = np.ones(pc.size, bool)
    knew[1:] = pc[1:] != pc[:-1]
    keys = pc[knew]
    offsets = np.concatenate([np.flatnonzero(knew), [pc.size]]).astype(np.uint64)
    return dict(keys=keys, offsets=offsets, post_node=pn, post_score=best)


def assert_csr_equal(a, b):
    for f in ("keys", "offsets", "post_node"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["post_score"].view(np.uint32), b["post_score"].view(np.uint32)), "post_score bits"
