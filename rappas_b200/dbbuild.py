"""Phylo-k-mer generation on the GPU (rp_dbbuild_*; the work is CUDA in csrc/rp_dbbuild.cu).

Mirrors main_v2/Main_DBBUILD_3.java:648-750 (the per-node, per-position WordExplorer_v3 loop) and
core/hash/CustomHash_v4_FastUtil81.java:73-90 (addTuple).  Output = the CSR arrays Database.from_arrays takes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, load


def prepare_posteriors(probs, state_of_column, site_pp_threshold=np.float32(1.4e-45), as_log10=True):
    """Raw AR posteriors [n_nodes][n_sites][n_states] (columns in the AR file's state order) -> (pp, states) as
    PProbasSorted holds them: clamped, log10, stable-sorted highest first (inputs/PHYMLWrapper.java:206-229)."""
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    soc = np.ascontiguousarray(state_of_column, dtype=np.uint8)
    pp, st = np.empty_like(probs), np.empty(probs.shape, np.uint8)
    check(load()["pp_prepare"](_abi.ptr(probs), _abi.ptr(soc), probs.shape[0], probs.shape[1], probs.shape[2],
                               C.c_float(site_pp_threshold), int(bool(as_log10)), _abi.ptr(pp), _abi.ptr(st)))
    return pp, st


def gap_intervals(rows):
    """rows: aligned sequences of equal length (str / bytes) -> (gap_off, gap_len), Alignment.updateGapIntervals."""
    m = np.ascontiguousarray([np.frombuffer(r.encode() if isinstance(r, str) else bytes(r), np.uint8) for r in rows], dtype=np.uint8)
    off = np.zeros(m.shape[1] + 1, np.uint64)
    n = C.c_uint64()
    fn = load()
    check(fn["gap_intervals"](_abi.ptr(m), m.shape[0], m.shape[1], _abi.ptr(off), None, 0, C.byref(n)))
    lens = np.zeros(max(1, n.value), np.int32)
    check(fn["gap_intervals"](_abi.ptr(m), m.shape[0], m.shape[1], _abi.ptr(off), _abi.ptr(lens), lens.shape[0], C.byref(n)))
    return off, lens[:n.value]


def build_db(alphabet, k, pp, states, original_id, thr_log10, gap_off=None, gap_len=None, gap_jumps=0, device=0):
    """pp, states: [n_nodes][n_sites][n_states] (PProbasSorted: log10 posteriors, descending per site, and their
    states); original_id[n_nodes]; gap_off/gap_len: CSR of Alignment.getGapIntervals(); gap_jumps 0 / 1 / 2 =
    off / every combination / at most one jump.  -> dict(keys, offsets, post_node, post_score, n_tuples, kernel_ms)"""
    fn = load()
    pp = np.ascontiguousarray(pp, dtype=np.float32)
    states = np.ascontiguousarray(states, dtype=np.uint8)
    original_id = np.ascontiguousarray(original_id, dtype=np.uint16)
    if gap_off is not None:
        gap_off = np.ascontiguousarray(gap_off, dtype=np.uint64)
        gap_len = np.ascontiguousarray(gap_len, dtype=np.int32)
    d = _abi.RpDbBuildDesc(int(alphabet), int(k), pp.shape[0], pp.shape[1], pp.shape[2], float(thr_log10), int(gap_jumps), 0)
    h = C.c_void_p()
    check(fn["dbbuild_run"](C.byref(d), _abi.ptr(pp), _abi.ptr(states), _abi.ptr(original_id),
                            _abi.ptr(gap_off) if gap_off is not None else None,
                            _abi.ptr(gap_len) if gap_len is not None else None, int(device), C.byref(h)))
    try:
        nk, npost, nt, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
        keys, offs, nodes, scores = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(fn["dbbuild_result"](h, C.byref(nk), C.byref(npost), C.byref(nt), C.byref(keys), C.byref(offs),
                                   C.byref(nodes), C.byref(scores), C.byref(ms)))

        def view(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            ct = np.ctypeslib.as_ctypes_type(dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy()
        return dict(keys=view(keys, nk.value, np.uint64), offsets=view(offs, nk.value + 1, np.uint64),
                    post_node=view(nodes, npost.value, np.uint16), post_score=view(scores, npost.value, np.float32),
                    n_tuples=nt.value, kernel_ms=ms.value)
    finally:
        fn["dbbuild_free"](h)
