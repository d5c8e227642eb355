"""Loader + numpy-friendly wrapper of the CPU oracle (oracle/librappas_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this.  Never reads /root/reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rappas_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_SO = os.path.join(ORACLE_DIR, "librappas_oracle.so")
_lib = None
_fn = None


def build(force=False):
    src = os.path.join(ORACLE_DIR, "rappas_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "librappas_oracle.so"])
    return _SO


def lib():
    global _lib, _fn
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _fn = _abi.bind(_lib, "rpo_", overrides=_abi.ORACLE_OVERRIDES, strict=False)
    return _fn


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle error %d: %s" % (rc, lib()["last_error"]().decode()))


def threshold(omega, alphabet, k):
    a, b = C.c_float(), C.c_float()
    lib()["threshold"](omega, alphabet, k, C.byref(a), C.byref(b))
    return np.float32(a.value), np.float32(b.value)


def pack_kmer(alphabet, states):
    s = np.ascontiguousarray(states, dtype=np.uint8)
    return int(lib()["pack_kmer"](alphabet, _abi.ptr(s), s.shape[0]))


class OracleDB:
    def __init__(self, db):
        """db: rappas_b200.synth.SynthDB (or anything with the same fields)"""
        self.db = db
        self.desc = _abi.RpDbDesc(db.alphabet, db.k, db.n_nodes, float(db.thr_log10), float(db.thr_lin), 0,
                                  db.n_keys, db.n_postings)
        self.h = C.c_void_p()
        _check(lib()["db_load"](C.byref(self.desc), _abi.ptr(np.ascontiguousarray(db.keys, dtype=np.uint64)),
                                _abi.ptr(np.ascontiguousarray(db.offsets, dtype=np.uint64)),
                                _abi.ptr(np.ascontiguousarray(db.post_node, dtype=np.uint16)),
                                _abi.ptr(np.ascontiguousarray(db.post_score, dtype=np.float32)), C.byref(self.h)))

    def close(self):
        if self.h:
            lib()["db_free"](self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def place(self, reads, cfg=None, threads=1):
        cfg = cfg or _abi.place_cfg()
        n, K = reads.n_reads, cfg.keep_at_most
        out = {
            "n_rows": np.zeros(n, np.int32), "node": np.zeros((n, K), np.uint16),
            "score": np.zeros((n, K), np.float32), "lwr": np.zeros((n, K), np.float64),
            "counts": np.zeros((n, 4), np.int32), "status": np.zeros(n, np.int32),
        }
        args = [self.h, C.byref(cfg), _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), n, _abi.ptr(out["n_rows"]),
                _abi.ptr(out["node"]), _abi.ptr(out["score"]), _abi.ptr(out["lwr"]), _abi.ptr(out["counts"]),
                _abi.ptr(out["status"])]
        if threads == 1:
            _check(lib()["place_batch"](*args))
        else:
            _check(lib()["place_batch_mt"](*args, threads))
        return out

    def extract(self, reads):
        k = self.db.k
        woff = reads.window_offsets(k)
        nw = int(woff[-1])
        out = {"win_off": woff, "code": np.zeros(nw, np.uint64), "kind": np.zeros(nw, np.uint8),
               "nalt": np.zeros(nw, np.int32), "hits": np.zeros(nw, np.int32),
               "status": np.zeros(reads.n_reads, np.int32)}
        _check(lib()["extract_kmers"](self.h, _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), reads.n_reads,
                                      _abi.ptr(woff), _abi.ptr(out["code"]), _abi.ptr(out["kind"]),
                                      _abi.ptr(out["nalt"]), _abi.ptr(out["hits"]), _abi.ptr(out["status"])))
        return out

    def node_scores(self, reads, cfg=None, hitcount=True):
        cfg = cfg or _abi.place_cfg()
        n, N = reads.n_reads, self.db.n_nodes
        S = np.zeros((n, N), np.float32)
        Cn = np.zeros((n, N), np.int32) if hitcount else None
        _check(lib()["node_scores"](self.h, C.byref(cfg), _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), n,
                                    _abi.ptr(S), _abi.ptr(Cn)))
        return S, Cn
