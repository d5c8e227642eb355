"""Host-side handle on the EXCHANGE FORM of a hash-partitioned DB (rappas_b200/csrc/rp_xchg.cu; north_star:
"NCCL all-to-all of k-mer probes over NVLink, only for DBs exceeding one GPU's HBM").

  Exchange.nccl(partition, rank, world, id)   one rank per process / GPU; `place([reads])` is a collective
  Exchange.local(partitions)                  every rank in this process on one GPU (tests): `place` takes
                                              one batch per rank
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, load


def unique_id() -> np.ndarray:
    """NCCL unique id (rank 0 creates it, the host program hands it to the other ranks)."""
    a = np.zeros(_abi.RP_XCHG_ID_BYTES, np.uint8)
    check(load()["xchg_unique_id"](_abi.ptr(a)))
    return a


class Exchange:
    def __init__(self, handle, n_local, partitions):
        self._h = handle
        self.n_local = n_local
        self._parts = partitions  # the partitions must outlive the exchange

    @classmethod
    def nccl(cls, partition, rank: int, world: int, uid: np.ndarray):
        h = C.c_void_p()
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        check(load()["xchg_create"](partition._h, int(rank), int(world), _abi.ptr(uid), C.byref(h)))
        return cls(h, 1, [partition])

    @classmethod
    def local(cls, partitions):
        arr = (C.c_void_p * len(partitions))(*[p._h for p in partitions])
        h = C.c_void_p()
        check(load()["xchg_create_local"](arr, len(partitions), C.byref(h)))
        return cls(h, len(partitions), list(partitions))

    def close(self):
        if self._h:
            load()["xchg_free"](self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def place(self, batches, cfg=None, counts=True):
        """batches: one ReadBatch per local rank -> one result dict per local rank (as Database.place)."""
        cfg = cfg or _abi.place_cfg()
        K = cfg.keep_at_most
        assert len(batches) == self.n_local
        outs = []
        for rb in batches:
            n = rb.n_reads
            outs.append({"n_rows": np.empty(n, np.int32), "node": np.empty((n, K), np.uint16),
                         "score": np.empty((n, K), np.float32), "lwr": np.empty((n, K), np.float64),
                         "counts": np.empty((n, 4), np.int32) if counts else None, "status": np.empty(n, np.int32)})
        nl = self.n_local
        P = C.c_void_p * nl

        def arr(f):
            return P(*[f(i) for i in range(nl)])
        nreads = (C.c_int64 * nl)(*[b.n_reads for b in batches])
        check(load()["xchg_place"](self._h, C.byref(cfg), nl,
                                   arr(lambda i: _abi.ptr(batches[i].seq)), arr(lambda i: _abi.ptr(batches[i].seq_off)), nreads,
                                   arr(lambda i: _abi.ptr(outs[i]["n_rows"])), arr(lambda i: _abi.ptr(outs[i]["node"])),
                                   arr(lambda i: _abi.ptr(outs[i]["score"])), arr(lambda i: _abi.ptr(outs[i]["lwr"])),
                                   arr(lambda i: _abi.ptr(outs[i]["counts"])) if counts else None,
                                   arr(lambda i: _abi.ptr(outs[i]["status"]))))
        return outs

    def stats(self):
        ms, pr, pb, hi, po = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(load()["xchg_stats"](self._h, C.byref(ms), C.byref(pr), C.byref(pb), C.byref(hi), C.byref(po)))
        return {"device_ms": ms.value, "probes": pr.value, "payload_bytes": pb.value, "owner_hits": hi.value,
                "owner_postings": po.value}
