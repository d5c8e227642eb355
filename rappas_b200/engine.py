"""Host-side handle on the GPU placement engine (thin ctypes layer over the C ABI).

`Database.place()` is the batch equivalent of the per-read body of
PlacementProcess.processQueries (core/algos/PlacementProcess.java:645-838, 974-1000); argument
names follow ArgumentsParser_v2 (keep_at_most, keep_factor, ...).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, load
from .synth import ReadBatch


class Database:
    def __init__(self, handle, desc):
        self._h = handle
        self.desc = desc

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, alphabet, k, n_nodes, thr_lin, thr_log10, keys, offsets, post_node, post_score,
                    devices=(0,), partitioned=False):
        fn = load()
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        post_node = np.ascontiguousarray(post_node, dtype=np.uint16)
        post_score = np.ascontiguousarray(post_score, dtype=np.float32)
        desc = _abi.RpDbDesc(int(alphabet), int(k), int(n_nodes), float(thr_log10), float(thr_lin), 0,
                             keys.shape[0], post_node.shape[0])
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        check(fn["db_load"](C.byref(desc), _abi.ptr(keys), _abi.ptr(offsets), _abi.ptr(post_node),
                            _abi.ptr(post_score), _abi.ptr(dev), dev.shape[0], int(partitioned), C.byref(h)))
        return cls(h, desc)

    @classmethod
    def from_synth(cls, db, devices=(0,), partitioned=False):
        """partitioned=True (1): every entry of `devices` holds one hash partition of the DB (an entry may repeat a
        device); the kernels of every distinct device reach all partitions through peer-mapped memory.
        partitioned=2: only the posting blocks are partitioned; every partition's device holds the whole table,
        so probes stay local and only the bulk gathers cross NVLink."""
        return cls.from_arrays(db.alphabet, db.k, db.n_nodes, db.thr_lin, db.thr_log10, db.keys, db.offsets,
                               db.post_node, db.post_score, devices=devices, partitioned=partitioned)

    @classmethod
    def from_synth_partitioned_dist(cls, db, device, group=None, replicate_table=False):
        """One process per GPU: this rank uploads only its hash partition of `db` to `device`, the ranks
        all_gather the CUDA-IPC blobs over torch.distributed, and every rank maps the others' partitions
        (peer memory over NVLink).  Any backend works for the 152-byte exchange."""
        import torch
        import torch.distributed as dist
        fn = load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        keys = np.ascontiguousarray(db.keys, dtype=np.uint64)
        offsets = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        post_node = np.ascontiguousarray(db.post_node, dtype=np.uint16)
        post_score = np.ascontiguousarray(db.post_score, dtype=np.float32)
        desc = _abi.RpDbDesc(int(db.alphabet), int(db.k), int(db.n_nodes), float(db.thr_log10), float(db.thr_lin), 0,
                             keys.shape[0], post_node.shape[0])
        blob = np.zeros(_abi.RP_PART_BLOB_BYTES, np.uint8)
        h = C.c_void_p()
        check(fn["db_load_partition"](C.byref(desc), _abi.ptr(keys), _abi.ptr(offsets), _abi.ptr(post_node),
                                      _abi.ptr(post_score), int(device), rank, world, int(bool(replicate_table)), _abi.ptr(blob),
                                      C.byref(h)))
        self = cls(h, desc)
        mine = torch.from_numpy(blob)
        if dist.get_backend(group) == "nccl":
            mine = mine.to(torch.device("cuda", int(device)))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        blobs = np.ascontiguousarray(torch.stack(gathered).cpu().numpy())
        check(fn["db_attach_partitions"](self._h, _abi.ptr(blobs), world))
        return self

    @classmethod
    def partition_of_synth(cls, db, device, part, n_parts):
        """Only hash partition `part` of `db`, on `device` (host-built): a rank of the exchange form
        (rappas_b200.exchange) or, after attach, of the peer-memory form."""
        fn = load()
        keys = np.ascontiguousarray(db.keys, dtype=np.uint64)
        offsets = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        post_node = np.ascontiguousarray(db.post_node, dtype=np.uint16)
        post_score = np.ascontiguousarray(db.post_score, dtype=np.float32)
        desc = _abi.RpDbDesc(int(db.alphabet), int(db.k), int(db.n_nodes), float(db.thr_log10), float(db.thr_lin), 0,
                             keys.shape[0], post_node.shape[0])
        blob = np.zeros(_abi.RP_PART_BLOB_BYTES, np.uint8)
        h = C.c_void_p()
        check(fn["db_load_partition"](C.byref(desc), _abi.ptr(keys), _abi.ptr(offsets), _abi.ptr(post_node),
                                      _abi.ptr(post_score), int(device), int(part), int(n_parts), 0, _abi.ptr(blob), C.byref(h)))
        return cls(h, desc)

    @classmethod
    def from_hash_db(cls, hdb, device=0, part=0, n_parts=1):
        """The hash-defined synthetic DB (rappas_b200.synth_hash.HashDB) generated ON THE DEVICE: the whole DB
        (n_parts == 1, ready for place()) or one partition of it."""
        fn = load()
        desc = _abi.RpDbDesc(int(hdb.alphabet), int(hdb.k), int(hdb.n_nodes), float(hdb.thr_log10), float(hdb.thr_lin), 0, 0, 0)
        plen = np.ascontiguousarray(hdb.plen, dtype=np.uint16)
        h = C.c_void_p()
        check(fn["db_synth_partition"](C.byref(desc), int(hdb.seed), float(hdb.occupancy), _abi.ptr(plen), int(device),
                                       int(part), int(n_parts), C.byref(h)))
        out = _abi.RpDbDesc()
        check(fn["db_describe"](h, C.byref(out)))
        return cls(h, out)

    def attach_partitions_dist(self, device, group=None):
        """Peer-memory form: all_gather the CUDA-IPC blobs of the ranks' partitions and map the others'."""
        import torch
        import torch.distributed as dist
        fn = load()
        world = dist.get_world_size(group)
        blob = np.zeros(_abi.RP_PART_BLOB_BYTES, np.uint8)
        check(fn["db_partition_blob"](self._h, _abi.ptr(blob)))
        mine = torch.from_numpy(blob)
        if dist.get_backend(group) == "nccl":
            mine = mine.to(torch.device("cuda", int(device)))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        blobs = np.ascontiguousarray(torch.stack(gathered).cpu().numpy())
        check(fn["db_attach_partitions"](self._h, _abi.ptr(blobs), world))
        return self

    @classmethod
    def from_file(cls, path, devices=(0,), partitioned=False):
        fn = load()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        check(fn["db_load_file"](str(path).encode(), _abi.ptr(dev), dev.shape[0], int(partitioned), C.byref(h)))
        desc = _abi.RpDbDesc()
        check(fn["db_describe"](h, C.byref(desc)))
        return cls(h, desc)

    def close(self):
        if self._h:
            load()["db_free"](self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- introspection ------------------------------------------------------------------------
    @property
    def k(self):
        return self.desc.k

    @property
    def n_nodes(self):
        return self.desc.n_nodes

    def device_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(load()["db_device_bytes"](self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_kernel_ms(self):
        return float(load()["last_kernel_ms"](self._h))

    # -- hot path -----------------------------------------------------------------------------
    def place(self, reads: ReadBatch, cfg=None, out=None, counts=True):
        """-> dict(n_rows, node, score, lwr, counts, status) of numpy arrays (host)."""
        cfg = cfg or _abi.place_cfg()
        n, K = reads.n_reads, cfg.keep_at_most
        if out is None:
            out = {
                "n_rows": np.empty(n, np.int32), "node": np.empty((n, K), np.uint16),
                "score": np.empty((n, K), np.float32), "lwr": np.empty((n, K), np.float64),
                "counts": np.empty((n, 4), np.int32) if counts else None, "status": np.empty(n, np.int32),
            }
        check(load()["place_batch"](self._h, C.byref(cfg), _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), n,
                                    _abi.ptr(out["n_rows"]), _abi.ptr(out["node"]), _abi.ptr(out["score"]),
                                    _abi.ptr(out["lwr"]), _abi.ptr(out.get("counts")), _abi.ptr(out["status"])))
        return out

    def place_device(self, cfg, d_seq, d_seq_off, n_reads, d_n_rows, d_node, d_score, d_lwr, d_counts, d_status,
                     stream=0, device_index=0):
        """All arguments are raw device pointers (ints); enqueues on `stream` and returns."""
        check(load()["place_batch_device"](self._h, int(device_index), C.byref(cfg), C.c_void_p(d_seq),
                                           C.c_void_p(d_seq_off), int(n_reads), C.c_void_p(d_n_rows),
                                           C.c_void_p(d_node), C.c_void_p(d_score), C.c_void_p(d_lwr),
                                           C.c_void_p(d_counts) if d_counts else None, C.c_void_p(d_status),
                                           C.c_void_p(stream) if stream else None))

    def extract(self, reads: ReadBatch):
        woff = reads.window_offsets(self.k)
        nw = int(woff[-1])
        out = {"win_off": woff, "code": np.zeros(nw, np.uint64), "kind": np.zeros(nw, np.uint8),
               "nalt": np.zeros(nw, np.int32), "hits": np.zeros(nw, np.int32),
               "status": np.zeros(reads.n_reads, np.int32)}
        check(load()["extract_kmers"](self._h, _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), reads.n_reads,
                                      _abi.ptr(woff), _abi.ptr(out["code"]), _abi.ptr(out["kind"]),
                                      _abi.ptr(out["nalt"]), _abi.ptr(out["hits"]), _abi.ptr(out["status"])))
        return out

    def place_windows(self, reads: ReadBatch, cfg=None):
        """K1 + K2 as the placement kernel's producer computes them: per window (code, hits); see rp_place_windows."""
        cfg = cfg or _abi.place_cfg()
        woff = reads.window_offsets(self.k)
        nw = int(woff[-1])
        code, hits = np.zeros(nw, np.uint64), np.zeros(nw, np.int32)
        check(load()["place_windows"](self._h, C.byref(cfg), _abi.ptr(reads.seq), _abi.ptr(reads.seq_off), reads.n_reads,
                                      _abi.ptr(woff), _abi.ptr(code), _abi.ptr(hits)))
        return {"win_off": woff, "code": code, "hits": hits}

    def node_scores(self, reads: ReadBatch, cfg=None):
        cfg = cfg or _abi.place_cfg()
        S = np.empty((reads.n_reads, self.n_nodes), np.float32)
        check(load()["node_scores"](self._h, C.byref(cfg), _abi.ptr(reads.seq), _abi.ptr(reads.seq_off),
                                    reads.n_reads, _abi.ptr(S), None))
        return S


def partition_of_keys(alphabet, k, keys, n_parts) -> np.ndarray:
    """Owner partition of each k-mer code (host helper, no GPU needed)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(keys.shape[0], np.int32)
    check(load()["partition_of_keys"](int(alphabet), int(k), _abi.ptr(keys), keys.shape[0], int(n_parts), _abi.ptr(out)))
    return out


def kernel_launch_count() -> int:
    return int(load()["kernel_launch_count"]())


def device_count() -> int:
    return int(load()["device_count"]())
