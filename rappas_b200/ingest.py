"""Host side either end of the hot path: FASTA ingest + duplicate structure, and the .jplace writer
(thin ctypes layer over rp_reads_* / rp_jplace_write; the work is C++ in csrc/rp_ingest.cpp).

Mirrors inputs/FASTAPointer.java:67-149 (records), core/algos/PlacementProcess.java:591-629 (duplicates)
and :974-1047 + main_v2/Main_PLACEMENT_v07.java:270-315 (jplace).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, load
from .synth import ReadBatch


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class QueryFile:
    """A parsed query FASTA: `.unique` (ReadBatch of the distinct exact sequences -- what is placed),
    `.headers`, `.unique_of[record]`, `.group_of[record]` (gap-stripped duplicate key)."""

    def __init__(self, handle):
        self._h = handle
        fn = load()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(fn["reads_describe"](self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_records, self.n_unique, self.n_groups = a.value, b.value, c.value
        seq, off = C.c_void_p(), C.c_void_p()
        check(fn["reads_unique"](self._h, C.byref(seq), C.byref(off)))
        seq_off = _view(off, self.n_unique + 1, np.uint64)
        self.unique = ReadBatch(_view(seq, int(seq_off[-1]), np.uint8), seq_off)
        hdr, hoff, uo, go = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(fn["reads_records"](self._h, C.byref(hdr), C.byref(hoff), C.byref(uo), C.byref(go)))
        self._hdr_off = _view(hoff, self.n_records + 1, np.uint64)
        self._hdr_raw = _view(hdr, int(self._hdr_off[-1]), np.uint8).tobytes()
        self._headers = None
        self.unique_of = _view(uo, self.n_records, np.uint32)
        self.group_of = _view(go, self.n_records, np.uint32)

    @property
    def headers(self):
        """Header strings of all records (built on first use: a million Python strings take longer than the parse)."""
        if self._headers is None:
            o, raw = self._hdr_off, self._hdr_raw
            self._headers = [raw[int(o[i]):int(o[i + 1])].decode("latin-1") for i in range(self.n_records)]
        return self._headers

    def header(self, i):
        return self._hdr_raw[int(self._hdr_off[i]):int(self._hdr_off[i + 1])].decode("latin-1")

    @classmethod
    def from_file(cls, path):
        h = C.c_void_p()
        check(load()["reads_load_fasta"](str(path).encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def from_text(cls, text):
        data = text.encode("latin-1") if isinstance(text, str) else bytes(text)
        buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        check(load()["reads_from_memory"](_abi.ptr(buf) if buf.size else None, buf.size, C.byref(h)))
        return cls(h)

    def close(self):
        if self._h:
            load()["reads_free"](self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def write_jplace(self, path, results, keep_at_most, edge_id, branch_len, tree_newick=None, invocation="",
                     guppy_compat=False, not_placed_path=None):
        """results: the dict Database.place(self.unique, cfg) returned.  -> number of placements written."""
        edge_id = np.ascontiguousarray(edge_id, dtype=np.int32)
        branch_len = np.ascontiguousarray(branch_len, dtype=np.float32)
        n = C.c_uint64()
        check(load()["jplace_write"](
            str(path).encode(), self._h, int(keep_at_most), _abi.ptr(np.ascontiguousarray(results["n_rows"], np.int32)),
            _abi.ptr(np.ascontiguousarray(results["node"], np.uint16)), _abi.ptr(np.ascontiguousarray(results["score"], np.float32)),
            _abi.ptr(np.ascontiguousarray(results["lwr"], np.float64)), _abi.ptr(np.ascontiguousarray(results["status"], np.int32)),
            _abi.ptr(edge_id), _abi.ptr(branch_len), edge_id.shape[0],
            tree_newick.encode() if tree_newick is not None else None, invocation.encode(), int(bool(guppy_compat)),
            str(not_placed_path).encode() if not_placed_path else None, C.byref(n)))
        return n.value


def java_number(v, as_float=False) -> str:
    """Float.toString / Double.toString of v as the jplace writer prints it."""
    buf = C.create_string_buffer(64)
    check(load()["java_number"](float(v), int(bool(as_float)), buf, 64))
    return buf.value.decode()
