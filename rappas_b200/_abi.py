"""ctypes view of include/rappas_b200.h (structs + prototypes).

`bind(lib, prefix)` attaches argtypes/restypes to a loaded library whose symbols carry
`prefix` ("rp_" for the CUDA product library; the test-suite reuses the same prototypes with
"rpo_" for the CPU oracle, which mirrors the ABI one to one).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

RP_OK = 0
RP_MAX_KEEP = 32
RP_PART_BLOB_BYTES = 152
STATUS_PLACED, STATUS_UNPLACED, STATUS_TOO_SHORT, STATUS_BAD_CHAR, STATUS_TOO_LONG = 0, 1, 2, 3, 4
RP_XCHG_ID_BYTES = 128
WIN_PLAIN, WIN_AMBIG, WIN_SKIPPED = 0, 1, 2
CNT_WINDOWS, CNT_MATCHED, CNT_AMBIG, CNT_SKIPPED = 0, 1, 2, 3


class RpDbDesc(C.Structure):
    _fields_ = [
        ("alphabet", C.c_int32), ("k", C.c_int32), ("n_nodes", C.c_int32),
        ("thr_log10", C.c_float), ("thr_lin", C.c_float), ("reserved0", C.c_int32),
        ("n_keys", C.c_uint64), ("n_postings", C.c_uint64),
    ]


class RpPlaceCfg(C.Structure):
    _fields_ = [
        ("keep_at_most", C.c_int32), ("keep_factor", C.c_float), ("treat_amb", C.c_int32),
        ("amb_with_max", C.c_int32), ("ns_bound", C.c_float), ("reserved0", C.c_int32),
    ]


class RpDbBuildDesc(C.Structure):
    _fields_ = [
        ("alphabet", C.c_int32), ("k", C.c_int32), ("n_nodes", C.c_int32), ("n_sites", C.c_int32),
        ("n_states", C.c_int32), ("thr_log10", C.c_float), ("gap_jumps", C.c_int32), ("reserved0", C.c_int32),
    ]


def place_cfg(keep_at_most=7, keep_factor=0.01, treat_amb=True, amb_with_max=False, ns_bound=-np.inf) -> RpPlaceCfg:
    """Defaults = ArgumentsParser_v2.java:86-91."""
    return RpPlaceCfg(int(keep_at_most), float(keep_factor), int(bool(treat_amb)), int(bool(amb_with_max)),
                      float(ns_bound), 0)


_P = C.c_void_p

# name (without prefix) -> (restype, argtypes)
PROTOTYPES = {
    "threshold": (None, [C.c_float, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pack_kmer": (C.c_uint64, [C.c_int32, _P, C.c_int32]),
    "db_load": (C.c_int, [C.POINTER(RpDbDesc), _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "db_load_file": (C.c_int, [C.c_char_p, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "db_save_file": (C.c_int, [C.c_char_p, C.POINTER(RpDbDesc), _P, _P, _P, _P]),
    "db_load_partition": (C.c_int, [C.POINTER(RpDbDesc), _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "db_attach_partitions": (C.c_int, [_P, _P, C.c_int32]),
    "partition_of_keys": (C.c_int, [C.c_int32, C.c_int32, _P, C.c_uint64, C.c_int32, _P]),
    "db_synth_partition": (C.c_int, [C.POINTER(RpDbDesc), C.c_uint64, C.c_double, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "db_partition_blob": (C.c_int, [_P, _P]),
    "xchg_unique_id": (C.c_int, [_P]),
    "xchg_create": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "xchg_create_local": (C.c_int, [_P, C.c_int32, C.POINTER(_P)]),
    "xchg_place": (C.c_int, [_P, C.POINTER(RpPlaceCfg), C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "xchg_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                         C.POINTER(C.c_uint64)]),
    "xchg_free": (None, [_P]),
    "xchg_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "db_free": (None, [_P]),
    "gap_intervals": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pp_prepare": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "dbbuild_run": (C.c_int, [C.POINTER(RpDbBuildDesc), _P, _P, _P, _P, _P, C.c_int32, C.POINTER(_P)]),
    "dbbuild_result": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(_P),
                                 C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_double)]),
    "dbbuild_free": (None, [_P]),
    "db_describe": (C.c_int, [_P, C.POINTER(RpDbDesc)]),
    "db_device_bytes": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "host_alloc": (C.c_int, [C.POINTER(_P), C.c_uint64]),
    "host_free": (None, [_P]),
    "host_register": (C.c_int, [_P, C.c_uint64]),
    "host_unregister": (C.c_int, [_P]),
    "place_batch": (C.c_int, [_P, C.POINTER(RpPlaceCfg), _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "place_batch_device": (C.c_int, [_P, C.c_int32, C.POINTER(RpPlaceCfg), _P, _P, C.c_int64,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "extract_kmers": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "place_windows": (C.c_int, [_P, C.POINTER(RpPlaceCfg), _P, _P, C.c_int64, _P, _P, _P]),
    "node_scores": (C.c_int, [_P, C.POINTER(RpPlaceCfg), _P, _P, C.c_int64, _P, _P]),
    "reads_load_fasta": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "reads_from_memory": (C.c_int, [_P, C.c_uint64, C.POINTER(_P)]),
    "reads_free": (None, [_P]),
    "reads_describe": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "reads_unique": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "reads_records": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "jplace_write": (C.c_int, [C.c_char_p, _P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_char_p,
                               C.c_char_p, C.c_int32, C.c_char_p, C.POINTER(C.c_uint64)]),
    "java_number": (C.c_int, [C.c_double, C.c_int32, C.c_char_p, C.c_int32]),
    "device_count": (C.c_int, []),
    "kernel_launch_count": (C.c_uint64, []),
    "last_kernel_ms": (C.c_double, [_P]),
    "version": (C.c_char_p, []),
    "last_error": (C.c_char_p, []),
}

# oracle-only extras (rpo_ prefix): db_load has no device arguments, plus the threaded variant
ORACLE_OVERRIDES = {
    "db_load": (C.c_int, [C.POINTER(RpDbDesc), _P, _P, _P, _P, C.POINTER(_P)]),
    "place_batch_mt": (C.c_int, [_P, C.POINTER(RpPlaceCfg), _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int32]),
    "max_ambig_per_mer": (C.c_int32, [C.c_int32, C.c_int32]),
    "char_class": (C.c_int32, [C.c_int32, C.c_int32]),
    "ambiguity_equivalence": (C.c_int32, [C.c_int32, C.c_int32, _P]),
}


def bind(lib: C.CDLL, prefix: str, names=None, overrides=None, strict=True):
    protos = dict(PROTOTYPES)
    if overrides:
        protos.update(overrides)
    bound = {}
    for name, (res, args) in protos.items():
        if names is not None and name not in names:
            continue
        sym = prefix + name
        try:
            fn = getattr(lib, sym)
        except AttributeError:
            if strict:
                raise
            continue
        fn.restype = res
        fn.argtypes = args
        bound[name] = fn
    return bound


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
