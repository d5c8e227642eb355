"""Loader of the CUDA product library (rappas_b200/librappas_b200.so).

There is no CPU fallback: if the library is missing it is built in-tree with nvcc; if that fails, or
if no CUDA device is visible when a DB is loaded, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "librappas_b200.so")
# development only (tools/sweep_variants.sh): another build of the same library, e.g. with a -D switch
if os.environ.get("RAPPAS_B200_LIB"):
    SO_PATH = os.path.abspath(os.environ["RAPPAS_B200_LIB"])
_fn = None
_lib = None


class RappasError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rappas_b200 error %d: %s" % (code, msg))
        self.code = code


def load(build_if_missing=True):
    global _fn, _lib
    if _fn is not None:
        return _fn
    if not os.path.exists(SO_PATH):
        if not build_if_missing:
            raise RappasError(-1, "librappas_b200.so not built (python -m rappas_b200.build)")
        from . import build as _build
        _build.build()
    _lib = C.CDLL(SO_PATH)
    _fn = _abi.bind(_lib, "rp_", strict=True)
    return _fn


def check(rc):
    if rc != 0:
        raise RappasError(rc, load()["last_error"]().decode(errors="replace"))
