"""CPU: the product library builds, loads, and exports every symbol include/rappas_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rappas_b200 import _abi, _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rappas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    so = build.build()
    lib = C.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    # every declared function has a ctypes prototype, and vice versa
    assert sorted("rp_" + k for k in _abi.PROTOTYPES) == names


def test_pure_host_entry_points_work_without_a_gpu():
    fn = _lib.load()
    a, b = C.c_float(), C.c_float()
    fn["threshold"](1.5, 0, 8, C.byref(a), C.byref(b))
    assert np.float32(b.value).view(np.uint32) == 0xC05A1893
    st = np.array([1, 3, 0, 2, 1, 1, 3, 0], np.uint8)
    assert fn["pack_kmer"](0, _abi.ptr(st), 8) == 0x358D
    assert b"sm_100a" in fn["version"]()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rappas_b200 as R
    from rappas_b200 import synth
    db = synth.make_db(0, 5, 17, n_keys=100, mean_postings=3, seed=1)
    with pytest.raises(_lib.RappasError) as e:
        R.Database.from_synth(db)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rappas_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("oracle_lib", "oracle_py", "librappas_oracle", "rappas_oracle.h", "rpo_place",
                               "rpo_db", "import oracle", "from oracle"):
                    assert needle not in text, (f, needle)
