"""CPU: the C-ABI shared library loads, exports every symbol include/sequnwinder_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sequnwinder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sequnwinder_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sqw_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 18
    lib = _lib.load()
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == names        # the Python binding covers the whole header


def test_abi_version_and_kmer_sizes():
    lib = _lib.load()
    assert lib.sqw_abi_version() == 2
    assert lib.sqw_num_kmers(4, 5) == 1280
    assert lib.sqw_num_kmers(4, 6) == 5376
    assert lib.sqw_num_kmers(4, 7) == 21760
    assert lib.sqw_num_kmers(5, 4) == -1


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.sqw_device_count() > 0:
        pytest.skip("a CUDA device is present")
    bases = np.frombuffer(b"ACGTACGT", dtype=np.uint8).copy()
    offs = np.array([0, 8], dtype=np.int64)
    counts = np.zeros((1, 1280), dtype=np.int32)
    has_n = np.zeros(1, dtype=np.uint8)
    rc = lib.sqw_kmer_count(bases.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(_lib.i64p), 1, 4, 5,
                            counts.ctypes.data_as(_lib.i32p), has_n.ctypes.data_as(_lib.u8p))
    assert rc == _lib.SQW_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.sqw_last_error()
    h = C.c_void_p()
    assert lib.sqw_admm_create(C.byref(h)) == _lib.SQW_ERR_NO_DEVICE
    assert counts.sum() == 0


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "sequnwinder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src, f
