"""CPU: the C-ABI shared library loads and exports every symbol include/b2r.h declares; the ctypes
binding lists exactly those symbols; without a GPU the entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2r.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rag_challenge_2_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run ./build.sh first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"libb2r.so does not export {name}"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert _lib.load().b2r_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rag_challenge_2_b200 as pkg
    from rag_challenge_2_b200 import _lib
    from rag_challenge_2_b200.bm25 import BM25Okapi

    with pytest.raises(_lib.B2RError):
        pkg.FlatIPIndex(1024)
    with pytest.raises(_lib.B2RError):
        BM25Okapi([["a", "b"], ["c"]])
    D = np.zeros((1, 1), np.float32)
    I = np.zeros((1, 1), np.int64)
    with pytest.raises(_lib.B2RError):
        _lib.check(_lib.load().b2r_topk_merge_host(_lib.ptr(D), _lib.ptr(I), 1, 1, 1, 1, _lib.ptr(D), _lib.ptr(I), 0))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "rag-challenge-2_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_tunables_round_trip():
    from rag_challenge_2_b200 import _lib

    _lib.tunable_set("flat.stages", 3)
    assert _lib.tunable_get("flat.stages") == 3
    _lib.tunable_set("flat.stages", 0)
    with pytest.raises(_lib.B2RError):
        _lib.tunable_set("no.such.key", 1)


def test_faiss_file_round_trip_without_gpu(tmp_path):
    """IxFI reader/writer (host-side parsing only)."""
    import struct

    from rag_challenge_2_b200 import flat_index

    rows = np.random.default_rng(0).standard_normal((7, 16)).astype(np.float32)
    path = tmp_path / "x.faiss"
    with open(path, "wb") as f:
        f.write(struct.pack("<4siqqqBi", b"IxFI", 16, 7, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", rows.size))
        f.write(rows.tobytes())
    d, got = flat_index.read_index_rows(path)
    assert d == 16 and np.array_equal(got, rows)
    with open(path, "r+b") as f:
        f.write(b"IxF2")
    with pytest.raises(RuntimeError):
        flat_index.read_index_rows(path)
