"""CPU tests: the C-ABI library builds, loads and exports exactly what
include/vdb_b200.h declares; argument validation works without a GPU; the
product package never touches the oracle."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")


def header_symbols(name="vdb_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg.build()
    lib = C.CDLL(pkg.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in vdb_b200.h but not exported"
    assert sorted(pkg.ABI) == syms, "python binding table and header disagree"


def test_invalid_config_is_invalid_argument():
    # ctor contract, ivf_flat_index.cpp:17-19 (checked before any device is touched)
    with pytest.raises(ValueError):
        pkg.IVFFlatIndex(pkg.Config(dimension=0, nlist=4))
    with pytest.raises(ValueError):
        pkg.IVFFlatIndex(pkg.Config(dimension=4, nlist=0))
    with pytest.raises(ValueError):
        pkg.IVFFlatIndex(pkg.Config(dimension=4, nlist=4, metric=pkg.Metric.Cosine))


def test_storage_library_exports_every_declared_symbol():
    storage = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.storage")
    pkg.build()
    lib = C.CDLL(storage.STORAGE_LIB_PATH)
    syms = header_symbols("vdb_b200_storage.h")
    assert len(syms) >= 6
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in vdb_b200_storage.h but not exported"
    assert sorted(storage.STORAGE_ABI) == syms, "python binding table and storage header disagree"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.VdbError) as e:
        pkg.IVFFlatIndex(pkg.Config(dimension=8, nlist=2))
    assert e.value.status == pkg.VDB_CUDA_ERROR


def test_product_never_references_oracle():
    pdir = os.path.join(ROOT, "cuda-acceleratedvectordatabaseengine_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "libvdbref" not in txt, f
