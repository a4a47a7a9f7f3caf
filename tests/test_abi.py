"""The C-ABI library builds, loads and exports every symbol include/dewi_b200.h declares.
No compute call is made here (no GPU in the build container)."""

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "dewi_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return re.findall(r"DEWI_API\s+[\w\s\*]+?\b(dewi_\w+)\s*\(", text)


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("dewi_index_create", "dewi_index_append", "dewi_index_search", "dewi_index_search_local", "dewi_rerank",
                 "dewi_fit_stats", "dewi_score", "dewi_similarity_dense", "dewi_join", "dewi_last_error"):
        assert must in syms
    assert len(syms) == len(set(syms))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_table_matches_header(lib_path):
    import dewi_b200._native as native

    assert set(native.SIGNATURES) == set(declared_symbols())
    lib = native.load_library()
    assert lib.dewi_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu(lib_path):
    import torch

    import dewi_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ImportError):
        dewi_b200.DewiIndex(dim=8)
    with pytest.raises(ImportError):
        dewi_b200.DewiScorer().fit_stats([{"ht_mean": 1.0}])
    with pytest.raises(ImportError):
        dewi_b200.cross_modal_similarity([[1.0, 0.0]], [[1.0, 0.0]])


def test_product_never_imports_the_oracle():
    pkg = ROOT / "dewi-design-for-an-entropy-weighted-index-for-text-image-corpora_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
        assert "/root/reference" not in text
