"""pytest configuration: the `gpu` marker, import paths and shared helpers.

`-m "not gpu"` covers the oracle against the golden fixtures, host-side logic and the C-ABI symbol
table; `-m gpu` holds the parity tests proper, which call libdewi_b200.so on a B200.
"""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a machine without a GPU: skip instead of failing in every constructor
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def lib_path() -> Path:
    """libdewi_b200.so, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry

    return entry.build_library()
