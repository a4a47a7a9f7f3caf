"""The reference-side binding of INTEGRATION.md section 1, executed as written.

The stub is the ~55 lines a maintainer of the reference would paste into `src/dewi/backends.py`: a `BaseIndex`
subclass that talks to libdewi_b200.so through plain ctypes (no torch, nothing from this repository's Python
package).  The CPU test checks that the block compiles and only calls symbols `include/dewi_b200.h` declares; the
GPU test executes it inside the namespace of the UNMODIFIED reference's `dewi.backends` (baseline/_ref, or
/root/reference/src in the build container), drives it through the reference's own `DewiIndex` facade
(src/dewi/index.py:62-92) and compares every result tuple with what the reference's `ExactIndex`
(backends.py:386-481) returns for the same documents -- the drop-in claim, checked on the reference itself."""

import importlib
import logging
import os
import re
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import check_topk

ROOT = Path(__file__).resolve().parent.parent


def _stub_source() -> str:
    text = (ROOT / "INTEGRATION.md").read_text(encoding="utf-8")
    return re.search(r"```python\n(.*?)```", text, re.S).group(1)


def _reference_backends():
    for p in (ROOT / "baseline" / "_ref", Path("/root/reference/src")):
        if (p / "dewi" / "backends.py").exists():
            if str(p) not in sys.path:
                sys.path.insert(0, str(p))
            logging.getLogger("dewi.backends").setLevel(logging.ERROR)
            return importlib.import_module("dewi.backends"), importlib.import_module("dewi.index"), \
                importlib.import_module("dewi.types")
    return None


def test_stub_compiles_and_calls_only_declared_symbols():
    src = _stub_source()
    compile(src, "INTEGRATION.md", "exec")
    header = (ROOT / "include" / "dewi_b200.h").read_text(encoding="utf-8")
    used = set(re.findall(r"_dewi_b200\.(dewi_\w+)", src))
    assert {"dewi_index_create", "dewi_index_append", "dewi_index_set_payload", "dewi_index_search",
            "dewi_index_destroy", "dewi_device_check", "dewi_last_error"} <= used
    for sym in used:
        assert re.search(rf"\b{sym}\(", header), f"{sym} is not declared in include/dewi_b200.h"
    imports = set(re.findall(r"^\s*(?:import|from)\s+(\w+)", src, re.M))
    assert imports == {"ctypes"}, imports  # numpy / os / Enum come from the reference module's own imports


@pytest.mark.gpu
def test_stub_runs_inside_the_unmodified_reference(lib_path):
    ref = _reference_backends()
    if ref is None:
        pytest.skip("the reference is not installed here (scripts/vendor_reference.sh)")
    backends, index_mod, types_mod = ref
    os.environ["DEWI_B200_LIB"] = str(lib_path)
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), backends.__dict__)
    assert backends._HAS_B200
    cuda_cls = backends.CudaIndex
    assert issubclass(cuda_cls, backends.BaseIndex)

    rng = np.random.RandomState(7)
    n, dim, k = 3000, 768, 10
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    exact = index_mod.DewiIndex(dim=dim, use_ann=False, rerank_eta=0.3, entropy_pref=0.5)
    cuda = index_mod.DewiIndex(dim=dim, use_ann=False, rerank_eta=0.3, entropy_pref=0.5)
    cuda._backend = cuda_cls(dim, "cosine")  # the branch index.py:44-60 gains
    for i in range(n):
        p = types_mod.Payload(dewi=float(rng.uniform()), ht_mean=float(rng.uniform(0, 10)), hi_mean=float(rng.uniform(0, 5)))
        exact.add(f"doc_{i}", emb[i], p)
        cuda.add(f"doc_{i}", emb[i], p)
    exact.build()
    cuda.build()
    assert len(cuda) == len(exact) == n
    for eta, pref in ((None, None), (0.0, 0.0), (1.0, 0.0), (0.5, -0.25)):
        for q in rng.standard_normal((6, dim)).astype(np.float32):
            want = exact.search(q, k=k, eta=eta, entropy_pref=pref)
            got = cuda.search(q, k=k, eta=eta, entropy_pref=pref)
            check_topk([int(r[0][4:]) for r in want], [r[1] for r in want], [int(r[0][4:]) for r in got],
                       [r[1] for r in got], what=f"eta={eta} pref={pref}")
            assert all(g[2] is cuda.get_payload(g[0]) for g in got)  # the shared Payload objects, as ExactIndex returns
    # error behaviour of the facade and of the final select (index.py:91-92, backends.py:468)
    with pytest.raises(ValueError):
        cuda.search(emb[:2], k=k)
    with pytest.raises(ValueError):
        exact.search(emb[0], k=n + 5)
    with pytest.raises(ValueError):
        cuda.search(emb[0], k=n + 5)
