"""SURVEY.md section 8f row N1, reverse direction: the UNMODIFIED reference loads directories that
`dewi_b200.DewiIndex.save` wrote on a B200 (committed fixtures, scripts/make_cuda_saved_fixture.py) and returns
the results the CUDA backend returned.  `DewiIndex.load` resolves the backend class by name and falls back to
`ExactIndex.load` for the unknown name "CudaIndex" (src/dewi/index.py:148-150).  Needs the reference
(baseline/_ref or /root/reference/src); skipped where neither exists (the GPU box has baseline/_ref)."""

import logging
import sys
from pathlib import Path

import numpy as np
import pytest

from _util import GOLD, check_topk

ROOT = Path(__file__).resolve().parent.parent


def _reference():
    for p in (ROOT / "baseline" / "_ref", Path("/root/reference/src")):
        if (p / "dewi" / "index.py").exists():
            if str(p) not in sys.path:
                sys.path.insert(0, str(p))
            logging.getLogger("dewi.backends").setLevel(logging.ERROR)
            from dewi.index import DewiIndex

            return DewiIndex
    return None


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reference_loads_a_directory_saved_by_the_cuda_backend(dtype):
    ref_index = _reference()
    if ref_index is None:
        pytest.skip("the reference is not installed here (scripts/vendor_reference.sh)")
    path = GOLD / f"cuda_saved_index_{dtype}"
    if not path.exists():
        pytest.skip("fixture not generated yet")
    g = np.load(GOLD / f"cuda_saved_index_{dtype}_queries.npz")
    ix = ref_index.load(path)
    assert type(ix._backend).__name__ == "ExactIndex" and len(ix) == 48
    assert ix.rerank_eta == 0.3 and ix.entropy_pref == 0.5 and ix.get_metadata("doc_007") == {"source": "file_7.txt"}
    assert ix.get_payload("doc_003") is not None
    for q, ids, scores in zip(g["queries"], g["ids"], g["scores"]):
        res = ix.search(q, k=5)  # eta / entropy_pref from the saved config (index.py:86-89)
        check_topk(np.array([int(i[4:]) for i in ids]), scores, [int(r[0][4:]) for r in res], [r[1] for r in res])
