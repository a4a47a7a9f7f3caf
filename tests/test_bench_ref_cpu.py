"""The CPU legs of bench.py (bench_ref.py): the harness around the unmodified reference must execute the reference's own
`search` body on a bulk-filled index and agree with the oracle port; the BLAS pool is pinned explicitly."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench_ref  # noqa: E402


def test_blas_pool_is_pinned_not_inherited(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")   # what torch.distributed.run exports
    want = min(bench_ref.host_threads(), 4)
    assert bench_ref.pin_blas(want) == want
    assert bench_ref.host_threads(3) == 3 and 1 <= bench_ref.host_threads() <= bench_ref.DEFAULT_MAX_THREADS


def test_ann_baselines_are_attempted_and_recorded():
    rec = bench_ref.ann_baselines()
    assert set(rec) == {"hnswlib", "faiss"}
    assert all(v == "importable" or v.startswith("unavailable -- not installed") for v in rec.values())


def test_bulk_filled_reference_index_runs_the_references_search_and_matches_the_port():
    ref = bench_ref.load_reference()
    if ref is None:
        pytest.skip("baseline/_ref not installed (scripts/vendor_reference.sh)")
    emb = bench_ref.gen_corpus(np, 5000, 64, threads=2)
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, rtol=1e-5)
    s_ref = bench_ref.ExactSearcher(np, emb, 10, 0.3, 0.5, ref)
    s_port = bench_ref.ExactSearcher(np, emb, 10, 0.3, 0.5, None)
    assert s_ref.kind == "reference" and s_port.kind == "port"
    assert type(s_ref.index._backend).__name__ == "ExactIndex" and s_ref.index._backend._embeddings is emb
    q = np.random.RandomState(3).standard_normal((4, 64)).astype(np.float32)
    for x in q:
        got = s_ref.search(x)
        ids, sc = s_port.search(x)
        assert [int(r[0][4:]) for r in got] == ids.tolist()
        np.testing.assert_array_equal(np.array([r[1] for r in got], np.float32), sc)
        assert got[0][2].dewi == float(s_ref.dewi[ids[0]])     # lazy payloads carry the row's columns
    ms = bench_ref.time_queries(np, s_ref, q, 1, 3)
    assert ms.shape == (3,) and np.all(ms > 0)


def test_largest_rows_respects_host_memory():
    n = bench_ref.largest_rows(768, 10_000_000)
    assert n % 1_000_000 == 0 and 1_000_000 <= n <= 10_000_000
