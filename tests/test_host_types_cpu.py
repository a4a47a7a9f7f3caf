"""Host-side pieces of the drop-in API that need no GPU: the reference's own type tests re-run on our types
(tests/test_index.py:73-101), the README's `Signals` contract (README.md:67,83-110), backend-name
resolution (backends.py:39-49), lazy document ids, payload column mirrors, the packed exchange layout
and the cluster metrics (metrics.py:173-212)."""

import json

import numpy as np
import pytest

import dewi_b200
from dewi_b200.backends import ColumnPayloads, ColumnStore, IndexBackend, LazyIds
from dewi_b200.types import PAYLOAD_FIELDS, SIGNAL_FIELDS


def test_payload_serialization_like_the_reference():
    payload = dewi_b200.Payload(dewi=0.5, ht_mean=1.2, ht_q90=1.8, hi_mean=0.8, hi_q90=1.2, I_hat=0.6, redundancy=0.1, noise=0.05)
    d = payload.to_dict()
    assert isinstance(d, dict) and "dewi" in d and list(d) == list(PAYLOAD_FIELDS)
    assert dewi_b200.Payload.from_dict(d) == payload
    raw = payload.to_bytes()
    assert isinstance(raw, bytes) and json.loads(raw.decode("utf-8")) == d
    assert dewi_b200.Payload.from_bytes(raw) == payload
    # extras are ignored and values are cast to float (types.py:28-30); missing fields default to 0.0
    p = dewi_b200.Payload.from_dict({"dewi": 1, "unknown": 3, "noise": "0.25"})
    assert p.dewi == 1.0 and isinstance(p.dewi, float) and p.noise == 0.25 and p.ht_mean == 0.0
    assert dewi_b200.Payload() == dewi_b200.Payload(**{f: 0.0 for f in PAYLOAD_FIELDS})


def test_weights_defaults():
    w = dewi_b200.Weights()
    assert (w.alpha_t, w.alpha_i, w.alpha_m, w.alpha_r, w.alpha_n, w.delta) == (1.0, 1.0, 1.0, 1.0, 1.0, 3.0)  # types.py:42-51


def test_signals_is_a_dataclass_and_a_mapping():
    """README.md:83-110 builds `Signals(...)`, unpacks `signals.__dict__` into `Payload(**...)` and hands the
    objects to `fit_stats` / `score`, which index them as mappings (scorer.py:20-21,53-57)."""
    s = dewi_b200.Signals(ht_mean=1.0, ht_q90=2.0, hi_mean=3.0, hi_q90=4.0, I_hat=0.5, redundancy=0.25, noise=0.125)
    assert tuple(dewi_b200.Signals.__annotations__) == SIGNAL_FIELDS
    assert list(s.keys()) == list(SIGNAL_FIELDS) and s["hi_q90"] == 4.0 and "noise" in s and len(s) == 7
    assert dict(s.items()) == s.__dict__ and list(iter(s)) == list(SIGNAL_FIELDS)
    with pytest.raises(KeyError):
        s["dewi"]
    p = dewi_b200.Payload(**s.__dict__, dewi=0.75)  # README.md:109
    assert p.dewi == 0.75 and p.redundancy == 0.25


def test_backend_names():
    assert IndexBackend.from_str("cuda") is IndexBackend.CUDA and IndexBackend.from_str("auto") is IndexBackend.CUDA
    assert IndexBackend.from_str("exact") is IndexBackend.EXACT and IndexBackend.from_str("ExactIndex") is IndexBackend.EXACT
    assert IndexBackend.from_str("hnsw") is IndexBackend.HNSW
    with pytest.raises(KeyError):
        IndexBackend.from_str("annoy")


def test_lazy_ids():
    ids = LazyIds(100_000_000)
    assert len(ids) == 100_000_000 and ids[0] == "doc_00000000" and ids[-1] == "doc_99999999"  # profile_index.py:52
    assert ids[3:6] == ["doc_00000003", "doc_00000004", "doc_00000005"]
    assert ids.index("doc_00012345") == 12345
    for bad in ("doc_12345", "doc_100000000", "row_00000001", "doc_0000000x", None):
        with pytest.raises(ValueError):
            ids.index(bad)
    with pytest.raises(IndexError):
        ids[100_000_000]


def test_column_store_and_payload_view():
    rng = np.random.RandomState(0)
    a, b = rng.rand(5, 8).astype(np.float32), rng.rand(3, 8).astype(np.float32)
    store = ColumnStore()
    store.append(5, a)
    store.append(4, None)       # four default payloads
    store.append(3, b)
    assert store.n == 12 and not store.all_default()
    np.testing.assert_array_equal(store.row(2), a[2])
    np.testing.assert_array_equal(store.row(7), np.zeros(8, np.float32))
    np.testing.assert_array_equal(store.row(10), b[1])
    np.testing.assert_array_equal(store.column(0), np.concatenate([a[:, 0], np.zeros(4, np.float32), b[:, 0]]))
    view = ColumnPayloads(LazyIds(12), store)
    p = view["doc_00000010"]
    assert isinstance(p, dewi_b200.Payload) and p.ht_mean == float(b[1, 1]) and len(view) == 12
    assert "doc_00000011" in view and "doc_00000012" not in view and view.get("nope") is None
    assert view["doc_00000006"] == dewi_b200.Payload()
    with pytest.raises(KeyError):
        view["doc_99999999"]


def test_packed_candidate_layout():
    """One rank block = [id i64 | sim f32 | dewi f32 | ent f32] x (B * kcand), padded to 8 bytes: the layout the
    finalize kernel writes (locally or into the peers' buffers) and `dewi_rerank` reads shard-strided."""
    torch = pytest.importorskip("torch")
    from dewi_b200.sharded import PackedCandidates

    pk = PackedCandidates(b=3, kcand=5, device="cpu", world=4)
    assert pk.words % 2 == 0 and pk.words >= 5 * 15 and pk.stride_bytes == pk.words * 4
    assert pk.gathered.numel() == 4 * pk.words
    ids, sim, dewi, ent = pk.views(pk.local)
    assert ids.dtype == torch.int64 and ids.shape == (3, 5) and sim.shape == dewi.shape == ent.shape == (3, 5)
    ids.fill_(7)
    sim.fill_(1.5)
    dewi.fill_(2.5)
    ent.fill_(3.5)
    raw = pk.local.numpy()
    assert np.all(raw[:30].view(np.int64) == 7) and np.all(raw[30:45].view(np.float32) == 1.5)
    assert np.all(raw[45:60].view(np.float32) == 2.5) and np.all(raw[60:75].view(np.float32) == 3.5)


def test_cluster_metrics_like_the_reference():
    clusters = [["a"], ["b", "c"], ["d"], ["e", "f", "g"]]
    assert dewi_b200.duplicate_rate(clusters) == 0.5                      # metrics.py:173-192
    assert dewi_b200.duplicate_rate([]) == 0.0 and dewi_b200.duplicate_rate([[], []]) == 0.0
    assert dewi_b200.cluster_coverage(["c", "g", "zzz"], clusters) == 0.5  # metrics.py:194-212
    assert dewi_b200.cluster_coverage([], clusters) == 0.0 and dewi_b200.cluster_coverage(["a"], []) == 0.0
    labels = np.array([0, 1, 1, 3, 0, 5])
    assert dewi_b200.clusters_from_labels(labels) == [[0, 4], [1, 2], [3], [5]]
    assert dewi_b200.clusters_from_labels(labels, doc_ids=list("abcdef")) == [["a", "e"], ["b", "c"], ["d"], ["f"]]


def test_candidate_limit_of_the_wrappers():
    """At most 400 candidates (min(2k, N), backends.py:440) per query -- 392 when `rerank_scope="full"` over-fetches
    eight slots; small corpora are bounded by N, not by k (documented divergence)."""
    from types import SimpleNamespace

    from dewi_b200.backends import MAX_CANDIDATES, CudaIndex

    check = CudaIndex._check_candidate_limit
    assert MAX_CANDIDATES == 400
    big = SimpleNamespace(_n_device=1_000_000, rerank_scope="candidates")
    check(big, 10)
    check(big, 200)
    with pytest.raises(ValueError, match="400 candidates"):
        check(big, 201)
    check(SimpleNamespace(_n_device=300, rerank_scope="candidates"), 300)      # k = N = 300: every row is a candidate
    full = SimpleNamespace(_n_device=1_000_000, rerank_scope="full")
    check(full, 196)
    with pytest.raises(ValueError, match="392 candidates"):
        check(full, 197)
    check(SimpleNamespace(_n_device=350, rerank_scope="full"), 350)


def test_fit_columns_splits_wide_inputs_over_several_library_calls(monkeypatch):
    """`RobustStats.fit` fits every key a row has (scorer.py:19-25); one dewi_fit_stats call takes 32 columns.  The
    wrapper's splitting (pointer offsets, result assembly) is checked against a stand-in for the library call that
    reads the very pointers it is handed -- the kernel itself is covered by the GPU tests."""
    import ctypes
    from types import SimpleNamespace

    import torch

    from dewi_b200 import _native, scorer

    calls = []

    def fake_fit_stats(ptr, n, f, ld, med, mad, device, stream):
        assert 1 <= f <= scorer.MAX_FIT_COLUMNS
        calls.append(f)
        flat = np.ctypeslib.as_array((ctypes.c_float * ((f - 1) * ld + n)).from_address(ptr.value))
        for c in range(f):
            v = flat[c * ld: c * ld + n]
            med[c] = float(np.median(v))
            mad[c] = float(np.median(np.abs(v - np.float32(med[c])))) or 1e-8
        return 0

    monkeypatch.setattr(_native, "load_library", lambda: SimpleNamespace(dewi_fit_stats=fake_fit_stats))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: SimpleNamespace(cuda_stream=0))
    rng = np.random.RandomState(3)
    for f, n in ((7, 101), (32, 50), (33, 64), (70, 33)):
        calls.clear()
        host = (rng.standard_normal((f, n)) * np.arange(1, f + 1)[:, None]).astype(np.float32)
        host[min(3, f - 1)] = 2.5                                  # a constant column: MAD 0 -> 1e-8 (scorer.py:24)
        wide = torch.from_numpy(np.ascontiguousarray(np.pad(host, ((0, 0), (0, 5)))))[:, :n]   # row pitch n + 5
        med, mad = scorer._fit_columns(wide, 0)
        assert calls == [32] * (f // 32) + ([f % 32] if f % 32 else [])
        assert np.array_equal(med, np.median(host, axis=1).astype(np.float64))
        want_mad = np.median(np.abs(host - np.median(host, axis=1, keepdims=True)), axis=1).astype(np.float64)
        want_mad[want_mad == 0] = 1e-8
        assert np.array_equal(mad, want_mad)
        med0, mad0 = scorer._fit_columns(wide, 0, zero_mad_as=0.0)  # robust.py:8-10 adds its own epsilon instead
        assert mad0[min(3, f - 1)] == 0.0 and np.array_equal(med0, med)


def test_column_payload_lookup_by_explicit_ids_is_indexed():
    """Explicit ids over a column store: lookups go through one dict (not `list.index` per call), a duplicated id
    resolves to its last row -- the payload the reference's `_payloads[doc_id] = payload` keeps (backends.py:400)."""
    store = ColumnStore()
    cols = np.arange(6 * len(PAYLOAD_FIELDS), dtype=np.float32).reshape(6, -1)
    store.append(6, cols)
    ids = ["a", "b", "c", "b", "d", "e"]
    view = ColumnPayloads(ids, store)
    assert view["a"].dewi == cols[0, 0] and view["e"].dewi == cols[5, 0]
    assert view["b"].dewi == cols[3, 0]                      # the later "b"
    assert view.get("zz") is None and "zz" not in view and "c" in view and len(view) == 6
    with pytest.raises(KeyError):
        view["zz"]
    n = 200_000                                               # 2e5 lookups: seconds with a dict, hours with list.index
    big = ColumnStore()
    big.append(n, None)
    bv = ColumnPayloads([f"id{i}" for i in range(n)], big)
    assert all(bv.get(f"id{i}") is not None for i in range(0, n, 1))
