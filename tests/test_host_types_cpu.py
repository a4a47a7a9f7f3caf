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
