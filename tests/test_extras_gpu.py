"""Rows N1-N4 of SURVEY.md section 8f: the reference's saved-index directory, bulk score -> payload, the
second RobustStats / local weights, and clustering of the join's pairs."""

import numpy as np
import pytest

import dewi_b200
from oracle import scorer as oscorer
from oracle import search as osearch

from _util import GOLD, PAYLOAD_FIELDS, SIGNAL_FIELDS, check_topk, entropy_column, make_corpus

pytestmark = pytest.mark.gpu


def test_directory_saved_by_the_reference_loads_and_searches():
    """N1: `DewiIndex.load` on a directory written by the reference's DewiIndex.save (ExactIndex backend)."""
    g = np.load(GOLD / "reference_saved_index_queries.npz")
    ix = dewi_b200.DewiIndex.load(GOLD / "reference_saved_index")
    assert len(ix) == 40 and ix.rerank_eta == 0.3 and ix.entropy_pref == 0.5
    assert ix.get_metadata("doc_007") == {"source": "file_7.txt"}
    for q, ids, scores in zip(g["queries"], g["ids"], g["scores"]):
        res = ix.search(q, k=5)  # eta / entropy_pref default to the saved config (index.py:86-89)
        check_topk(np.array([int(i[4:]) for i in ids]), scores, [int(r[0][4:]) for r in res], [r[1] for r in res])


def test_save_load_round_trip_between_dtypes(tmp_path):
    emb, pay = make_corpus(300, 64, seed=12)
    ix = dewi_b200.DewiIndex(dim=64)
    ix.add_batch([f"d{i}" for i in range(300)], emb, payloads=[dewi_b200.Payload(dewi=float(pay[i, 0])) for i in range(300)],
                 normalized=True)
    ix.build()
    ix.save(tmp_path / "ix")
    back = dewi_b200.DewiIndex.load(tmp_path / "ix")
    q = np.random.RandomState(3).standard_normal(64).astype(np.float32)
    assert [r[0] for r in back.search(q, k=7)] == [r[0] for r in ix.search(q, k=7)]
    np.testing.assert_array_equal(back.get_embedding("d17"), emb[17])


def test_bulk_score_to_payload_matches_the_readme_loop():
    """N2: fit_stats -> score -> payload.dewi for every document, as one device-side call."""
    n, d = 5000, 64
    emb, pay = make_corpus(n, d, seed=33, style="readme")
    sig = pay[:, 1:].T.astype(np.float32)
    ix = dewi_b200.CudaIndex(d)
    ix.add_batch(None, emb, normalized=True)
    scorer = dewi_b200.DewiScorer()
    dewi = ix.set_payload_from_signals(sig, scorer).cpu().numpy()
    ix.build()
    cols = {k: sig[i] for i, k in enumerate(SIGNAL_FIELDS)}
    med, mad = oscorer.robust_fit(cols)
    ref_dewi = oscorer.score_rows(cols, med, mad)
    np.testing.assert_allclose(dewi, ref_dewi, rtol=1e-6)
    q = np.random.RandomState(4).standard_normal((6, d)).astype(np.float32)
    ids, sc = ix.search_batch(q, k=10, eta=0.5, entropy_pref=0.25)
    rid, rsc = osearch.exact_search_batch(emb, dewi.astype(np.float64), entropy_column(pay), q, 10, 0.5, 0.25, True)
    for i in range(6):
        check_topk(rid[i], rsc[i], ids[i], sc[i], what=f"q{i}")


def test_payload_robust_stats_and_local_weights_match_the_reference():
    """N3: fixtures produced by the reference's robust.RobustStats and local_weights_from_surprisal."""
    g = np.load(GOLD / "extras_robust_localweights.npz")
    pay = g["payload"]
    payloads = [dewi_b200.Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}) for i in range(len(pay))]
    st = dewi_b200.PayloadRobustStats.from_payloads(payloads)
    keys = dewi_b200.PayloadRobustStats.KEYS
    assert [st.fields[k][0] for k in keys] == g["med"].tolist()
    assert [st.fields[k][1] for k in keys] == g["mad"].tolist()
    assert [st.z(k, float(v)) for k, v in zip(keys, g["probe"])] == g["z"].tolist()
    with pytest.raises(ValueError):
        dewi_b200.PayloadRobustStats.from_payloads([])
    w = dewi_b200.local_weights_from_surprisal(g["surprisal"])
    assert w.dtype == np.float32 and w.shape == g["surprisal"].shape and np.all(w > 0)
    np.testing.assert_allclose(w, g["weights"], rtol=2e-6)
    np.testing.assert_allclose(dewi_b200.local_weights_from_surprisal(g["const"]), g["const_weights"], rtol=2e-6)


def test_clusters_of_join_pairs():
    """N4: connected components of the near-duplicate pairs vs a host union-find; metric helpers."""
    rng = np.random.RandomState(7)
    n, d = 4000, 128
    x = rng.standard_normal((n, d)).astype(np.float32)
    groups = [rng.choice(n, size, replace=False) for size in (2, 3, 5, 9, 2, 2)]
    for gidx in groups:  # chains: each member is a noisy copy of the previous one
        for a, b in zip(gidx[:-1], gidx[1:]):
            x[b] = x[a] + 0.03 * rng.standard_normal(d).astype(np.float32)
    out = dewi_b200.redundancy_join(x, tau=0.95)
    labels = dewi_b200.cluster_pairs(out["pairs_i"], out["pairs_j"], n).cpu().numpy()
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for i, j in zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()):
        ra, rb = find(i), find(j)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    want = np.array([find(i) for i in range(n)])
    _, inv_w = np.unique(want, return_inverse=True)
    _, inv_g = np.unique(labels, return_inverse=True)
    assert np.array_equal(inv_w, inv_g)
    assert np.all(labels <= np.arange(n)) and np.all(labels[labels] == labels)  # label = smallest member, a fixed point
    clusters = dewi_b200.clusters_from_labels(labels)
    assert sum(len(c) for c in clusters) == n
    multi = [c for c in clusters if len(c) > 1]
    assert len(multi) >= 4
    assert dewi_b200.duplicate_rate(clusters) == pytest.approx(len(multi) / len(clusters))
    assert dewi_b200.cluster_coverage([c[0] for c in clusters[: len(clusters) // 2]], clusters) == pytest.approx(
        (len(clusters) // 2) / len(clusters))
    assert dewi_b200.cluster_pairs([], [], 5).cpu().tolist() == [0, 1, 2, 3, 4]


def test_signals_scored_in_bulk_survive_rebuild_search_and_save(tmp_path):
    """add() -> set_payload_from_signals -> build -> search -> save / load: the scores written on the device are
    mirrored into the shared Payload objects, so a rebuild re-uploads the SAME values, search() tuples carry
    them and a reloaded index ranks identically."""
    n, d = 600, 32
    emb, pay = make_corpus(n, d, seed=44, style="readme")
    sig = pay[:, 1:].T.astype(np.float32)
    ix = dewi_b200.DewiIndex(dim=d)
    objs = [dewi_b200.Payload() for _ in range(n)]
    for i in range(n):
        ix.add(f"doc_{i:08d}", emb[i], objs[i])
    scorer = dewi_b200.DewiScorer()
    dewi = ix._backend.set_payload_from_signals(sig, scorer).cpu().numpy()
    assert objs[5].dewi == pytest.approx(float(dewi[5])) and objs[5].ht_mean == float(sig[0, 5]) and objs[5].noise == float(sig[6, 5])
    q = np.random.RandomState(6).standard_normal(d).astype(np.float32)
    rid, rsc = osearch.exact_search(emb, dewi.astype(np.float64), entropy_column(pay), q, 10, 0.5, 0.25, True)
    res = ix.search(q, k=10, eta=0.5, entropy_pref=0.25)                 # lazy build() re-reads the host payloads
    check_topk(rid, rsc, [int(r[0][4:]) for r in res], [r[1] for r in res], what="after build")
    assert res[0][2] is objs[int(res[0][0][4:])] and res[0][2].dewi == pytest.approx(float(dewi[int(res[0][0][4:])]))
    ix.add("late", emb[0] * 0.5 + emb[1], dewi_b200.Payload(dewi=0.0))   # any add() forces another rebuild
    res = ix.search(q, k=10, eta=0.5, entropy_pref=0.25)
    assert [r[0] for r in res if r[0] != "late"][:9] == [f"doc_{i:08d}" for i in rid][:9]
    ix.save(tmp_path / "ix")
    back = dewi_b200.DewiIndex.load(tmp_path / "ix")
    res2 = back.search(q, k=10, eta=0.5, entropy_pref=0.25)
    assert [(r[0], r[1]) for r in res2] == [(r[0], r[1]) for r in res]
    assert back.get_payload("doc_00000005").dewi == pytest.approx(float(dewi[5]))
    # bulk-column ingest: the column store is the mirror
    bx = dewi_b200.CudaIndex(d)
    bx.add_batch(None, emb, payload_columns=np.zeros((n, 8), np.float32), normalized=True)
    bx.set_payload_from_signals(sig, dewi_b200.DewiScorer())
    bx.refresh_payloads()                                                # must not clobber the device columns
    ids, sc = bx.search_batch(q[None], k=10, eta=0.5, entropy_pref=0.25)
    check_topk(rid, rsc, ids[0], sc[0], what="column store mirror")
    assert bx._payloads["doc_00000005"].dewi == pytest.approx(float(dewi[5]))
    # no mirror: the device columns are authoritative, refresh is a no-op, save refuses to persist stale payloads
    cx = dewi_b200.CudaIndex(d)
    cx.add_batch(None, emb, normalized=True)
    cx.set_payload_from_signals(sig, dewi_b200.DewiScorer(), mirror=False)
    cx.build()
    ids, sc = cx.search_batch(q[None], k=10, eta=0.5, entropy_pref=0.25)
    check_topk(rid, rsc, ids[0], sc[0], what="device-authoritative columns")
    with pytest.raises(ValueError):
        cx.save(tmp_path / "stale")


def test_add_batch_rejects_bad_arguments_without_touching_the_index():
    emb, pay = make_corpus(50, 16, seed=2)
    ix = dewi_b200.CudaIndex(16)
    ix.add_batch([f"a{i}" for i in range(50)], emb, payload_columns=pay.astype(np.float32), normalized=True)
    for kwargs in ({"doc_ids": ["x"] * 3}, {"doc_ids": [f"b{i}" for i in range(50)], "payload_columns": np.zeros((50, 3), np.float32)},
                   {"doc_ids": [f"b{i}" for i in range(50)], "payloads": [dewi_b200.Payload()] * 50}):
        with pytest.raises(ValueError):
            ix.add_batch(kwargs.pop("doc_ids"), emb, normalized=True, **kwargs)
        assert len(ix) == 50 and len(ix._doc_ids) == 50 and ix._columns.n == 50
    ix.build()
    q = np.random.RandomState(1).standard_normal(16).astype(np.float32)
    res = ix.search(q, k=5, eta=0.3, entropy_pref=0.5)
    rid, rsc = osearch.exact_search(emb, pay[:, 0], entropy_column(pay), q, 5, 0.3, 0.5, True)
    check_topk(rid, rsc, [int(r[0][1:]) for r in res], [r[1] for r in res])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_bulk_export_and_sidecar_round_trip(tmp_path, dtype):
    """N1 at scale: rows leave the device in bulk (never per row); a bf16 index writes a sharded sidecar that
    reloads bit-exactly; the fp32 file is what ExactIndex.load reads."""
    from _util import bf16_round

    n, d = 3000, 64
    emb, pay = make_corpus(n, d, seed=13)
    ix = dewi_b200.CudaIndex(d, dtype=dtype)
    ix.SIDECAR_SHARD_ROWS = 1024                       # several shards at this size
    ix.FP32_EXPORT_CHUNK = 700
    ix.add_batch([f"d{i}" for i in range(n)], emb, payload_columns=pay.astype(np.float32), normalized=True)
    ix.build()
    want = emb if dtype == "fp32" else bf16_round(emb)
    np.testing.assert_array_equal(ix.export_rows(100, 1500), want[100:1600])
    np.testing.assert_array_equal(ix._embeddings, want)
    dewi_c, ent_c = ix.get_payload_columns(10, 20)
    np.testing.assert_array_equal(dewi_c, pay[10:30, 0].astype(np.float32))
    np.testing.assert_array_equal(ent_c, entropy_column(pay)[10:30].astype(np.float32))
    ix._host_rows = None
    ix.save(tmp_path / "ix")
    meta = __import__("json").loads((tmp_path / "ix" / "metadata.json").read_text())
    np.testing.assert_array_equal(np.load(tmp_path / "ix" / "embeddings.npy"), want)
    if dtype == "bf16":
        assert [s["rows"] for s in meta["bf16_sidecar"]] == [1024, 1024, 952]
        raw = np.concatenate([np.load(tmp_path / "ix" / s["file"]) for s in meta["bf16_sidecar"]])
        assert raw.dtype == np.uint16
        np.testing.assert_array_equal((raw.astype(np.uint32) << 16).view(np.float32), want)
    else:
        assert "bf16_sidecar" not in meta
    back = dewi_b200.CudaIndex.load(tmp_path / "ix")
    np.testing.assert_array_equal(back.export_rows(), want)
    q = np.random.RandomState(3).standard_normal((4, d)).astype(np.float32)
    a = ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
    b = back.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    if dtype == "bf16":                                # sidecar only (no fp32 file): still loads here
        ix.save(tmp_path / "side", fp32=False)
        assert not (tmp_path / "side" / "embeddings.npy").exists()
        np.testing.assert_array_equal(dewi_b200.CudaIndex.load(tmp_path / "side").export_rows(), want)
