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
