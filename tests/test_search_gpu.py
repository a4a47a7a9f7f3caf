"""Parity of the CUDA search path with the reference's exact numpy re-rank path (through the C ABI).

fp32 mode: identical top-k ids outside ties, scores within 1e-5 relative (BASELINE.json north star).
bf16-storage mode: recall@10 >= 0.999 against the exact result on the same bf16-representable rows."""

import numpy as np
import pytest

import dewi_b200
from dewi_b200 import _native
from oracle import search as osearch

from _util import (PAYLOAD_FIELDS, bf16_round, check_topk, entropy_column, load_search_golden, make_corpus,
                   recall_at_k)

pytestmark = pytest.mark.gpu

GOLDEN = ["cos_n100_d128", "cos_n2000_d64", "cos_n3000_d768", "cos_n15_d16_k5", "cos_n12_d8_k10", "l2_n500_d32"]


def payload_objects(pay):
    return [dewi_b200.Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}) for i in range(len(pay))]


def build_index(emb, pay, space="cosine", **kw):
    """Per-document add() exactly as a reference user would (tests/test_index.py:103-127)."""
    ix = dewi_b200.DewiIndex(dim=emb.shape[1], space=space, backend="cuda", **kw)
    for i, p in enumerate(payload_objects(pay)):
        ix.add(f"doc_{i:08d}", emb[i], p)
    ix.build()
    return ix


def bulk_index(emb, pay, dtype="fp32", normalized=True, **kw):
    ix = dewi_b200.DewiIndex(dim=emb.shape[1], space="cosine", backend="cuda", dtype=dtype, **kw)
    ix.add_batch(None, emb, payload_columns=pay.astype(np.float32), normalized=normalized)
    ix.build()
    return ix


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_single_query_api(name):
    """`DewiIndex.search` (1-D query contract) against outputs of the unmodified reference."""
    g = load_search_golden(name)
    ix = build_index(g["emb"], g["payload"], g["space"])
    for gi, (eta, pref) in enumerate(g["grid"]):
        for qi, q in enumerate(g["queries"]):
            res = ix.search(q, k=g["k"], eta=eta, entropy_pref=pref)
            assert all(isinstance(r[0], str) and isinstance(r[1], float) and isinstance(r[2], dewi_b200.Payload) for r in res)
            got_idx = [int(r[0][4:]) for r in res]
            check_topk(g["ref_idx"][gi, qi], g["ref_scores"][gi, qi], got_idx, [r[1] for r in res],
                       what=f"{name} grid{gi} q{qi}")


@pytest.mark.parametrize("name", ["cos_n2000_d64", "cos_n3000_d768"])
@pytest.mark.parametrize("path", ["simt", "tc"])
def test_golden_batch_both_sweeps(name, path):
    """The tcgen05 sweep and the CUDA-core sweep are interchangeable: both reproduce the reference."""
    g = load_search_golden(name)
    ix = build_index(g["emb"], g["payload"], g["space"])
    flag = _native.FLAG_FORCE_TC if path == "tc" else _native.FLAG_FORCE_SIMT
    for gi, (eta, pref) in enumerate(g["grid"]):
        ids, sc = ix._backend.search_batch(g["queries"], g["k"], eta, pref, flags=flag)
        for qi in range(len(g["queries"])):
            check_topk(g["ref_idx"][gi, qi], g["ref_scores"][gi, qi], ids[qi], sc[qi], what=f"{name}/{path} grid{gi} q{qi}")


@pytest.mark.parametrize("n,d,b,k", [(50_000, 768, 33, 10), (20_011, 128, 130, 10), (4_096, 512, 5, 64), (300_000, 64, 7, 10)])
def test_fp32_differential_vs_oracle(n, d, b, k):
    """Seeded corpora incl. ragged tile tails and B that is not a multiple of the query block."""
    emb, pay = make_corpus(n, d, seed=100 + d)
    rng = np.random.RandomState(5)
    queries = rng.standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay)
    ent = entropy_column(pay)
    for eta, pref in [(0.3, 0.5), (0.0, 0.0), (1.0, -1.0)]:
        ids, sc = ix.search_batch(queries, k=k, eta=eta, entropy_pref=pref)
        rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], ent, queries, k, eta, pref, True)
        for qi in range(b):
            check_topk(rid[qi], rsc[qi], ids[qi], sc[qi], what=f"n{n} d{d} eta{eta} pref{pref} q{qi}")


def test_device_tensor_io_matches_host_io():
    import torch

    emb, pay = make_corpus(30_000, 256, seed=3)
    ix = bulk_index(emb, pay)
    q = np.random.RandomState(1).standard_normal((16, 256)).astype(np.float32)
    ids_h, sc_h = ix.search_batch(q, k=10, eta=0.25, entropy_pref=0.0)
    ids_d, sc_d = ix.search_batch(torch.from_numpy(q).cuda(), k=10, eta=0.25, entropy_pref=0.0)
    assert np.array_equal(ids_h, ids_d.cpu().numpy()) and np.array_equal(sc_h, sc_d.cpu().numpy())
    assert ix._backend.last_launches() >= 3   # query prep, sweep, fused tail (+ pre-pass and seed on larger corpora)


def test_device_side_normalisation_close_to_numpy():
    """add_batch(normalized=False) normalises on the device; rows agree with numpy to 1 ulp-ish."""
    rng = np.random.RandomState(8)
    raw = (rng.standard_normal((5000, 96)) * rng.uniform(0.1, 30, (5000, 1))).astype(np.float32)
    pay = np.zeros((5000, 8))
    ix = bulk_index(raw, pay, normalized=False)
    ref = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    got = np.stack([ix._backend.get_row(i) for i in (0, 17, 4999)])
    np.testing.assert_allclose(got, ref[[0, 17, 4999]], rtol=3e-7, atol=1e-9)
    with pytest.raises(ValueError):
        ix.add_batch(None, np.zeros((2, 96), np.float32), normalized=False)  # zero-norm row


def test_bf16_storage_recall_gate():
    """bf16 corpus: recall@10 >= 0.999 vs the exact search over the same bf16-representable rows
    (bulk-assigned into the oracle); queries stay fp32 on both sides."""
    n, d, b, k = 200_000, 768, 1024, 10
    emb, pay = make_corpus(n, d, seed=77)
    emb16 = bf16_round(emb)
    queries = np.random.RandomState(78).standard_normal((b, d)).astype(np.float32)
    ent = entropy_column(pay)
    rid, rsc = osearch.exact_search_batch(emb16, pay[:, 0], ent, queries, k, 0.3, 0.5, True)
    for precise in (False, True):
        ix = bulk_index(emb, pay, dtype="bf16", precise_query=precise)
        np.testing.assert_array_equal(ix._backend.get_row(123), emb16[123])
        ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
        rec = recall_at_k(rid, ids)
        assert rec >= 0.999, f"recall@{k} = {rec} (precise_query={precise})"
        same = ids == rid
        np.testing.assert_allclose(sc[same], rsc[same], rtol=1e-5)


def test_edge_semantics():
    g = load_search_golden("cos_n15_d16_k5")
    ix = build_index(g["emb"], g["payload"])
    with pytest.raises(ValueError):  # k > N (backends.py:468)
        ix.search(g["queries"][0], k=16)
    with pytest.raises(ValueError):  # 2-D query (index.py:91-92)
        ix.search(g["queries"][:2], k=3)
    with pytest.raises(ValueError):  # wrong embedding shape (backends.py:395-396)
        ix.add("x", np.zeros(3, np.float32), dewi_b200.Payload())
    assert len(ix.search(g["queries"][0], k=15)) == 15  # k == N: every document, sorted
    # zero query: all similarities are 0, ranking by eta*dewi + pref*ent over the candidate set
    res = ix.search(np.zeros(16, np.float32), k=15, eta=0.5, entropy_pref=0.25)
    pay = g["payload"]
    adj = np.float32(0.5) * pay[:, 0].astype(np.float32) + np.float32(0.25) * entropy_column(pay).astype(np.float32)
    assert [int(r[0][4:]) for r in res] == np.argsort(-adj, kind="stable").tolist()
    empty = dewi_b200.DewiIndex(dim=4)
    with pytest.raises(ValueError):  # backends.py:409-410
        empty.build()


def test_payload_mutation_and_refresh():
    """README flow: payload.dewi is assigned after add(); build()/refresh_payloads() snapshot it."""
    emb, pay = make_corpus(500, 32, seed=9)
    ix = dewi_b200.DewiIndex(dim=32)
    objs = payload_objects(pay)
    for i, p in enumerate(objs):
        p.dewi = 0.0
        ix.add(f"doc_{i:08d}", emb[i], p)
    for i, p in enumerate(objs):
        p.dewi = float(pay[i, 0])  # README.md:106-110
    q = np.random.RandomState(2).standard_normal(32).astype(np.float32)
    res = ix.search(q, k=10, eta=0.5, entropy_pref=0.0)  # lazy build reads the updated payloads
    rid, rsc = osearch.exact_search(emb, pay[:, 0], entropy_column(pay), q, 10, 0.5, 0.0, True)
    check_topk(rid, rsc, [int(r[0][4:]) for r in res], [r[1] for r in res], what="after mutation")
    assert res[0][2] is objs[int(res[0][0][4:])]  # payloads are returned by reference
    for p in objs:
        p.dewi = 1.0 - p.dewi
    ix.refresh_payloads()
    res2 = ix.search(q, k=10, eta=0.5, entropy_pref=0.0)
    rid2, rsc2 = osearch.exact_search(emb, 1.0 - pay[:, 0], entropy_column(pay), q, 10, 0.5, 0.0, True)
    check_topk(rid2, rsc2, [int(r[0][4:]) for r in res2], [r[1] for r in res2], what="after refresh")


def test_reference_behaviour_tests_rerun_on_cuda_backend(tmp_path):
    """The reference's own test_index.py assertions (shape/order/monotonicity/persistence), re-run
    with the CUDA backend substituted (tests/test_index.py:103-127,206-353)."""
    np.random.seed(42)
    dim, n, k = 128, 100, 10
    emb = np.random.randn(n, dim).astype(np.float32)
    ix = dewi_b200.DewiIndex(dim=dim, space="cosine", use_ann=False)
    for i in range(n):
        ix.add(f"doc_{i}", emb[i], dewi_b200.Payload(dewi=np.random.rand(), ht_mean=np.random.rand() * 3,
                                                     hi_mean=np.random.rand() * 3))
    q = np.random.randn(dim).astype(np.float32)
    res = ix.search(q, k=k)
    assert len(res) == k and len(ix) == n
    assert all(res[i][1] >= res[i + 1][1] for i in range(k - 1))
    assert all(r[0].startswith("doc_") for r in res)
    mean_ent = [np.mean([(r[2].ht_mean + r[2].hi_mean) / 2 for r in ix.search(q, k=k, eta=0.0, entropy_pref=p)])
                for p in (-1.0, 0.0, 1.0)]
    assert mean_ent[0] <= mean_ent[1] <= mean_ent[2]
    mean_dewi = [np.mean([r[2].dewi for r in ix.search(q, k=k, eta=e)]) for e in (0.0, 0.5, 1.0)]
    assert mean_dewi[0] <= mean_dewi[1] <= mean_dewi[2]
    ix.save(tmp_path / "idx")
    for f in ("config.json", "ann_index/metadata.json", "ann_index/payloads.jsonl", "ann_index/embeddings.npy"):
        assert (tmp_path / "idx" / f).exists()
    back = dewi_b200.DewiIndex.load(tmp_path / "idx")
    res2 = back.search(q, k=5)
    assert [r[0] for r in res2] == [r[0] for r in ix.search(q, k=5)]
    assert ix.get_payload("doc_3") is not None and ix.get_payload("nope") is None
    np.testing.assert_allclose(ix.get_embedding("doc_3"), emb[3] / np.linalg.norm(emb[3]), rtol=1e-6)


@pytest.mark.parametrize("n_shards", [2, 5])
def test_virtual_shards_merge_equals_single_index(n_shards):
    """The multi-GPU merge on one device: S shard indexes with global id bases write their stage-1
    blocks into one 'gathered' buffer, `dewi_rerank` reads it shard-strided."""
    import ctypes

    import torch

    from dewi_b200.sharded import PackedCandidates, shard_range

    n, d, b, k = 40_000, 128, 9, 10
    emb, pay = make_corpus(n, d, seed=31)
    queries = torch.from_numpy(np.random.RandomState(32).standard_normal((b, d)).astype(np.float32)).cuda()
    kcand = 2 * k
    pk = PackedCandidates(b, kcand, queries.device, world=n_shards)
    blocks = pk.gathered.view(n_shards, pk.words)
    shards = []
    for g in range(n_shards):
        lo, hi = shard_range(n, n_shards, g, align=64)
        ix = dewi_b200.CudaIndex(d)
        ix.add_batch(None, emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        ix.set_id_base(lo)
        ix.build()
        ids, sim, dewi, ent = pk.views(blocks[g])
        ix.search_local_into(queries, kcand, sim, ids, dewi, ent)
        shards.append(ix)
    ids, sim, dewi, ent = pk.views(pk.gathered)
    out_ids = torch.empty((b, k), dtype=torch.int64, device="cuda")
    out_sc = torch.empty((b, k), dtype=torch.float32, device="cuda")
    lib = _native.load_library()
    _native.check(lib.dewi_rerank(ctypes.c_void_p(sim.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                                  ctypes.c_void_p(dewi.data_ptr()), ctypes.c_void_p(ent.data_ptr()), b, n_shards, kcand,
                                  pk.stride_bytes, kcand, k, 0.3, 0.5, ctypes.c_void_p(out_ids.data_ptr()),
                                  ctypes.c_void_p(out_sc.data_ptr()), 0, _native.stream_ptr()))
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), queries.cpu().numpy(), k, 0.3, 0.5, True)
    for q in range(b):
        check_topk(rid[q], rsc[q], out_ids[q].cpu().numpy(), out_sc[q].cpu().numpy(), what=f"S{n_shards} q{q}")


@pytest.mark.parametrize("dtype,b", [("fp32", 300), ("bf16", 1024), ("bf16", 129)])
def test_cta_pair_sweep_matches_single_cta_sweep_and_oracle(dtype, b):
    """B > 128 runs the cta_group::2 sweep (two query blocks per corpus tile); DEWI_FLAG_NO_PAIR keeps
    the 1-CTA sweep.  Both must agree with each other and with the oracle."""
    n, d, k = 70_001, 256, 10
    emb, pay = make_corpus(n, d, seed=55)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    queries = np.random.RandomState(56).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype=dtype)
    ids2, sc2 = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    ids1, sc1 = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_NO_PAIR)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries[:64], k, 0.3, 0.5, True)
    if dtype == "fp32":
        for q in range(64):
            check_topk(rid[q], rsc[q], ids2[q], sc2[q], what=f"pair q{q}")
            check_topk(rid[q], rsc[q], ids1[q], sc1[q], what=f"single q{q}")
        assert (ids1 == ids2).mean() > 0.999
    else:
        assert recall_at_k(rid, ids2[:64]) >= 0.999 and recall_at_k(rid, ids1[:64]) >= 0.999
        assert recall_at_k(ids1, ids2) >= 0.999


@pytest.mark.parametrize("dtype,b", [("fp32", 512), ("bf16", 640)])
def test_pair_sweep_with_rendezvous_matches_oracle(dtype, b):
    """Several query-block pairs over a corpus long enough (>= 64 tiles per work item) for the clusters that
    stream one chunk to rendezvous at their arrival counters (search_tc2.cu): results must not depend on
    it -- the same answer as the oracle and as the 1-CTA sweep, call after call."""
    n, d, k = 750_000, 64, 10
    emb, pay = make_corpus(n, d, seed=91)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    queries = np.random.RandomState(92).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype=dtype)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    for _ in range(3):
        ids_again, sc_again = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
        assert np.array_equal(ids, ids_again) and np.array_equal(sc, sc_again)
    ids1, sc1 = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_NO_PAIR)
    sel = np.arange(0, b, 7)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries[sel], k, 0.3, 0.5, True)
    if dtype == "fp32":
        for i, q in enumerate(sel):
            check_topk(rid[i], rsc[i], ids[q], sc[q], what=f"rendezvous q{q}")
        assert (ids1 == ids).mean() > 0.999
    else:
        assert recall_at_k(rid, ids[sel]) >= 0.999 and recall_at_k(ids1, ids) >= 0.999


@pytest.mark.parametrize("b", [5, 200])
def test_seeded_sweep_equals_unseeded_and_oracle(b):
    """Large corpora run a sample pre-pass that seeds the admission thresholds (DEWI_FLAG_NO_SEED turns
    it off).  Duplicated rows make exact score ties at the threshold: nothing at the seed may be lost."""
    n, d, k = 700_000, 64, 10
    emb, pay = make_corpus(n, d, seed=91)
    emb[5000:5040] = emb[100]          # 41 identical rows, inside the sample
    emb[600_000:600_020] = emb[100]    # ... and far outside it
    queries = np.random.RandomState(92).standard_normal((b, d)).astype(np.float32)
    queries[0] = emb[100] + 0.01 * queries[0]
    ix = bulk_index(emb, pay)
    ent = entropy_column(pay)
    ids_s, sc_s = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    ids_u, sc_u = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_NO_SEED)
    np.testing.assert_allclose(sc_s, sc_u, rtol=1e-6)
    nq = min(b, 12)
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], ent, queries[:nq], k, 0.3, 0.5, True)
    for q in range(1, nq):
        check_topk(rid[q], rsc[q], ids_s[q], sc_s[q], what=f"seeded q{q}")
    # query 0 sits on a 61-way exact tie at the top: the reference keeps an arbitrary 20 of the tied rows
    # as candidates, so any tied rows are valid -- but they must BE tied rows, with the blended score
    tied = {100, *range(5000, 5040), *range(600_000, 600_020)}
    assert set(ids_s[0].tolist()) <= tied and set(ids_u[0].tolist()) <= tied
    qn = queries[0] / np.linalg.norm(queries[0])
    for i, s in zip(ids_s[0], sc_s[0]):
        want = np.float32(0.7) * np.float32(emb[i] @ qn) + np.float32(0.3) * np.float32(pay[i, 0]) + np.float32(0.5) * np.float32(ent[i])
        assert abs(s - want) <= 1e-5 * max(1.0, abs(want))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_m64_and_m128_sweeps_agree(dtype):
    """B <= 64 runs M = 64 MMAs (accumulator rows on TMEM lanes 0-15 of each quarter); DEWI_FLAG_NO_M64
    keeps M = 128.  Same results either way, and both match the oracle."""
    n, d, b, k = 90_000, 192, 64, 10
    emb, pay = make_corpus(n, d, seed=61)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    queries = np.random.RandomState(62).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype=dtype)
    ids64, sc64 = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    ids128, sc128 = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_NO_M64)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries, k, 0.3, 0.5, True)
    if dtype == "fp32":
        for q in range(b):
            check_topk(rid[q], rsc[q], ids64[q], sc64[q], what=f"M64 q{q}")
            check_topk(rid[q], rsc[q], ids128[q], sc128[q], what=f"M128 q{q}")
    else:
        assert recall_at_k(rid, ids64) >= 0.999 and recall_at_k(rid, ids128) >= 0.999


def test_precise_query_small_batch():
    """bf16 corpus, hi+lo query planes (mode 1), B <= 64 (M = 64 MMAs)."""
    n, d, b, k = 60_000, 256, 40, 10
    emb, pay = make_corpus(n, d, seed=71)
    emb16 = bf16_round(emb)
    queries = np.random.RandomState(72).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype="bf16", precise_query=True)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    rid, rsc = osearch.exact_search_batch(emb16, pay[:, 0], entropy_column(pay), queries, k, 0.3, 0.5, True)
    for q in range(b):  # with both query planes the sweep is as exact as the fp32 one
        check_topk(rid[q], rsc[q], ids[q], sc[q], what=f"q{q}")


@pytest.mark.parametrize("b", [5, 200])
def test_certified_single_plane_sweep_is_exact(b):
    """fp32 corpus swept through its bf16 hi plane only (half the bytes / a third of the MMAs) + certificate + exact
    re-score: same answers as the full hi/lo product and as the oracle; the certificate held for every query."""
    n, d, k = 120_000, 768, 10
    emb, pay = make_corpus(n, d, seed=91)
    ix = bulk_index(emb, pay)
    q = np.random.RandomState(92).standard_normal((b, d)).astype(np.float32)
    used0, failed0 = ix._backend.cert_stats()
    ids, sc = ix.search_batch(q, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_FORCE_TC)
    used1, failed1 = ix._backend.cert_stats()
    assert used1 == used0 + 1 and failed1 == failed0, "the certified sweep did not run, or its certificate failed"
    ids3, sc3 = ix.search_batch(q, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_FORCE_TC | _native.FLAG_NO_CERT)
    assert ix._backend.cert_stats() == (used1, failed1)
    np.testing.assert_array_equal(ids, ids3)
    np.testing.assert_array_equal(sc, sc3)          # both orders come from the same exact fp32 re-score
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), q[:24], k, 0.3, 0.5, True)
    for i in range(min(b, 24)):
        check_topk(rid[i], rsc[i], ids[i], sc[i], what=f"certified B={b} q{i}")


def test_certificate_failure_falls_back_to_the_full_product():
    """Scores packed more densely than the rounding bound resolves: 300 rows within 2e-4 of each other at the top.
    The certificate cannot be given, the batch is re-run with the hi/lo product, and the answer is still exact."""
    n, d, k = 60_000, 768, 10
    rng = np.random.RandomState(93)
    emb, pay = make_corpus(n, d, seed=94)
    target = rng.standard_normal(d).astype(np.float32)
    target /= np.linalg.norm(target)
    for j, row in enumerate(rng.choice(n, 300, replace=False)):
        v = target + (2e-3 + 1e-6 * j) * rng.standard_normal(d).astype(np.float32)
        emb[row] = v / np.linalg.norm(v)
    ix = bulk_index(emb, pay)
    q = np.stack([target, rng.standard_normal(d).astype(np.float32)])
    used0, failed0 = ix._backend.cert_stats()
    ids, sc = ix.search_batch(q, k=k, eta=0.0, entropy_pref=0.0, flags=_native.FLAG_FORCE_TC)
    used1, failed1 = ix._backend.cert_stats()
    assert used1 == used0 + 1 and failed1 == failed0 + 1
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), q, k, 0.0, 0.0, True)
    for i in range(2):
        check_topk(rid[i], rsc[i], ids[i], sc[i], what=f"fallback q{i}")


def _sweep_kind(ix, queries, k, flags=0):
    be = ix._backend
    be.set_profiling(True)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=flags)
    kind = be.sweep_ms(0)[1]
    be.set_profiling(False)
    return ids, sc, kind


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("b", [1, 7, 16, 17, 33, 64])
def test_rows_on_m_sweep_matches_queries_on_m_and_oracle(dtype, b):
    """B <= 64 puts the corpus rows on the MMA M dimension and the queries on N = 16 / 32 / 64 (search_tcr.cu);
    DEWI_FLAG_NO_ROWS_ON_M keeps the queries on M.  Ragged corpus tail (n % 256 != 0), with and without the seeding
    pre-pass: same results from both kernels, and both match the oracle."""
    n, d, k = 300_077, 128, 10
    emb, pay = make_corpus(n, d, seed=161 + b)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    queries = np.random.RandomState(162).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype=dtype)
    force = _native.FLAG_FORCE_TC | (_native.FLAG_FORCE_CERT if dtype == "fp32" else 0)   # fp32: certified fp16-plane sweep
    ids_r, sc_r, kind_r = _sweep_kind(ix, queries, k, force)
    ids_q, sc_q, kind_q = _sweep_kind(ix, queries, k, force | _native.FLAG_NO_ROWS_ON_M)
    ids_u, sc_u, kind_u = _sweep_kind(ix, queries, k, force | _native.FLAG_NO_SEED)
    if dtype == "bf16":   # (an fp32 batch whose certificate fails is re-run with both planes, queries on M)
        assert kind_r == "tcgen05-rows" and kind_u == "tcgen05-rows"
    assert kind_q == "tcgen05"
    np.testing.assert_array_equal(ids_r, ids_q)
    np.testing.assert_array_equal(sc_r, sc_q)
    np.testing.assert_array_equal(ids_r, ids_u)
    np.testing.assert_array_equal(sc_r, sc_u)
    nq = min(b, 16)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries[:nq], k, 0.3, 0.5, True)
    if dtype == "fp32":
        for q in range(nq):
            check_topk(rid[q], rsc[q], ids_r[q], sc_r[q], what=f"rows-on-M B={b} q{q}")
    else:
        assert recall_at_k(rid, ids_r[:nq]) >= 0.999


@pytest.mark.parametrize("b", [3, 40])
def test_rows_on_m_certified_sweep_is_exact(b):
    """fp32 corpus, certified single-plane sweep (48-slot lists) on the rows-on-M kernel."""
    n, d, k = 150_000, 768, 10
    emb, pay = make_corpus(n, d, seed=171)
    ix = bulk_index(emb, pay)
    queries = np.random.RandomState(172).standard_normal((b, d)).astype(np.float32)
    flags = _native.FLAG_FORCE_TC | _native.FLAG_FORCE_CERT
    used0, failed0 = ix._backend.cert_stats()
    ids, sc, kind = _sweep_kind(ix, queries, k, flags)
    used1, failed1 = ix._backend.cert_stats()
    assert kind == "tcgen05-rows" and used1 == used0 + 1 and failed1 == failed0
    ids_q, sc_q, kind_q = _sweep_kind(ix, queries, k, flags | _native.FLAG_NO_ROWS_ON_M)
    assert kind_q == "tcgen05"
    np.testing.assert_array_equal(ids, ids_q)
    np.testing.assert_array_equal(sc, sc_q)
    nq = min(b, 12)
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), queries[:nq], k, 0.3, 0.5, True)
    for q in range(nq):
        check_topk(rid[q], rsc[q], ids[q], sc[q], what=f"certified rows-on-M B={b} q{q}")


@pytest.mark.parametrize("b", [2, 64])
def test_rows_on_m_sweep_with_lists_overflowing_all_the_time(b):
    """Worst case for the shared candidate lists: the corpus is sorted by ASCENDING similarity to query 0, so every
    row beats the running threshold, every list overflows in every half-tile and the prune / retry rounds of the four
    epilogue warps run continuously (no seed: DEWI_FLAG_NO_SEED).  Results stay exact."""
    n, d, k = 60_000, 64, 10
    emb, pay = make_corpus(n, d, seed=181)
    emb16 = bf16_round(emb)
    queries = np.random.RandomState(182).standard_normal((b, d)).astype(np.float32)
    qn = queries[0] / np.linalg.norm(queries[0])
    order = np.argsort(emb16 @ qn, kind="stable")
    emb, pay = np.ascontiguousarray(emb[order]), np.ascontiguousarray(pay[order])
    ix = bulk_index(emb, pay, dtype="bf16")
    ids, sc, kind = _sweep_kind(ix, queries, k, _native.FLAG_FORCE_TC | _native.FLAG_NO_SEED)
    assert kind == "tcgen05-rows"
    ids_q, sc_q, _ = _sweep_kind(ix, queries, k, _native.FLAG_FORCE_TC | _native.FLAG_NO_SEED | _native.FLAG_NO_ROWS_ON_M)
    np.testing.assert_array_equal(ids, ids_q)
    np.testing.assert_array_equal(sc, sc_q)
    rid, rsc = osearch.exact_search_batch(bf16_round(emb), pay[:, 0], entropy_column(pay), queries[:8], k, 0.3, 0.5, True)
    assert recall_at_k(rid, ids[:8]) >= 0.999


# ---- rerank_scope="full": the blend applied over the whole corpus (opt-in, not the reference's semantics) ----------
def _full_oracle(rows, pay, queries, k, eta, pref, normalize=True):
    ent = entropy_column(pay)
    out = [osearch.full_scope_search(rows, pay[:, 0], ent, q, k, eta, pref, normalize) for q in queries]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


@pytest.mark.parametrize("b", [3, 70])
def test_full_scope_fp32_matches_the_blend_over_every_row(b):
    """fp32 corpus -> exact CUDA-core sweep selecting by the blended key; compared with the reference's blend
    statements applied to every row (oracle.search.full_scope_search).  The answer differs from the two-stage one."""
    n, d, k = 40_000, 96, 10
    emb, pay = make_corpus(n, d, seed=201, style="readme")
    queries = np.random.RandomState(202).standard_normal((b, d)).astype(np.float32)
    ix = dewi_b200.DewiIndex(dim=d, backend="cuda", rerank_scope="full")
    ix.add_batch(None, emb, payload_columns=pay.astype(np.float32), normalized=True)
    ix.build()
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    rid, rsc = _full_oracle(emb, pay, queries[:16], k, 0.3, 0.5)
    for q in range(min(b, 16)):
        check_topk(rid[q], rsc[q], ids[q], sc[q], what=f"full scope q{q}")
    two_stage, _ = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), queries[:4], k, 0.3, 0.5, True)
    assert (two_stage != rid[:4]).any(), "the two scopes should not coincide on this corpus"
    # single-query facade and eta = 1 / pref = 0 (pure dewi ranking: the same ids for every query)
    res = ix.search(queries[0], k=k, eta=0.3, entropy_pref=0.5)
    assert [int(r[0].split("_")[1]) for r in res] == ids[0].tolist()
    ids1, _ = ix.search_batch(queries[:2], k=k, eta=1.0, entropy_pref=0.0)
    dewi32 = pay[:, 0].astype(np.float32)
    np.testing.assert_array_equal(np.sort(dewi32[ids1[0]]), np.sort(dewi32)[-k:])
    np.testing.assert_array_equal(ids1[0], ids1[1])


def test_full_scope_l2_space_and_odd_dim():
    n, d, k, b = 9_000, 50, 7, 5
    rng = np.random.RandomState(211)
    emb = (0.25 * rng.standard_normal((n, d))).astype(np.float32)
    _, pay = make_corpus(n, 8, seed=212, style="readme")
    queries = (0.25 * rng.standard_normal((b, d))).astype(np.float32)
    ix = dewi_b200.DewiIndex(dim=d, space="l2", backend="cuda", rerank_scope="full")
    ix.add_batch(None, emb, payload_columns=pay.astype(np.float32))
    ix.build()
    ids, sc = ix.search_batch(queries, k=k, eta=0.25, entropy_pref=-0.5)
    rid, rsc = _full_oracle(emb, pay, queries, k, 0.25, -0.5, normalize=False)
    for q in range(b):
        check_topk(rid[q], rsc[q], ids[q], sc[q], what=f"full scope l2 q{q}")


@pytest.mark.parametrize("b", [1, 64, 150])
def test_full_scope_bf16_runs_on_the_tensor_cores(b):
    """bf16 corpus -> rows-on-M sweep with the blended key evaluated per row in its epilogue (batches above 64 in
    groups of 64); recall against the blend over every (bf16-representable) row."""
    n, d, k = 400_123, 128, 10
    emb, pay = make_corpus(n, d, seed=221, style="profile")
    rows = bf16_round(emb)
    queries = np.random.RandomState(222).standard_normal((b, d)).astype(np.float32)
    ix = dewi_b200.DewiIndex(dim=d, backend="cuda", dtype="bf16", rerank_scope="full")
    ix.add_batch(None, emb, payload_columns=pay.astype(np.float32), normalized=True)
    ix.build()
    be = ix._backend
    be.set_profiling(True)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    assert be.sweep_ms(0)[1] == "tcgen05-rows"
    be.set_profiling(False)
    nq = min(b, 24)
    sel = np.linspace(0, b - 1, nq).astype(int)
    rid, rsc = _full_oracle(rows, pay, queries[sel], k, 0.3, 0.5)
    assert recall_at_k(rid, ids[sel]) >= 0.999
    np.testing.assert_allclose(sc[sel], rsc, rtol=2e-5, atol=2e-6)
    # similarity only (eta = 0, no entropy term): the full scope then IS the top-k by similarity
    ids0, sc0 = ix.search_batch(queries[:4], k=k, eta=0.0, entropy_pref=0.0)
    be._flags &= ~_native.FLAG_SCOPE_FULL   # the same index in the reference's candidate scope
    ids_c, sc_c = ix.search_batch(queries[:4], k=k, eta=0.0, entropy_pref=0.0)
    np.testing.assert_array_equal(ids0, ids_c)
    np.testing.assert_array_equal(sc0, sc_c)


@pytest.mark.parametrize("n_shards", [3])
def test_full_scope_virtual_shards_merge_equals_single_index(n_shards):
    """The global top-k by the blend lies in the union of the shards' local ones: per-shard search_local (blended
    key) + one dewi_rerank over the concatenated blocks == the single-index full-scope answer."""
    import torch
    n, d, k, b = 90_000, 64, 10, 9
    emb, pay = make_corpus(n, d, seed=231, style="readme")
    queries = np.random.RandomState(232).standard_normal((b, d)).astype(np.float32)
    single = dewi_b200.DewiIndex(dim=d, backend="cuda", dtype="bf16", rerank_scope="full")
    single.add_batch(None, emb, payload_columns=pay.astype(np.float32), normalized=True)
    single.build()
    ids1, sc1 = single.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    bounds = np.linspace(0, n, n_shards + 1).astype(int)
    q_dev = torch.from_numpy(queries).cuda()
    kcand = 2 * k
    blocks = []
    for g in range(n_shards):
        lo, hi = bounds[g], bounds[g + 1]
        sh = dewi_b200.CudaIndex(d, dtype="bf16", rerank_scope="full")
        sh.set_id_base(int(lo))
        sh.add_batch(None, emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        sh.build()
        sh.set_blend(0.3, 0.5)
        blocks.append(sh.search_local(q_dev, kcand))
    sim, gid, dewi, ent = (torch.cat([blk[i] for blk in blocks], dim=1).contiguous() for i in range(4))
    lib = _native.load_library()
    out_ids = torch.empty((b, k), dtype=torch.int64, device="cuda")
    out_sc = torch.empty((b, k), dtype=torch.float32, device="cuda")
    import ctypes
    _native.check(lib.dewi_rerank(ctypes.c_void_p(sim.data_ptr()), ctypes.c_void_p(gid.data_ptr()), ctypes.c_void_p(dewi.data_ptr()),
                                  ctypes.c_void_p(ent.data_ptr()), b, 1, kcand * n_shards, 0, kcand * n_shards, k, 0.3, 0.5,
                                  ctypes.c_void_p(out_ids.data_ptr()), ctypes.c_void_p(out_sc.data_ptr()), 0, _native.stream_ptr()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out_ids.cpu().numpy(), ids1)
    np.testing.assert_array_equal(out_sc.cpu().numpy(), sc1)


@pytest.mark.parametrize("dtype,b", [("bf16", 1024), ("fp32", 1500)])
def test_staged_pair_sweep_equals_unseeded_and_oracle(dtype, b):
    """Many query pairs over a small corpus: the CTA-pair sweep runs its first round of chunks as a launch of its own and
    seeds the rest from the finished lists (tc2_make_plan: first_items).  Same answers as the unseeded single launch
    (DEWI_FLAG_NO_SEED), exact ties at the seed included, and the oracle's."""
    n, d, k = 700_000, 64, 10
    emb, pay = make_corpus(n, d, seed=241)
    emb[3000:3030] = emb[77]            # 31 identical rows inside the first chunks ...
    emb[650_000:650_025] = emb[77]      # ... and in the last ones
    queries = np.random.RandomState(242).standard_normal((b, d)).astype(np.float32)
    queries[5] = emb[77] + 0.01 * queries[5]
    rows = emb if dtype == "fp32" else bf16_round(emb)
    ix = bulk_index(emb, pay, dtype=dtype)
    force = _native.FLAG_FORCE_TC | (_native.FLAG_FORCE_CERT if dtype == "fp32" else 0)
    n0 = ix._backend.last_launches()
    ids_s, sc_s = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=force)
    launches_staged = ix._backend.last_launches()
    ids_u, sc_u = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=force | _native.FLAG_NO_SEED)
    launches_plain = ix._backend.last_launches()
    assert launches_staged >= launches_plain + 3, "pre-pass + seed + a second sweep launch were expected"
    sel = [q for q in range(b) if q != 5]
    np.testing.assert_array_equal(ids_s[sel], ids_u[sel])
    np.testing.assert_array_equal(sc_s[sel], sc_u[sel])
    tied = {77, *range(3000, 3030), *range(650_000, 650_025)}
    assert set(ids_s[5].tolist()) <= tied and set(ids_u[5].tolist()) <= tied
    pick = np.linspace(0, b - 1, 12).astype(int)
    pick = pick[pick != 5]
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries[pick], k, 0.3, 0.5, True)
    if dtype == "fp32":
        for i, q in enumerate(pick):
            check_topk(rid[i], rsc[i], ids_s[q], sc_s[q], what=f"staged q{q}")
    else:
        assert recall_at_k(rid, ids_s[pick]) >= 0.999


def test_many_query_pairs_with_long_lists_skip_the_staged_sweep():
    """k = 60 makes the candidate lists of one staged first round (18 chunks x 136 entries at B = 1024) longer than the
    2048-entry window the seed selection ranks in shared memory: tc2_make_plan must keep the single launch instead of
    planning a first round launch_seed_from_partials would refuse (the search used to fail with an error)."""
    n, d, k, b = 120_000, 64, 60, 1024
    emb, pay = make_corpus(n, d, seed=251)
    queries = np.random.RandomState(252).standard_normal((b, d)).astype(np.float32)
    ix = bulk_index(emb, pay, dtype="bf16")
    ix._backend.set_profiling(True)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5, flags=_native.FLAG_FORCE_TC)
    assert ix._backend.sweep_ms()[1] == "tcgen05-pair"
    pick = np.linspace(0, b - 1, 8).astype(int)
    rid, _ = osearch.exact_search_batch(bf16_round(emb), pay[:, 0], entropy_column(pay), queries[pick], k, 0.3, 0.5, True)
    assert recall_at_k(rid, ids[pick]) >= 0.995   # (480 ids: bf16 query rounding may swap a pair at the 2k-th boundary)
    assert np.all(np.diff(sc, axis=1) <= 0)


@pytest.mark.parametrize("dtype,n,d,k,b", [("fp32", 5000, 128, 100, 3), ("fp32", 5000, 128, 200, 3), ("bf16", 9000, 64, 200, 3),
                                            ("fp32", 300, 100, 150, 3), ("fp32", 5000, 128, 150, 8), ("bf16", 9000, 64, 120, 40),
                                            ("bf16", 6000, 128, 110, 100), ("fp32", 6000, 64, 160, 300)])
def test_large_k_up_to_the_documented_limit(dtype, n, d, k, b):
    """Up to 400 candidates (min(2k, N)) per query: long lists take the exact CUDA-core sweep (the tensor-core sweeps'
    shared memory holds ~150-220) and the fused tail; checked against the oracle like every other search.  k = 110-160
    used to plan an M = 64 sweep on 128-row tiles that has no kernel (tc_make_plan) -- with few queries, with one
    query block and with CTA pairs."""
    emb, pay = make_corpus(n, d, seed=261)
    queries = np.random.RandomState(262).standard_normal((b, d)).astype(np.float32)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    ix = bulk_index(emb, pay, dtype=dtype)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries, k, 0.3, 0.5, True)
    for i in range(len(queries)):
        check_topk(rid[i], rsc[i], ids[i], sc[i], what=f"k={k} q{i}")
    res = ix.search(queries[0], k=k, eta=0.3, entropy_pref=0.5)
    assert len(res) == k and [int(r[0][4:]) for r in res] == ids[0].tolist()


def test_k_above_the_candidate_limit_raises():
    emb, pay = make_corpus(3000, 64, seed=263)
    ix = bulk_index(emb, pay, dtype="fp32")
    q = np.random.RandomState(264).standard_normal((2, 64)).astype(np.float32)
    with pytest.raises(ValueError, match="400 candidates"):
        ix.search_batch(q, k=201)
    with pytest.raises(ValueError, match="400 candidates"):
        ix.search(q[0], k=201)
    # the library refuses as well (a caller that binds the C ABI directly)
    be = ix._backend
    with pytest.raises(_native.NativeError, match="at most 400 candidates"):
        be._search_host(q, 201, 0.3, 0.5)


def _shape_sweep():
    """Edge shapes of the plan selection (tile / query-block / minimum-row boundaries) + seeded random ones."""
    shapes = [(2047, 64, 1, 10), (2048, 64, 1, 10), (2049, 128, 64, 10), (4097, 64, 65, 3), (16385, 128, 128, 10),
              (8191, 192, 129, 7), (33000, 64, 256, 10), (12345, 256, 257, 10), (70001, 64, 513, 5), (1, 64, 3, 1),
              (31, 128, 2, 20), (5000, 1024, 9, 10), (40000, 320, 31, 40), (9000, 64, 1025, 2), (2500, 448, 700, 10)]
    rng = np.random.RandomState(271)
    for _ in range(14):
        shapes.append((int(rng.randint(1, 60000)), int(rng.choice([64, 128, 192, 256, 384, 100, 72])), int(rng.randint(1, 600)),
                       int(rng.randint(1, 48))))
    return shapes


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_shape_sweep_default_path_selection_vs_oracle(dtype):
    """No forcing flags: whatever sweep the planner picks for the shape (rows-on-M, queries-on-M, CTA pairs, staged,
    certified, CUDA cores) must reproduce the oracle -- shapes sit on the planner's boundaries (2048 rows, 64 / 128 /
    256 queries, ragged tiles, dims that are not a multiple of 64, one-row corpora)."""
    failures = []
    for si, (n, d, b, k) in enumerate(_shape_sweep()):
        k = min(k, n)
        what = f"shape {si}: n={n} d={d} B={b} k={k} {dtype}"
        try:
            _check_shape(si, n, d, b, k, dtype, what)
        except Exception as exc:  # noqa: BLE001 -- every shape is reported, not only the first that fails
            failures.append(f"{what}: {type(exc).__name__}: {str(exc)[:300]}")
    assert not failures, "\n".join(failures)


def _check_shape(si, n, d, b, k, dtype, what):
    emb, pay = make_corpus(n, d, seed=300 + si)
    queries = np.random.RandomState(400 + si).standard_normal((b, d)).astype(np.float32)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    ix = bulk_index(emb, pay, dtype=dtype)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    assert ids.shape == (b, k)
    pick = np.unique(np.linspace(0, b - 1, min(b, 24)).astype(int))
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries[pick], k, 0.3, 0.5, True)
    if dtype == "fp32":
        for i, q in enumerate(pick):
            check_topk(rid[i], rsc[i], ids[q], sc[q], what=f"{what} q{q}")
    else:
        # bf16 rows: the tensor-core sweeps round the query too (one bf16 plane + exact re-score of the candidates)
        assert recall_at_k(rid, ids[pick]) >= 0.995, what
        assert np.max(np.abs(np.sort(sc[pick], axis=1) - np.sort(rsc, axis=1))) <= 2e-3, what
