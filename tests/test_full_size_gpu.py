"""Parity at BASELINE.json's full sizes.

C2 (1M x 768 fp32, exact): compared directly with the oracle (`ExactIndex.search` restated) -- the oracle
needs ~60 ms per query at this size, so a handful of queries run in seconds.
C3 (100M x 768 bf16, one GPU) and C4 (100M Signals rows): too big for the oracle, so they are checked through
size-independent properties: planted needles at tile / chunk / shard boundaries must come back first,
returned similarities must equal an exact recomputation from the stored rows, every slab of the corpus
searched exactly on its own must be dominated by the global answer, all sweep kernels (M = 64, M = 128,
CTA pairs) must agree, and the scorer's medians must split the column exactly in half.
"""

import numpy as np
import pytest
import torch

import dewi_b200
from oracle import scorer as oscorer
from oracle import search as osearch

from _util import check_topk, entropy_column, synth_payload_columns

pytestmark = pytest.mark.gpu

DIM = 768
CHUNK = 1_000_000


def _gen_chunk(chunk: int, rows: int, device):
    g = torch.Generator(device=device)
    g.manual_seed(77_000_000 + chunk)
    emb = torch.randn((rows, DIM), generator=g, device=device, dtype=torch.float32)
    pay = torch.rand((2, rows), generator=g, device=device, dtype=torch.float32)
    return emb, pay[0].contiguous(), (pay[1] * 3.0).contiguous()


def test_c2_full_size_fp32_equals_the_oracle():
    """1M x 768 fp32: identical top-10 ids (outside exact ties) and scores within 1e-5, single-query API
    and a small batch, against the exact numpy path on the same rows."""
    n, k = 1_000_000, 10
    dev = torch.device("cuda", 0)
    emb_d, _, _ = _gen_chunk(0, n, dev)
    emb = emb_d.cpu().numpy()
    del emb_d
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)  # backends.py:403-405, in bulk
    rng = np.random.RandomState(11)
    pay = synth_payload_columns(rng, n, "profile")
    ent = entropy_column(pay)
    ix = dewi_b200.CudaIndex(DIM, dtype="fp32", device=0)
    ix.add_batch(None, emb, normalized=True)
    ix.set_payload_columns(pay[:, 0].astype(np.float32), ent.astype(np.float32))
    ix.build()
    queries = rng.standard_normal((6, DIM)).astype(np.float32)
    for eta, pref in [(0.3, 0.5), (0.0, 0.0)]:
        ids, sc = ix.search_batch(queries, k=k, eta=eta, entropy_pref=pref)
        for qi in range(len(queries)):
            rid, rsc = osearch.exact_search(emb, pay[:, 0], ent, queries[qi], k, eta, pref, True)
            check_topk(rid, rsc, ids[qi], sc[qi], what=f"C2 eta{eta} pref{pref} q{qi}")
    one = ix.search(queries[0], k=k, eta=0.3, entropy_pref=0.5)
    rid, rsc = osearch.exact_search(emb, pay[:, 0], ent, queries[0], k, 0.3, 0.5, True)
    check_topk(rid, rsc, [int(t[0].split("_")[1]) for t in one], [t[1] for t in one], what="C2 single-query API")


def test_c3_full_size_bf16_properties():
    """100M x 768 bf16 on one GPU (153.6 GB)."""
    torch.cuda.empty_cache()   # blocks cached by earlier tests in this process would count against the 180 GB
    free, _ = torch.cuda.mem_get_info(0)
    if free < 168e9:
        pytest.skip("needs ~165 GB of free device memory")
    n, k = 100_000_000, 10
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(3)
    n_q = 12
    queries = rng.standard_normal((n_q, DIM)).astype(np.float32)
    qn = queries / np.linalg.norm(queries, axis=1, keepdims=True)
    # needles: query j's own direction planted at positions that straddle every boundary the sweep has
    # (first / last row, MMA tile edges, generation chunks, the 8-way shard cuts of the multi-GPU layout)
    needle_rows = [0, 255, 256, 12_499_999, 12_500_000, 37_500_001, 49_999_999, 50_000_000, 87_654_321, n - 257, n - 2, n - 1]
    needles = dict(zip(needle_rows, range(n_q)))
    ix = dewi_b200.CudaIndex(DIM, dtype="bf16", device=0)
    ix.reserve(n)
    for c in range(n // CHUNK):
        emb, dewi, ent = _gen_chunk(c, CHUNK, dev)
        for row, j in needles.items():
            if c * CHUNK <= row < (c + 1) * CHUNK:
                emb[row - c * CHUNK] = torch.from_numpy(queries[j] * 2.5).to(dev)  # any positive scale: rows are normalised
        ix.add_batch(None, emb, normalized=False)
        ix.set_payload_columns(dewi, ent, offset=c * CHUNK)
        del emb, dewi, ent
    ix.build()
    assert len(ix) == n

    # (1) needles first, with similarity 1 up to bf16 rounding of the stored row (rel. 2^-9 per element, averaged out)
    ids, sc = ix.search_batch(queries, k=k, eta=0.0, entropy_pref=0.0)
    for row, j in needles.items():
        assert ids[j][0] == row, f"needle of query {j} at row {row} not returned first: {ids[j]}"
        assert abs(sc[j][0] - 1.0) < 2e-3
    assert all(len(set(r.tolist())) == k and np.all(np.diff(s) <= 0) for r, s in zip(ids, sc))

    # (2) returned similarities == exact fp32 dot of the STORED (bf16) row with the fp32-normalised query
    for j in range(n_q):
        for row, s in zip(ids[j], sc[j]):
            exact = float(np.dot(ix.get_row(int(row)).astype(np.float64), qn[j].astype(np.float64)))
            assert abs(exact - s) <= 1e-5, f"query {j} row {row}: returned {s}, recomputed {exact}"

    # (3) the ORACLE on slabs of the stored rows.  A 1M-row slab is pulled back from the device (bulk export of the
    # bf16 rows, widened exactly) and handed to `oracle.search.exact_search` (numpy, backends.py:414-481) together
    # with the rows of the query's global top-2k candidates: the global top-2k by similarity is the top-2k of any
    # subset that contains it, so the oracle over {slab + candidates} must return the GPU's global answer -- and it
    # would surface any slab row the sweep wrongly dropped.  Checked with the DEWI re-rank on (eta, entropy_pref).
    rid_g, rsc_g = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    qd = torch.from_numpy(queries).to(dev)
    c_sim, c_id, c_dewi, c_ent = (t.cpu().numpy() for t in ix.search_local(qd, 2 * k))
    cand_rows = np.stack([np.stack([ix.get_row(int(r)) for r in c_id[j]]) for j in range(n_q)])
    for c in (0, 37, 99):
        slab = ix.export_rows(c * CHUNK, CHUNK)
        s_dewi, s_ent = ix.get_payload_columns(c * CHUNK, CHUNK)
        for j in range(n_q):
            out_of_slab = [t for t in range(2 * k) if not (c * CHUNK <= c_id[j, t] < (c + 1) * CHUNK)]
            emb_j = np.concatenate([slab, cand_rows[j, out_of_slab]])
            gid_j = np.concatenate([c * CHUNK + np.arange(CHUNK, dtype=np.int64), c_id[j, out_of_slab]])
            dewi_j = np.concatenate([s_dewi, c_dewi[j, out_of_slab]])
            ent_j = np.concatenate([s_ent, c_ent[j, out_of_slab]]).astype(np.float64)
            oi, osc = osearch.exact_search(emb_j, dewi_j, ent_j, queries[j], k, 0.3, 0.5, True)
            check_topk(gid_j[oi], osc, rid_g[j], rsc_g[j], what=f"slab {c} q{j} vs the oracle")
        del slab

    # (4) all sweep kernels agree on the re-ranked answer: B = 1 (single-query API), B = 12 (M = 64 MMAs),
    # B = 100 (M = 128), B = 200 (CTA pairs); padding queries are copies
    ref_ids, ref_sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    for b in (100, 200):
        big = np.concatenate([queries] * (b // n_q + 1))[:b]
        bid, bsc = ix.search_batch(big, k=k, eta=0.3, entropy_pref=0.5)
        for j in range(b):
            check_topk(ref_ids[j % n_q], ref_sc[j % n_q], bid[j], bsc[j], what=f"B={b} q{j}")
    one = ix.search(queries[5], k=k, eta=0.3, entropy_pref=0.5)
    check_topk(ref_ids[5], ref_sc[5], [int(t[0].split("_")[1]) for t in one], [t[1] for t in one], what="B=1")


def test_c4_full_size_scorer_properties():
    """100M Signals rows x 7: the medians / MADs are exact order statistics (the counts below and above
    split the column in half), they equal the oracle on a column the oracle can still handle, and the
    scores of a row sample equal the oracle's."""
    n = 100_000_000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    hi = torch.tensor([10, 15, 5, 8, 1, 1, 0.2], device=dev).view(7, 1)
    sig = torch.rand((7, n), generator=g, device=dev) * hi
    s = dewi_b200.DewiScorer()
    s.fit_stats_columns(sig)
    med = np.array([s.stats.medians[k] for k in oscorer.SIGNAL_KEYS])
    mad = np.array([s.stats.mads[k] for k in oscorer.SIGNAL_KEYS])
    for c in range(7):
        col = sig[c]
        m = float(med[c])
        below, above = int((col < m).sum()), int((col > m).sum())
        assert below <= n // 2 and above <= n // 2, f"column {c}: {below} below / {above} above the median"
        dev_c = (col - np.float32(m)).abs()
        below, above = int((dev_c < float(mad[c])).sum()), int((dev_c > float(mad[c])).sum())
        assert below <= n // 2 and above <= n // 2, f"column {c}: MAD is not the median deviation"
    # the oracle on one full column (np.median of 100M floats takes a few seconds)
    col0 = sig[0].cpu().numpy()
    omed, omad = oscorer.robust_fit({"x": col0})
    assert med[0] == omed["x"] and mad[0] == omad["x"]
    # scores of a strided row sample against the oracle, 1e-6 relative
    out = s.score_batch(sig)
    pick = torch.arange(0, n, 9973, device=dev)
    cols = {k: sig[i, pick].cpu().numpy() for i, k in enumerate(oscorer.SIGNAL_KEYS)}
    ref = oscorer.score_rows(cols, dict(s.stats.medians), dict(s.stats.mads))
    got = out[pick].cpu().numpy().astype(np.float64)
    assert np.max(np.abs(got - ref) / ref) <= 1e-6


def test_bf16_recall_gate_at_10m_rows_against_the_oracle_on_the_host():
    """bf16 storage, 10M x 768: recall@10 >= 0.999 against `oracle.search.exact_search` run on the HOST over the same
    bf16-representable rows (exported in bulk from the device, 30.7 GB fp32), 128 fp32 queries, DEWI re-rank on."""
    import psutil

    n, k, n_q = 10_000_000, 10, 128
    if psutil.virtual_memory().available < 44e9:
        pytest.skip("needs ~40 GB of free host memory for the fp32 oracle corpus")
    torch.cuda.empty_cache()
    dev = torch.device("cuda", 0)
    ix = dewi_b200.CudaIndex(DIM, dtype="bf16", device=0)
    ix.reserve(n)
    for c in range(n // CHUNK):
        emb, dewi, ent = _gen_chunk(500 + c, CHUNK, dev)
        ix.add_batch(None, emb, normalized=False)
        ix.set_payload_columns(dewi, ent, offset=c * CHUNK)
        del emb, dewi, ent
    ix.build()
    rows = ix.export_rows(0, n)                      # what the device stores, widened exactly
    dewi_h, ent_h = ix.get_payload_columns(0, n)
    queries = np.random.RandomState(17).standard_normal((n_q, DIM)).astype(np.float32)
    ids, sc = ix.search_batch(queries, k=k, eta=0.3, entropy_pref=0.5)
    ent64 = ent_h.astype(np.float64)
    hits, worst = 0, 0.0
    for j in range(n_q):
        rid, rsc = osearch.exact_search(rows, dewi_h, ent64, queries[j], k, 0.3, 0.5, True)
        hits += len(set(rid.tolist()) & set(ids[j].tolist()))
        same = rid == ids[j]
        if same.any():
            worst = max(worst, float(np.max(np.abs(rsc[same] - sc[j][same]) / np.maximum(1.0, np.abs(rsc[same])))))
    recall = hits / (n_q * k)
    assert recall >= 0.999, f"bf16 recall@{k} = {recall:.5f} at {n} rows"
    assert worst <= 1e-5
