"""The oracle (CPU restatement) against the fixtures produced by the unmodified reference
(oracle/make_golden.py).  Runs without a GPU; this is what pins the oracle (BASELINE tier rule 3)."""

import json

import numpy as np
import pytest

from oracle import redundancy as ored
from oracle import scorer as oscorer
from oracle import search as osearch

from _util import GOLD, entropy_column, load_search_golden

SEARCH_CASES = ["cos_n100_d128", "cos_n2000_d64", "cos_n3000_d768", "cos_n15_d16_k5", "cos_n12_d8_k10", "l2_n500_d32"]
SCORER_CASES = ["readme_n1001", "readme_n1000", "profile_n4096"]


def test_manifest_lists_every_fixture():
    man = json.loads((GOLD / "MANIFEST.json").read_text())
    assert len(man["cases"]) == 15
    # 15 fixtures written from the reference by oracle/make_golden.py + 2 query sets of the CUDA-saved directories
    # (scripts/make_cuda_saved_fixture.py, run on a B200)
    assert len([p for p in GOLD.glob("*.npz") if not p.name.startswith("cuda_saved_index")]) == 15
    assert (GOLD / "reference_saved_index" / "ann_index" / "embeddings.npy").exists()


@pytest.mark.parametrize("name", SEARCH_CASES)
def test_search_oracle_bit_exact(name):
    g = load_search_golden(name)
    cosine = g["space"] == "cosine"
    rows = osearch.normalize_rows_like_add(g["emb"]) if cosine else g["emb"]
    ent = entropy_column(g["payload"])
    for gi, (eta, pref) in enumerate(g["grid"]):
        for qi, q in enumerate(g["queries"]):
            idx, sc = osearch.exact_search(rows, g["payload"][:, 0], ent, q, g["k"], eta, pref, cosine)
            assert np.array_equal(idx, g["ref_idx"][gi, qi])
            assert np.array_equal(sc, g["ref_scores"][gi, qi])


def test_search_oracle_object_form_and_errors():
    g = load_search_golden("cos_n15_d16_k5")

    class P:
        def __init__(self, row):
            self.dewi, self.ht_mean, self.hi_mean = float(row[0]), float(row[1]), float(row[3])

    ix = osearch.OracleExactIndex(16)
    with pytest.raises(ValueError):
        ix.build()
    for i, e in enumerate(g["emb"]):
        ix.add(f"doc_{i:08d}", e, P(g["payload"][i]))
    with pytest.raises(ValueError):
        ix.add("bad", np.zeros(3, np.float32), P(g["payload"][0]))
    ix.build()
    eta, pref = g["grid"][0]
    res = ix.search(g["queries"][0], k=g["k"], eta=eta, entropy_pref=pref)
    assert [int(r[0][4:]) for r in res] == g["ref_idx"][0, 0].tolist()
    with pytest.raises(ValueError):  # k > N (backends.py:468)
        ix.search(g["queries"][0], k=16)


@pytest.mark.parametrize("name", SCORER_CASES)
def test_scorer_oracle_bit_exact(name):
    g = np.load(GOLD / f"scorer_{name}.npz")
    cols = {k: g["signals"][i] for i, k in enumerate(oscorer.SIGNAL_KEYS)}
    med, mad = oscorer.robust_fit(cols)
    assert [med[k] for k in oscorer.SIGNAL_KEYS] == g["med"].tolist()
    assert [mad[k] for k in oscorer.SIGNAL_KEYS] == g["mad"].tolist()
    w = tuple(g["weights"])
    assert np.array_equal(oscorer.score_rows(cols, med, mad, w, False), g["score"])
    assert np.array_equal(oscorer.score_rows(cols, med, mad, w, True), g["score_conditional"])


def test_scorer_oracle_one_row_zero_mad():
    g = np.load(GOLD / "scorer_onerow.npz")
    sig = dict(zip([str(k) for k in g["keys"]], g["row"].tolist()))
    w = g["weights"]
    o = oscorer.OracleScorer(w[:5], w[5])
    o.fit_stats([sig])
    assert all(o.mad[k] == 1e-8 for k in sig)  # scorer.py:24 `or 1e-8`
    assert o.score(sig) == float(g["score"]) and o.score_conditional(sig) == float(g["score_conditional"])


@pytest.mark.parametrize("name", ["t37_i53_d512", "t64_i64_d64"])
def test_redundancy_oracle(name):
    g = np.load(GOLD / f"redundancy_{name}.npz")
    sim = ored.cross_modal_similarity(g["tfeat"], g["ifeat"])
    assert sim.shape == g["sim"].shape and sim.dtype == np.float32
    np.testing.assert_allclose(sim, g["sim"], atol=5e-7, rtol=0)
    mx, am, cnt, pairs = ored.join_rowstats(g["tfeat"], g["ifeat"], 0.9)
    np.testing.assert_array_equal(am, np.argmax(sim, axis=1))
    assert cnt.sum() == len(pairs) == int((sim >= np.float32(0.9)).sum()) > 0


def test_quickstart_flow_oracle_bit_exact():
    """BASELINE.json config 1 (README quick-start at 10K documents): fit -> score -> search restated by the oracle
    equals what the unmodified reference produced for the fixture, bit for bit."""
    import hashlib

    g = np.load(GOLD / "quickstart_c1.npz")
    n, d, k = int(g["n"]), int(g["d"]), int(g["k"])
    rng = np.random.RandomState(int(g["seed"]))
    emb = rng.rand(n, d).astype(np.float32)
    assert hashlib.sha256(emb.tobytes()).hexdigest() == str(g["emb_sha256"])
    sig = g["signals"]
    cols = {key: sig[j] for j, key in enumerate(oscorer.SIGNAL_KEYS)}
    med, mad = oscorer.robust_fit(cols)
    assert [med[key] for key in oscorer.SIGNAL_KEYS] == g["med"].tolist()
    assert [mad[key] for key in oscorer.SIGNAL_KEYS] == g["mad"].tolist()
    dewi = oscorer.score_rows(cols, med, mad)
    assert np.array_equal(dewi, g["dewi"])
    rows = osearch.normalize_rows_like_add(emb)
    ent = (sig[0].astype(np.float64) + sig[2].astype(np.float64)) * 0.5
    for qi, q in enumerate(g["queries"]):
        idx, sc = osearch.exact_search(rows, dewi, ent, q, k, 0.3, 0.5, True)
        assert np.array_equal(idx, g["ref_idx"][qi]) and np.array_equal(sc, g["ref_scores"][qi])
