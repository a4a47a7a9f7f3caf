#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference, and pin the oracle to it.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

For every case the script (1) drives the reference's own classes
(`dewi.index.DewiIndex(use_ann=False)` -> `ExactIndex`, `dewi.scorer.DewiScorer`) and torch's
`F.normalize` + matmul as in `dewi/signals/redundancy.py:36-38`, (2) asserts that the numpy
restatement in `oracle/` reproduces those outputs bit-for-bit, and (3) stores inputs + reference
outputs as small fixtures.  The fixtures travel to the GPU box; the reference does not.
"""

from __future__ import annotations

import hashlib
import json
import logging
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path(os.environ.get("DEWI_REFERENCE_SRC", "/root/reference/src"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF_SRC))
logging.disable(logging.WARNING)

from dewi.index import DewiIndex  # noqa: E402  (reference)
from dewi.scorer import DewiScorer  # noqa: E402  (reference)
from dewi.types import Payload, Weights  # noqa: E402  (reference)

from oracle import search as osearch  # noqa: E402
from oracle import scorer as oscorer  # noqa: E402
from oracle import redundancy as ored  # noqa: E402

GOLD = ROOT / "tests" / "golden"
PAYLOAD_FIELDS = ("dewi", "ht_mean", "ht_q90", "hi_mean", "hi_q90", "I_hat", "redundancy", "noise")


def synth_payload_columns(rng: np.random.RandomState, n: int, style: str) -> np.ndarray:
    """[n, 8] float64 whose values are float32-representable."""
    if style == "profile":  # scripts/profile_index.py:58-70
        cols = [
            np.clip(rng.beta(2, 2, n), 0, 1), rng.gamma(2, 0.5, n), rng.gamma(2, 0.5, n) * 1.5,
            rng.gamma(2, 0.3, n), rng.gamma(2, 0.3, n) * 1.5, rng.beta(2, 2, n), rng.beta(1, 5, n), rng.beta(1, 10, n),
        ]
    else:  # README.md:83-91 ranges
        cols = [
            rng.uniform(0, 1, n), rng.uniform(0, 10, n), rng.uniform(0, 15, n), rng.uniform(0, 5, n),
            rng.uniform(0, 8, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 0.2, n),
        ]
    return np.stack(cols, axis=1).astype(np.float32).astype(np.float64)


def search_case(name, n, d, nq, k, space, style, seed, grid):
    rng = np.random.RandomState(seed)
    emb = rng.randn(n, d).astype(np.float32)
    if space == "l2":
        emb *= 0.25
    pay = synth_payload_columns(rng, n, style)
    queries = rng.randn(nq, d).astype(np.float32)
    ids = [f"doc_{i:08d}" for i in range(n)]

    ref = DewiIndex(dim=d, space=space, use_ann=False)
    payloads = [Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}) for i in range(n)]
    for i in range(n):
        ref.add(ids[i], emb[i], payloads[i])
    ref.build()
    stored = np.asarray(ref._backend._embeddings)

    # oracle restatement of add(): must store the same rows
    mine_rows = osearch.normalize_rows_like_add(emb) if space == "cosine" else emb
    assert np.array_equal(mine_rows, stored), f"{name}: add() restatement differs"
    ent = (pay[:, 1] + pay[:, 3]) * 0.5

    out_idx = np.zeros((len(grid), nq, k), dtype=np.int64)
    out_sc = np.zeros((len(grid), nq, k), dtype=np.float32)
    for g, (eta, pref) in enumerate(grid):
        for qi in range(nq):
            res = ref.search(queries[qi], k=k, eta=eta, entropy_pref=pref)
            ridx = np.array([int(r[0][4:]) for r in res], dtype=np.int64)
            rsc = np.array([r[1] for r in res], dtype=np.float32)
            oidx, osc = osearch.exact_search(stored, pay[:, 0], ent, queries[qi], k, eta, pref, space == "cosine")
            assert np.array_equal(ridx, oidx), f"{name}: ids differ (grid {g}, q {qi})"
            assert np.array_equal(rsc, osc), f"{name}: scores differ (grid {g}, q {qi})"
            out_idx[g, qi], out_sc[g, qi] = ridx, rsc
    # Large corpora are stored as (seed, sha256): RandomState's stream is frozen across numpy
    # versions, so tests regenerate the rows with `regen_search_inputs` and verify the digest.
    big = emb.nbytes > (1 << 20)
    np.savez_compressed(
        GOLD / f"search_{name}.npz", emb=(np.zeros((0, d), np.float32) if big else emb),
        emb_sha256=np.array(hashlib.sha256(emb.tobytes()).hexdigest()), seed=np.int64(seed), n=np.int64(n),
        style=np.array(style), payload=pay, queries=queries,
        grid=np.asarray(grid, dtype=np.float64), k=np.int64(k), space=np.array(space),
        ref_idx=out_idx, ref_scores=out_sc,
    )
    return {"case": name, "n": n, "d": d, "nq": nq, "k": k, "space": space, "grid": len(grid)}


def scorer_case(name, n, style, seed, weights, delta):
    rng = np.random.RandomState(seed)
    pay = synth_payload_columns(rng, n, style)
    cols = {f: pay[:, j] for j, f in enumerate(PAYLOAD_FIELDS) if f != "dewi"}
    rows = [{f: float(cols[f][i]) for f in cols} for i in range(n)]
    ref = DewiScorer(Weights(*weights), delta=delta)
    ref.fit_stats(rows)
    med = np.array([ref.stats.medians[f] for f in oscorer.SIGNAL_KEYS])
    mad = np.array([ref.stats.mads[f] for f in oscorer.SIGNAL_KEYS])
    sc = np.array([ref.score(r) for r in rows])
    scc = np.array([ref.score_conditional(r) for r in rows])

    omed, omad = oscorer.robust_fit(cols)
    assert all(omed[f] == ref.stats.medians[f] and omad[f] == ref.stats.mads[f] for f in cols), f"{name}: fit differs"
    w6 = tuple(weights) + (delta,)
    osc = oscorer.score_rows(cols, omed, omad, w6, conditional=False)
    oscc = oscorer.score_rows(cols, omed, omad, w6, conditional=True)
    assert np.array_equal(osc, sc) and np.array_equal(oscc, scc), f"{name}: vectorised score differs"
    np.savez_compressed(
        GOLD / f"scorer_{name}.npz", signals=np.stack([cols[f] for f in oscorer.SIGNAL_KEYS], axis=0),
        weights=np.asarray(w6), med=med, mad=mad, score=sc, score_conditional=scc,
    )
    return {"case": name, "n": n, "style": style}


def scorer_degenerate():
    """tests/test_scorer_weights.py:6-14: one row, MAD == 0 -> 1e-8, extra key `dewi` is fitted too."""
    w = Weights(alpha_t=0.6, alpha_i=0.2, alpha_r=0.2, alpha_n=0.1)
    s = DewiScorer(weights=w)
    p = Payload(ht_mean=1.0, hi_mean=0.5, redundancy=0.2, noise=0.1, ht_q90=1.2, hi_q90=0.7)
    sig = p.to_dict()
    sig["I_hat"] = 0.0
    s.fit_stats([sig])
    keys = list(sig.keys())
    np.savez_compressed(
        GOLD / "scorer_onerow.npz", keys=np.array(keys), row=np.array([sig[k] for k in keys]),
        med=np.array([s.stats.medians[k] for k in keys]), mad=np.array([s.stats.mads[k] for k in keys]),
        weights=np.array([w.alpha_t, w.alpha_i, w.alpha_m, w.alpha_r, w.alpha_n, w.delta]),
        score=np.float64(s.score(sig)), score_conditional=np.float64(s.score_conditional(sig)),
    )
    o = oscorer.OracleScorer((w.alpha_t, w.alpha_i, w.alpha_m, w.alpha_r, w.alpha_n), w.delta)
    o.fit_stats([sig])
    assert o.score(sig) == s.score(sig) and o.score_conditional(sig) == s.score_conditional(sig)
    return {"case": "onerow"}


def redundancy_case(name, t, i, d, seed):
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(seed)
    tf = (rng.randn(t, d) * rng.uniform(0.5, 3.0, (t, 1))).astype(np.float32)
    imf = (rng.randn(i, d) * rng.uniform(0.5, 3.0, (i, 1))).astype(np.float32)
    # plant near-duplicates so a threshold has true positives
    for j in range(0, min(t, i), 7):
        imf[j] = tf[j] * 1.7 + 0.05 * rng.randn(d).astype(np.float32)
    with torch.no_grad():  # redundancy.py:36-38
        a = F.normalize(torch.from_numpy(tf), p=2, dim=1)
        b = F.normalize(torch.from_numpy(imf), p=2, dim=1)
        sim = (a @ b.T).cpu().numpy()
    mine = ored.cross_modal_similarity(tf, imf)
    err = float(np.max(np.abs(mine - sim)))
    assert err <= 5e-7, f"{name}: similarity restatement off by {err}"
    np.savez_compressed(GOLD / f"redundancy_{name}.npz", tfeat=tf, ifeat=imf, sim=sim)
    return {"case": name, "t": t, "i": i, "d": d, "oracle_vs_torch_maxabs": err}


def extras_case():
    """robust.RobustStats.from_payloads / .z and local_weights_from_surprisal of the reference."""
    from dewi.local_weights import local_weights_from_surprisal
    from dewi.robust import RobustStats as PayloadStats

    rng = np.random.RandomState(51)
    pay = synth_payload_columns(rng, 777, "profile")
    payloads = [Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}) for i in range(len(pay))]
    st = PayloadStats.from_payloads(payloads)
    keys = ["ht_mean", "hi_mean", "redundancy", "noise"]
    probe = {k: float(np.float32(pay[5, PAYLOAD_FIELDS.index(k)])) for k in keys}
    surpr = np.concatenate([rng.gamma(2.0, 1.5, 4099), [0.0, 50.0, -3.0]]).astype(np.float32)
    const = np.full(17, 2.5, np.float32)  # MAD == 0 path
    np.savez_compressed(
        GOLD / "extras_robust_localweights.npz", payload=pay,
        med=np.array([st.fields[k][0] for k in keys]), mad=np.array([st.fields[k][1] for k in keys]),
        probe=np.array([probe[k] for k in keys]), z=np.array([st.z(k, probe[k]) for k in keys]),
        surprisal=surpr, weights=local_weights_from_surprisal(surpr),
        const=const, const_weights=local_weights_from_surprisal(const),
    )
    return {"case": "extras_robust_localweights"}


def saved_index_case():
    """A directory written by the reference's own DewiIndex.save (index.py:121-141 -> ExactIndex.save,
    backends.py:483-515), to be loaded by the CUDA backend."""
    import shutil

    rng = np.random.RandomState(61)
    n, d, k = 40, 16, 5
    emb = rng.randn(n, d).astype(np.float32)
    pay = synth_payload_columns(rng, n, "profile")
    ref = DewiIndex(dim=d, space="cosine", use_ann=False, rerank_eta=0.3, entropy_pref=0.5)
    for i in range(n):
        ref.add(f"doc_{i:03d}", emb[i], Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}),
                meta={"source": f"file_{i}.txt"})
    ref.build()
    out = GOLD / "reference_saved_index"
    shutil.rmtree(out, ignore_errors=True)
    ref.save(out)
    q = rng.randn(3, d).astype(np.float32)
    res = [ref.search(q[i], k=k) for i in range(3)]
    np.savez_compressed(GOLD / "reference_saved_index_queries.npz", queries=q,
                        ids=np.array([[r[0] for r in rr] for rr in res]), scores=np.array([[r[1] for r in rr] for rr in res]))
    return {"case": "reference_saved_index", "n": n, "d": d}


def quickstart_case():
    """BASELINE.json config 1: the README quick-start (README.md:67-110) at 10K documents, driven through the
    reference exactly as the README writes it -- add with `dewi=0.0`, `fit_stats(rows)`, per-document
    `payload.dewi = scorer.score(signals)`, `build()`, `search(k=10, eta=0.3, entropy_pref=0.5)`.  `Signals` does
    not exist in the reference's code, so rows are the plain dicts `fit_stats` takes;
    hnswlib is not installed, so the exact backend answers (`use_ann=False`)."""
    rng = np.random.RandomState(101)
    n, d, k, nq = 10_000, 768, 10, 8
    emb = rng.rand(n, d).astype(np.float32)                      # README.md:76 `np.random.rand(768)`
    hi = (10, 15, 5, 8, 1, 1, 0.2)                               # README.md:83-91
    sig = np.stack([rng.uniform(0, h, n) for h in hi]).astype(np.float32)
    queries = rng.rand(nq, d).astype(np.float32)
    index = DewiIndex(dim=d, space="cosine", use_ann=False)
    rows = []
    for i in range(n):
        signals = {key: float(sig[j, i]) for j, key in enumerate(oscorer.SIGNAL_KEYS)}
        rows.append(signals)
        index.add(f"doc{i}", emb[i], Payload(dewi=0.0, **signals))
    scorer = DewiScorer(Weights())
    scorer.fit_stats(rows)
    for i in range(n):
        payload = index.get_payload(f"doc{i}")
        payload.dewi = scorer.score({key: getattr(payload, key) for key in oscorer.SIGNAL_KEYS})
    index.build()
    dewi = np.array([index.get_payload(f"doc{i}").dewi for i in range(n)])
    ids = np.zeros((nq, k), dtype=np.int64)
    scores = np.zeros((nq, k), dtype=np.float32)
    for qi in range(nq):
        res = index.search(queries[qi], k=k, eta=0.3, entropy_pref=0.5)
        ids[qi] = [int(r[0][3:]) for r in res]
        scores[qi] = [r[1] for r in res]
    # the restatement agrees bit for bit
    cols = {key: sig[j] for j, key in enumerate(oscorer.SIGNAL_KEYS)}
    omed, omad = oscorer.robust_fit(cols)
    assert np.array_equal(oscorer.score_rows(cols, omed, omad), dewi), "quickstart: scorer restatement differs"
    stored = np.asarray(index._backend._embeddings)
    ent = (sig[0].astype(np.float64) + sig[2].astype(np.float64)) * 0.5
    for qi in range(nq):
        oi, osc = osearch.exact_search(stored, dewi, ent, queries[qi], k, 0.3, 0.5, True)
        assert np.array_equal(oi, ids[qi]) and np.array_equal(osc, scores[qi]), "quickstart: search restatement differs"
    np.savez_compressed(GOLD / "quickstart_c1.npz", seed=np.int64(101), n=np.int64(n), d=np.int64(d), k=np.int64(k),
                        emb_sha256=np.array(hashlib.sha256(emb.tobytes()).hexdigest()), signals=sig, queries=queries,
                        dewi=dewi, med=np.array([omed[key] for key in oscorer.SIGNAL_KEYS]),
                        mad=np.array([omad[key] for key in oscorer.SIGNAL_KEYS]), ref_idx=ids, ref_scores=scores)
    return {"case": "quickstart_c1", "n": n, "d": d, "k": k, "nq": nq}


def main() -> None:
    GOLD.mkdir(parents=True, exist_ok=True)
    grid_full = [(e, p) for e in (0.0, 0.25, 0.5, 1.0) for p in (-1.0, 0.0, 0.5, 1.0)]
    manifest = {"reference_src": str(REF_SRC), "numpy": np.__version__, "cases": []}
    m = manifest["cases"]
    m.append(search_case("cos_n100_d128", 100, 128, 5, 10, "cosine", "profile", 42, grid_full))
    m.append(search_case("cos_n2000_d64", 2000, 64, 8, 10, "cosine", "readme", 7, [(0.3, 0.5), (0.25, 0.0), (0.0, 0.0)]))
    m.append(search_case("cos_n3000_d768", 3000, 768, 4, 10, "cosine", "readme", 11, [(0.3, 0.5)]))
    m.append(search_case("cos_n15_d16_k5", 15, 16, 3, 5, "cosine", "profile", 3, [(0.5, 0.0), (0.3, -1.0)]))
    m.append(search_case("cos_n12_d8_k10", 12, 8, 3, 10, "cosine", "profile", 5, [(0.25, 0.0)]))  # 2k > N
    m.append(search_case("l2_n500_d32", 500, 32, 4, 7, "l2", "profile", 9, [(0.5, 0.0), (0.3, 0.5)]))
    m.append(scorer_case("readme_n1001", 1001, "readme", 21, (1.0, 1.0, 1.0, 1.0, 1.0), 3.0))
    m.append(scorer_case("readme_n1000", 1000, "readme", 22, (0.6, 0.2, 1.0, 0.2, 0.1), 3.0))
    m.append(scorer_case("profile_n4096", 4096, "profile", 23, (1.0, 0.5, 2.0, 1.0, 0.25), 1.5))
    m.append(scorer_degenerate())
    m.append(redundancy_case("t37_i53_d512", 37, 53, 512, 31))
    m.append(redundancy_case("t64_i64_d64", 64, 64, 64, 32))
    m.append(extras_case())
    m.append(saved_index_case())
    m.append(quickstart_case())
    (GOLD / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
