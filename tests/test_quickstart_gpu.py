"""BASELINE.json config 1 as ONE flow: the README quick-start (README.md:67-110) at 10K x 768, written against
dewi_b200 exactly as the README writes it against dewi, checked against the fixture the unmodified reference
produced for the same inputs (tests/golden/quickstart_c1.npz, oracle/make_golden.py: quickstart_case)."""

import hashlib

import numpy as np
import pytest

from _util import GOLD, check_topk

pytestmark = pytest.mark.gpu


def test_readme_quickstart_flow_matches_the_reference():
    from dewi_b200 import DewiIndex, DewiScorer, Payload, Signals, Weights  # README.md:67-68 imports, one package

    g = np.load(GOLD / "quickstart_c1.npz")
    n, d, k = int(g["n"]), int(g["d"]), int(g["k"])
    rng = np.random.RandomState(int(g["seed"]))
    embeddings = rng.rand(n, d).astype(np.float32)
    assert hashlib.sha256(embeddings.tobytes()).hexdigest() == str(g["emb_sha256"])
    sig = g["signals"]
    ids = [f"doc{i}" for i in range(n)]

    index = DewiIndex(dim=d, space="cosine")                                    # README.md:79
    rows = []
    for i, doc_id in enumerate(ids):
        signals = Signals(ht_mean=float(sig[0, i]), ht_q90=float(sig[1, i]), hi_mean=float(sig[2, i]), hi_q90=float(sig[3, i]),
                          I_hat=float(sig[4, i]), redundancy=float(sig[5, i]), noise=float(sig[6, i]))
        rows.append(signals)
        payload = Payload(dewi=0.0, **signals.__dict__)                         # README.md:94-97
        index.add(doc_id, embeddings[i], payload)
    scorer = DewiScorer(Weights())                                              # README.md:101-102
    scorer.fit_stats(rows)
    keys = list(Signals.__annotations__)
    assert [scorer.stats.medians[key] for key in keys] == g["med"].tolist()     # medians / MADs bit-equal
    assert [scorer.stats.mads[key] for key in keys] == g["mad"].tolist()
    for doc_id in ids:                                                          # README.md:105-110, verbatim
        payload = index.get_payload(doc_id)
        if payload:
            signals = Signals(**{key: getattr(payload, key) for key in Signals.__annotations__})
            payload.dewi = scorer.score(signals)
    dewi = np.array([index.get_payload(doc_id).dewi for doc_id in ids])
    assert np.max(np.abs(dewi - g["dewi"]) / g["dewi"]) <= 1e-6                 # scorer gate
    index.build()                                                               # README.md:113
    for qi, q in enumerate(g["queries"]):
        res = index.search(q, k=k, eta=0.3, entropy_pref=0.5)
        assert all(isinstance(r[0], str) and isinstance(r[1], float) and isinstance(r[2], Payload) for r in res)
        check_topk(g["ref_idx"][qi], g["ref_scores"][qi], [int(r[0][3:]) for r in res], [r[1] for r in res], what=f"C1 q{qi}")
        assert res[0][2] is index.get_payload(res[0][0])                        # payloads come back by reference
    assert len(index) == n


def test_bulk_form_of_the_quickstart_equals_the_loop():
    """The same flow without per-document Python: add_batch + set_payload_from_signals (mirrored)."""
    from dewi_b200 import DewiIndex, DewiScorer, Weights

    g = np.load(GOLD / "quickstart_c1.npz")
    n, d, k = int(g["n"]), int(g["d"]), int(g["k"])
    embeddings = np.random.RandomState(int(g["seed"])).rand(n, d).astype(np.float32)
    index = DewiIndex(dim=d, space="cosine")
    index.add_batch([f"doc{i}" for i in range(n)], embeddings, payload_columns=np.zeros((n, 8), np.float32))
    scorer = DewiScorer(Weights())
    index._backend.set_payload_from_signals(g["signals"], scorer)
    index.build()
    for qi, q in enumerate(g["queries"]):
        res = index.search(q, k=k, eta=0.3, entropy_pref=0.5)
        check_topk(g["ref_idx"][qi], g["ref_scores"][qi], [int(r[0][3:]) for r in res], [r[1] for r in res], what=f"bulk q{qi}")
        top = res[0][2]
        row = int(res[0][0][3:])
        assert abs(top.dewi - g["dewi"][row]) <= 1e-6 and top.ht_mean == float(g["signals"][0, row])
