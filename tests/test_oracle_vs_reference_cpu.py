"""The oracle against the UNMODIFIED reference, live, on fresh seeds (the committed fixtures in tests/golden/ pin a fixed
set of cases; this re-derives the pin on inputs no fixture holds).  Needs the reference -- baseline/_ref
(scripts/vendor_reference.sh; present in the build container and on the GPU box) or /root/reference/src -- and is
skipped where neither exists.  Bit equality is demanded: same numpy, same BLAS, same statements
(src/dewi/backends.py:394-481, src/dewi/scorer.py:18-89)."""

import importlib
import logging
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import scorer as oscorer
from oracle import search as osearch

ROOT = Path(__file__).resolve().parent.parent
SIGNAL_KEYS = ("ht_mean", "ht_q90", "hi_mean", "hi_q90", "I_hat", "redundancy", "noise")


def _reference():
    for p in (ROOT / "baseline" / "_ref", Path("/root/reference/src")):
        if (p / "dewi" / "backends.py").exists():
            if str(p) not in sys.path:
                sys.path.insert(0, str(p))
            logging.getLogger("dewi.backends").setLevel(logging.ERROR)
            return (importlib.import_module("dewi.index"), importlib.import_module("dewi.types"),
                    importlib.import_module("dewi.scorer"))
    pytest.skip("the reference is not installed here (scripts/vendor_reference.sh)")


@pytest.mark.parametrize("seed", range(6))
def test_search_oracle_equals_the_reference_on_fresh_seeds(seed):
    index_mod, types_mod, _ = _reference()
    rng = np.random.RandomState(9000 + seed)
    n = int(rng.choice([7, 40, 333, 1500]))
    d = int(rng.choice([8, 64, 200]))
    space = "cosine" if seed % 3 else "l2"
    emb = rng.standard_normal((n, d)).astype(np.float32) * (1.0 if space == "cosine" else 0.3)
    pay = [types_mod.Payload(dewi=float(np.float32(rng.uniform())), ht_mean=float(np.float32(rng.uniform(0, 10))),
                             hi_mean=float(np.float32(rng.uniform(0, 5)))) for _ in range(n)]
    ref = index_mod.DewiIndex(dim=d, space=space, use_ann=False)
    ora = osearch.OracleExactIndex(d, space)
    for i in range(n):
        ref.add(f"doc_{i}", emb[i], pay[i])
        ora.add(f"doc_{i}", emb[i], pay[i])
    ref.build()
    ora.build()
    assert np.array_equal(ref._backend._embeddings, ora._embeddings)   # rows normalised the same way (backends.py:403-405)
    dewi = np.array([p.dewi for p in pay])
    ent = np.array([(p.ht_mean + p.hi_mean) * 0.5 for p in pay])
    for _ in range(8):
        q = rng.standard_normal(d).astype(np.float32)
        k = int(rng.randint(1, min(n, 25) + 1))
        eta = float(rng.choice([0.0, 0.25, 0.3, 1.0]))
        pref = float(rng.choice([0.0, 0.5, -0.4]))
        want = ref.search(q, k=k, eta=eta, entropy_pref=pref)
        got = ora.search(q, k=k, eta=eta, entropy_pref=pref)
        assert [w[0] for w in want] == [g[0] for g in got]
        assert [w[1] for w in want] == [g[1] for g in got]                                 # bit-equal Python floats
        ids, sc = osearch.exact_search(ora._embeddings, dewi, ent, q, k, eta, pref, space == "cosine")
        assert [f"doc_{i}" for i in ids] == [w[0] for w in want] and [float(s) for s in sc] == [w[1] for w in want]
    with pytest.raises(ValueError):
        ref.search(q, k=n + 1)
    with pytest.raises(ValueError):
        ora.search(q, k=n + 1)


@pytest.mark.parametrize("seed", range(4))
def test_scorer_oracle_equals_the_reference_on_fresh_seeds(seed):
    _, types_mod, scorer_mod = _reference()
    rng = np.random.RandomState(9100 + seed)
    n = int(rng.choice([1, 2, 17, 400, 1001]))
    rows = [{k: float(np.float32(rng.uniform(0, hi))) for k, hi in zip(SIGNAL_KEYS, (10, 15, 5, 8, 1, 1, 0.2))} for _ in range(n)]
    if n > 4:
        for r in rows[: n // 2 + 1]:
            r["noise"] = 0.125                        # the median sits in a constant run: MAD = 0 -> 1e-8 (scorer.py:24)
    w = [float(x) for x in rng.uniform(0.1, 2.0, 5)]
    delta = float(rng.choice([0.5, 3.0, 40.0]))
    ref = scorer_mod.DewiScorer(types_mod.Weights(*w), delta=delta)
    ora = oscorer.OracleScorer(w, delta=delta)
    ref.fit_stats(rows)
    ora.fit_stats(rows)
    assert ref.stats.medians == ora.med and ref.stats.mads == ora.mad
    cols = {k: np.asarray([r[k] for r in rows], dtype=np.float64) for k in SIGNAL_KEYS}
    bulk = oscorer.score_rows(cols, ora.med, ora.mad, tuple(w) + (delta,))
    bulk_c = oscorer.score_rows(cols, ora.med, ora.mad, tuple(w) + (delta,), conditional=True)
    for i, r in enumerate(rows[:50]):
        assert ref.score(r) == ora.score(r) and ref.score_conditional(r) == ora.score_conditional(r)
        # the vectorised restatement (what the 100M-row tests use) agrees to the last few ulps of float64
        assert bulk[i] == pytest.approx(ref.score(r), rel=1e-12) and bulk_c[i] == pytest.approx(ref.score_conditional(r), rel=1e-12)
