"""fit_stats / score parity: medians and MADs bit-exact, scores within 1e-6 relative (north star)."""

import numpy as np
import pytest

import dewi_b200
from oracle import scorer as oscorer

from _util import GOLD, SIGNAL_FIELDS, synth_payload_columns

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-6


@pytest.mark.parametrize("name", ["readme_n1001", "readme_n1000", "profile_n4096"])
def test_golden_scorer(name):
    g = np.load(GOLD / f"scorer_{name}.npz")
    w = g["weights"]
    s = dewi_b200.DewiScorer(dewi_b200.Weights(*w[:5]), delta=float(w[5]))
    s.fit_stats_columns(g["signals"].astype(np.float32))
    assert [s.stats.medians[k] for k in SIGNAL_FIELDS] == g["med"].tolist()
    assert [s.stats.mads[k] for k in SIGNAL_FIELDS] == g["mad"].tolist()
    for cond, key in ((False, "score"), (True, "score_conditional")):
        out = s.score_batch(g["signals"].astype(np.float32), conditional=cond, out_dtype="float64").cpu().numpy()
        np.testing.assert_allclose(out, g[key], rtol=SCORE_RTOL, atol=0)
        out32 = s.score_batch(g["signals"].astype(np.float32), conditional=cond).cpu().numpy()
        np.testing.assert_allclose(out32, g[key], rtol=SCORE_RTOL, atol=0)


def test_per_row_api_and_one_row_fit():
    """tests/test_scorer_weights.py: fit on ONE row (MAD = 0 -> 1e-8), extra key fitted, floats out."""
    g = np.load(GOLD / "scorer_onerow.npz")
    sig = dict(zip([str(k) for k in g["keys"]], g["row"].tolist()))
    w = g["weights"]
    s = dewi_b200.DewiScorer(weights=dewi_b200.Weights(*w[:5]))
    assert not s.is_fitted()
    with pytest.raises(AssertionError):
        s.score(sig)
    s.fit_stats([sig])
    assert s.is_fitted() and set(s.stats.medians) == set(sig)
    assert all(s.stats.mads[k] == 1e-8 for k in sig)
    a, b = s.score(sig), s.score_conditional(sig)
    assert isinstance(a, float) and isinstance(b, float)
    assert a == pytest.approx(float(g["score"]), rel=SCORE_RTOL)
    assert b == pytest.approx(float(g["score_conditional"]), rel=SCORE_RTOL)


def test_signals_rows_fit_like_the_readme():
    rng = np.random.RandomState(4)
    pay = synth_payload_columns(rng, 257, "readme")
    rows = [dewi_b200.Signals(**{f: float(pay[i, j + 1]) for j, f in enumerate(SIGNAL_FIELDS)}) for i in range(257)]
    s = dewi_b200.DewiScorer()
    s.fit_stats(rows)
    o = oscorer.OracleScorer()
    o.fit_stats([dict(r.items()) for r in rows])
    assert s.stats.medians == o.med and s.stats.mads == o.mad
    for r in rows[:5]:
        assert s.score(r) == pytest.approx(o.score(r), rel=SCORE_RTOL)
        assert dewi_b200.Payload(**r.__dict__, dewi=s.score(r)).ht_mean == r.ht_mean  # README.md:109


@pytest.mark.parametrize("n", [2, 3, 1 << 20, (1 << 22) + 5])
@pytest.mark.parametrize("style", ["readme", "profile"])
def test_differential_fit_and_score(n, style):
    """Odd/even n, heavy duplicates, negative values and constants stress the radix selection."""
    rng = np.random.RandomState(n % 1000 + len(style))
    pay = synth_payload_columns(rng, n, style)
    sig = pay[:, 1:].T.astype(np.float32).copy()
    sig[4] = np.round(sig[4] * 8) / 8          # massive ties
    sig[5] -= 0.5                               # negative values
    if n > 3:
        sig[6, : n // 2 + 1] = 0.125            # median sits inside a constant run -> MAD = 0
    cols = {k: sig[i] for i, k in enumerate(SIGNAL_FIELDS)}
    med, mad = oscorer.robust_fit(cols)
    s = dewi_b200.DewiScorer(dewi_b200.Weights(0.6, 0.2, 1.0, 0.2, 0.1), delta=2.0)
    s.fit_stats_columns(sig)
    assert s.stats.medians == med and s.stats.mads == mad
    ref = oscorer.score_rows(cols, med, mad, (0.6, 0.2, 1.0, 0.2, 0.1, 2.0))
    out = s.score_batch(sig, out_dtype="float64").cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=SCORE_RTOL, atol=0)
    assert out.min() >= 1 / (1 + np.exp(2.0)) - 1e-12 and out.max() <= 1 / (1 + np.exp(-2.0)) + 1e-12


def test_nan_signals_and_extreme_deltas_follow_numpy():
    """np.clip and the sigmoid propagate NaN (scorer.py:62,74); CUDA's fmin / fmax would not.  Any delta is accepted,
    as in the reference: exp overflows to inf (score 0) / underflows (score 1) instead of being rejected."""
    rng = np.random.RandomState(8)
    pay = synth_payload_columns(rng, 4096, "readme")
    sig = pay[:, 1:].T.astype(np.float32).copy()
    cols = {k: sig[i] for i, k in enumerate(SIGNAL_FIELDS)}
    med, mad = oscorer.robust_fit(cols)
    bad = sig.copy()
    bad[2, 7] = np.nan
    bad[5, 100] = np.nan
    bad[0, 200] = np.inf
    bad[0, 201] = -np.inf
    bcols = {k: bad[i] for i, k in enumerate(SIGNAL_FIELDS)}
    for delta in (3.0, 705.0, 800.0, 5000.0):
        s = dewi_b200.DewiScorer(dewi_b200.Weights(), delta=delta)
        s.stats = dewi_b200.RobustStats(medians=dict(med), mads=dict(mad))
        with np.errstate(over="ignore", invalid="ignore"):
            ref = oscorer.score_rows(bcols, med, mad, (1, 1, 1, 1, 1, delta))
        for dt in ("float64", "float32"):
            out = s.score_batch(bad, out_dtype=dt).cpu().numpy().astype(np.float64)
            assert np.isnan(out[7]) and np.isnan(out[100]) and np.array_equal(np.isnan(out), np.isnan(ref))
            ok = ~np.isnan(ref)
            want = ref[ok] if dt == "float64" else ref[ok].astype(np.float32).astype(np.float64)
            np.testing.assert_allclose(out[ok], want, rtol=SCORE_RTOL, atol=1e-300 if dt == "float64" else 1e-45)
    w = dewi_b200.local_weights_from_surprisal(np.array([0.5, np.nan, 2.0, 1.0, 3.0], np.float32))
    assert np.isnan(w[1]) and not np.isnan(w[[0, 2, 3, 4]]).any()


@pytest.mark.parametrize("n", [(1 << 22) - 1, 5_000_002])
def test_adversarial_columns_keep_fit_stats_exact(n):
    """Columns built to defeat the sampled window of the large-n selection (scorer.cu): sorted and periodic data (a
    strided sample is then biased), a constant column, two values, a heavy tail, values one ulp apart, denormals and
    infinities.  Whatever path the kernel takes (window hit, window miss -> plain radix passes) the medians / MADs must
    equal np.median's bit for bit (scorer.py:18-26)."""
    rng = np.random.RandomState(n % 977)
    base = rng.standard_normal(n).astype(np.float32)
    one = np.float32(1.0)
    cols = [
        np.sort(base),                                                     # ascending
        (np.arange(n) % 4096).astype(np.float32),                          # sawtooth, period = a power of two
        np.full(n, 0.375, np.float32),                                     # constant: MAD = 0 -> 1e-8
        np.where(rng.rand(n) < 0.5, np.float32(-2.0), np.float32(7.0)),    # two values
        rng.standard_cauchy(n).astype(np.float32),                         # heavy tail
        np.where(rng.rand(n) < 0.5, one, np.nextafter(one, np.float32(2.0), dtype=np.float32)),   # 1 or 1 + one ulp
        base * np.float32(1e-41),                                          # denormals around zero
    ]
    cols[4][:3] = [np.inf, -np.inf, np.inf]
    sig = np.stack(cols).astype(np.float32)
    named = {k: sig[i] for i, k in enumerate(SIGNAL_FIELDS)}
    med, mad = oscorer.robust_fit(named)
    s = dewi_b200.DewiScorer()
    s.fit_stats_columns(sig)
    assert s.stats.medians == med, (s.stats.medians, med)
    assert s.stats.mads == mad, (s.stats.mads, mad)


def test_more_columns_than_one_library_call_fits():
    """`RobustStats.fit` fits every key a row has (scorer.py:19-25); one dewi_fit_stats call takes 32 columns, so the
    wrapper splits wider inputs -- 40 columns here, medians / MADs bit-equal to numpy's."""
    rng = np.random.RandomState(17)
    n, f = 3001, 40
    cols = {f"s{j:02d}": (rng.standard_normal(n) * (j + 1)).astype(np.float32) for j in range(f)}
    got = dewi_b200.RobustStats.fit_columns(cols)
    med, mad = oscorer.robust_fit(cols)
    assert got.medians == med and got.mads == mad
