"""The arithmetic of `score_kernel` (csrc/scorer.cu) restated in numpy and held against the oracle
(`DewiScorer.score` / `score_conditional`, scorer.py:49-89): the folded form U = sum_c k_c (v_c - med_c) and
the range-reduced degree-9 polynomial exp stay within 1e-9 of the reference's float64 arithmetic (gate: 1e-6),
including the zero-MAD (1e-8) column, and the float32-output variant within 2e-7."""

import numpy as np
import pytest

from oracle import scorer as oscorer


def exp_small(x):
    """exp via round-to-nearest range reduction + degree-9 Taylor polynomial (scorer.cu: exp_small)."""
    kf = np.rint(x * 1.4426950408889634)
    r = x - kf * 6.93147180369123816490e-01
    r = r - kf * 1.90821492927058770002e-10
    p = np.full_like(r, 1.0 / 362880.0)
    for c in (1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        p = p * r + c
    return p * np.exp2(kf)


def kernel_scores(sig, med, mad, w6, conditional):
    a_t, a_i, a_m, a_r, a_n, delta = w6
    inv = np.array([1.0 / (1.4826 * mad[k]) for k in oscorer.SIGNAL_KEYS])
    m = np.array([med[k] for k in oscorer.SIGNAL_KEYS])
    wk = np.array([0.5 * a_t, 0.5 * a_t, 0.5 * a_i, 0.5 * a_i, -(a_t + a_i) if conditional else -a_m, -a_r, -a_n]) * inv
    u = np.zeros(sig.shape[1])
    for c in range(7):
        u = u + wk[c] * (sig[c].astype(np.float64) - m[c])
    e = exp_small(-np.clip(u, -delta, delta))
    return 1.0 / (1.0 + e), (np.float32(1.0) / (1.0 + e).astype(np.float32)).astype(np.float32)


def test_exp_polynomial_accuracy():
    x = np.linspace(-700.0, 700.0, 2_000_001)
    assert np.max(np.abs(exp_small(x) - np.exp(x)) / np.exp(x)) < 5e-11


@pytest.mark.parametrize("conditional", [False, True])
@pytest.mark.parametrize("w6", [(1.0, 1.0, 1.0, 1.0, 1.0, 3.0), (0.6, 0.2, 1.0, 0.2, 0.1, 2.0), (2.0, 0.0, 0.5, 3.0, 0.0, 30.0)])
def test_folded_score_matches_the_oracle(w6, conditional):
    rng = np.random.RandomState(7)
    n = 200_000
    hi = np.array([10, 15, 5, 8, 1, 1, 0.2])
    sig = (rng.rand(7, n) * hi[:, None]).astype(np.float32)
    sig[5] -= 0.5                       # negative values
    sig[6, : n // 2 + 1] = 0.125        # median inside a constant run -> MAD = 0 -> 1e-8 (scorer.py:24)
    cols = {k: sig[i] for i, k in enumerate(oscorer.SIGNAL_KEYS)}
    med, mad = oscorer.robust_fit(cols)
    assert mad["noise"] == 1e-8
    ref = oscorer.score_rows(cols, med, mad, w6, conditional=conditional)
    s64, s32 = kernel_scores(sig, med, mad, w6, conditional)
    assert np.max(np.abs(s64 - ref) / ref) < 1e-9
    big = ref > 1e-30                   # float32 output cannot hold sigma(-30) to 2e-7 relative once it is denormal
    assert np.max(np.abs(s32[big].astype(np.float64) - ref[big]) / ref[big]) < 2e-7
