"""Oracle: robust median/MAD statistics and the DEWI score.  TEST INFRASTRUCTURE ONLY.

Restates `RobustStats.fit/z` and `DewiScorer._components/score/score_conditional`
of the reference (src/dewi/scorer.py:18-31, 49-89).  `score_rows` is the
vectorised float64 form of the per-row Python arithmetic; SURVEY.md section 7 item 6
records that it is bit-identical to the reference on fp32-representable inputs
(re-checked by oracle/make_golden.py).
"""

from __future__ import annotations

from typing import Dict, List, Mapping, Sequence

import numpy as np

SIGNAL_KEYS = ("ht_mean", "ht_q90", "hi_mean", "hi_q90", "I_hat", "redundancy", "noise")


def robust_fit(columns: Mapping[str, np.ndarray]) -> tuple[Dict[str, float], Dict[str, float]]:
    """scorer.py:18-26 on column arrays.

    `arr[k] = np.asarray([...], dtype=np.float32)`; `med = float(np.median(v))`;
    `mad = float(np.median(np.abs(v - med))) or 1e-8`.  The subtraction of the Python-float
    median from the float32 array stays float32 (weak scalar), so MAD is a float32 median.
    """
    arr = {k: np.asarray(v, dtype=np.float32) for k, v in columns.items()}
    med = {k: float(np.median(v)) for k, v in arr.items()}
    mad = {k: float(np.median(np.abs(v - med[k]))) or 1e-8 for k, v in arr.items()}
    return med, mad


def robust_fit_rows(rows: Sequence[Mapping[str, float]]):
    """scorer.py:19-21: keys come from rows[0]; values are pulled per row."""
    keys = rows[0].keys()
    return robust_fit({k: [r[k] for r in rows] for k in keys})


def score_rows(
    columns: Mapping[str, np.ndarray],
    med: Mapping[str, float],
    mad: Mapping[str, float],
    weights: Sequence[float] = (1.0, 1.0, 1.0, 1.0, 1.0, 3.0),
    conditional: bool = False,
) -> np.ndarray:
    """scorer.py:28-31, 49-89 vectorised in float64.  weights = (a_t, a_i, a_m, a_r, a_n, delta)."""
    a_t, a_i, a_m, a_r, a_n, delta = (float(w) for w in weights)

    def z(name):
        v = np.asarray(columns[name], dtype=np.float64)
        return (v - med[name]) / (1.4826 * mad[name])  # scorer.py:31

    Ht = 0.5 * (z("ht_mean") + z("ht_q90"))  # scorer.py:53
    Hi = 0.5 * (z("hi_mean") + z("hi_q90"))  # scorer.py:54
    I = z("I_hat")
    R = z("redundancy")
    N = z("noise")
    if not conditional:
        U = a_t * Ht + a_i * Hi - a_m * I - a_r * R - a_n * N  # scorer.py:67-73
    else:
        U = a_t * (Ht - I) + a_i * (Hi - I) - a_r * R - a_n * N  # scorer.py:80-87
    U = np.clip(U, -delta, delta)  # scorer.py:74
    return 1.0 / (1.0 + np.exp(-U))  # scorer.py:62


class OracleScorer:
    """Per-row form with the reference's method names (scorer.py:34-89)."""

    def __init__(self, weights: Sequence[float] = (1.0, 1.0, 1.0, 1.0, 1.0), delta: float = 3.0):
        self.w = tuple(float(x) for x in weights[:5]) + (float(delta),)
        self.med = None
        self.mad = None

    def fit_stats(self, rows: List[Mapping[str, float]]) -> None:
        self.med, self.mad = robust_fit_rows(rows)

    def _z(self, name, val):
        return float((val - self.med[name]) / (1.4826 * self.mad[name]))

    def _components(self, sig):
        assert self.med is not None, "Call fit_stats() before scoring."
        return {
            "Ht": 0.5 * (self._z("ht_mean", sig["ht_mean"]) + self._z("ht_q90", sig["ht_q90"])),
            "Hi": 0.5 * (self._z("hi_mean", sig["hi_mean"]) + self._z("hi_q90", sig["hi_q90"])),
            "I": self._z("I_hat", sig["I_hat"]),
            "R": self._z("redundancy", sig["redundancy"]),
            "N": self._z("noise", sig["noise"]),
        }

    def score(self, sig) -> float:
        c = self._components(sig)
        a_t, a_i, a_m, a_r, a_n, delta = self.w
        U = a_t * c["Ht"] + a_i * c["Hi"] - a_m * c["I"] - a_r * c["R"] - a_n * c["N"]
        U = float(np.clip(U, -delta, delta))
        return float(1.0 / (1.0 + np.exp(-U)))

    def score_conditional(self, sig) -> float:
        c = self._components(sig)
        a_t, a_i, a_m, a_r, a_n, delta = self.w
        U = a_t * (c["Ht"] - c["I"]) + a_i * (c["Hi"] - c["I"]) - a_r * c["R"] - a_n * c["N"]
        U = float(np.clip(U, -delta, delta))
        return float(1.0 / (1.0 + np.exp(-U)))
