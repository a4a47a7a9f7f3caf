"""Oracle: redundancy similarity pass.  TEST INFRASTRUCTURE ONLY.

`cross_modal_similarity` restates the only arithmetic of
`RedundancyEstimator.compute_cross_modal_similarity`
(src/dewi/signals/redundancy.py:36-38): `F.normalize(x, p=2, dim=1)` on both
feature matrices (torch eps 1e-12: `x / max(||x||_2, eps)`), then `T @ I.T`, float32.
The CLIP forward that produces the features is out of scope.

PARITY UNPINNED beyond that product: the reference defines no threshold, no
per-row reduction and no pair list.  `join_rowstats` is
*this repository's* definition of the thresholded join, restated on the CPU so
the CUDA join can be checked against something.
"""

from __future__ import annotations

import numpy as np


def l2_normalize_rows(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(x, p=2, dim=1, eps) in float32 (redundancy.py:36-37)."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt(np.sum(x * x, axis=1, keepdims=True, dtype=np.float32))
    return x / np.maximum(n, np.float32(eps))


def cross_modal_similarity(tfeat: np.ndarray, ifeat: np.ndarray) -> np.ndarray:
    """redundancy.py:36-38: normalised text x image similarity, dense [T, I] float32."""
    return l2_normalize_rows(tfeat) @ l2_normalize_rows(ifeat).T


def join_rowstats(a: np.ndarray, b: np.ndarray, tau: float, self_join: bool = False, block: int = 4096):
    """Thresholded join of normalised rows of `a` against normalised rows of `b`.

    Returns (max_sim[M] f32, argmax[M] i64, count_ge_tau[M] i64, pairs[(i, j, sim)]).
    self_join: `b` is `a`; the diagonal is excluded from the row statistics and only
    pairs with j > i are emitted.
    """
    an = l2_normalize_rows(a)
    bn = an if self_join else l2_normalize_rows(b)
    m, n = an.shape[0], bn.shape[0]
    max_sim = np.full(m, -np.inf, dtype=np.float32)
    argmax = np.full(m, -1, dtype=np.int64)
    count = np.zeros(m, dtype=np.int64)
    pairs = []
    for i0 in range(0, m, block):
        s = an[i0 : i0 + block] @ bn.T
        rows = np.arange(i0, min(i0 + block, m))
        if self_join:
            s[rows - i0, rows] = -np.inf
        j = np.argmax(s, axis=1)
        max_sim[rows] = s[rows - i0, j]
        argmax[rows] = j
        hit = s >= np.float32(tau)
        count[rows] = hit.sum(axis=1)
        ii, jj = np.nonzero(hit)
        if self_join:
            keep = jj > (ii + i0)
            ii, jj = ii[keep], jj[keep]
        pairs.extend(zip((ii + i0).tolist(), jj.tolist(), s[ii, jj].tolist()))
    return max_sim, argmax, count, pairs
