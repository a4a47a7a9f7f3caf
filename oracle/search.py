"""Oracle: exact DEWI-re-ranked search.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates `ExactIndex.add/build/search` of the reference
(src/dewi/backends.py:394-481).  The statements below follow the reference
line by line so that, on the same inputs and the same numpy/BLAS, outputs are
bit-identical (checked by oracle/make_golden.py against the real class).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def normalize_rows_like_add(emb: np.ndarray) -> np.ndarray:
    """Row-normalise the way `ExactIndex.add` does, one row at a time.

    backends.py:403-405: `emb = embedding.astype(np.float32); emb = emb / np.linalg.norm(emb)`.
    Per-row calls keep the BLAS `sdot` reduction order identical to the reference's.
    """
    emb = np.asarray(emb, dtype=np.float32)
    out = np.empty_like(emb)
    for i in range(emb.shape[0]):
        out[i] = emb[i] / np.linalg.norm(emb[i])
    return out


def exact_search(
    embeddings: np.ndarray,
    dewi: np.ndarray,
    entropy: np.ndarray,
    query: np.ndarray,
    k: int = 10,
    eta: float = 0.5,
    entropy_pref: float = 0.0,
    normalize: bool = True,
) -> Tuple[np.ndarray, np.ndarray]:
    """One query against a built corpus.  Returns (row_indices[k], scores[k]) sorted desc.

    embeddings : [N, D] float32, already normalised when `normalize` (backends.py:408-411)
    dewi       : [N] float64/32 -- `payload.dewi` per row            (backends.py:457)
    entropy    : [N] float64    -- `(ht_mean + hi_mean) * 0.5` evaluated in Python floats
                                   (backends.py:458) *before* the store into a float32 array
    """
    # backends.py:420-424
    query = np.asarray(query, dtype=np.float32)
    if normalize:
        query_norm = np.linalg.norm(query)
        if query_norm > 0:
            query = query / query_norm
    # backends.py:427-428
    if query.ndim == 1:
        query = query.reshape(1, -1)
    # backends.py:431-436
    if normalize:
        scores = np.dot(embeddings, query.T).flatten()
    else:
        scores = -np.sum((embeddings - query) ** 2, axis=1)
    # backends.py:439-441
    candidate_count = min(2 * k, len(scores))
    if candidate_count <= 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
    # backends.py:444-447
    top_indices = np.argpartition(scores, -candidate_count)[-candidate_count:]
    candidate_scores = scores[top_indices]
    # backends.py:450-458 (payload gather: Python floats stored into float32 arrays)
    dewi_scores = np.zeros(candidate_count, dtype=np.float32)
    entropies = np.zeros(candidate_count, dtype=np.float32)
    for i, idx in enumerate(top_indices):
        dewi_scores[i] = dewi[idx]
        entropies[i] = entropy[idx]
    # backends.py:461-465  (eta / entropy_pref must be Python floats: weak scalars keep float32)
    adjusted_scores = (1 - eta) * candidate_scores + eta * dewi_scores
    if entropy_pref != 0:
        adjusted_scores += entropy_pref * entropies
    # backends.py:468-471 (raises ValueError when k > candidate_count, i.e. k > N)
    top_k_indices = np.argpartition(adjusted_scores, -k)[-k:]
    sorted_indices = top_k_indices[np.argsort(-adjusted_scores[top_k_indices])]
    return top_indices[sorted_indices].astype(np.int64), adjusted_scores[sorted_indices]


def exact_search_batch(embeddings, dewi, entropy, queries, k=10, eta=0.5, entropy_pref=0.0, normalize=True):
    """The reference has no batch API (index.py:91-92): a batch is B sequential calls."""
    ids, scs = [], []
    for q in np.asarray(queries):
        i, s = exact_search(embeddings, dewi, entropy, q, k, eta, entropy_pref, normalize)
        ids.append(i)
        scs.append(s)
    return np.stack(ids), np.stack(scs)


def full_scope_search(embeddings, dewi, entropy, query, k=10, eta=0.5, entropy_pref=0.0, normalize=True):
    """`rerank_scope="full"` -- NOT a reference code path (parity unpinned by definition): the
    reference's own blend statements (backends.py:461-465, float32 arrays, weak Python-float scalars) applied to EVERY
    row instead of to the top-2k by similarity, then its final select / sort (backends.py:468-471)."""
    query = np.asarray(query, dtype=np.float32)
    if normalize:
        query_norm = np.linalg.norm(query)
        if query_norm > 0:
            query = query / query_norm
    query = query.reshape(1, -1)
    if normalize:
        scores = np.dot(embeddings, query.T).flatten()
    else:
        scores = -np.sum((embeddings - query) ** 2, axis=1)
    dewi_scores = np.asarray(dewi).astype(np.float32)
    entropies = np.asarray(entropy).astype(np.float32)
    adjusted_scores = (1 - eta) * scores + eta * dewi_scores
    if entropy_pref != 0:
        adjusted_scores += entropy_pref * entropies
    top_k_indices = np.argpartition(adjusted_scores, -k)[-k:]
    sorted_indices = top_k_indices[np.argsort(-adjusted_scores[top_k_indices])]
    return sorted_indices.astype(np.int64), adjusted_scores[sorted_indices]


def candidate_sims(embeddings, query, k, normalize=True):
    """Raw similarities and the 2k candidate set of backends.py:431-447 (for tie-window checks)."""
    query = np.asarray(query, dtype=np.float32)
    if normalize:
        n = np.linalg.norm(query)
        if n > 0:
            query = query / n
    q = query.reshape(1, -1)
    scores = np.dot(embeddings, q.T).flatten() if normalize else -np.sum((embeddings - q) ** 2, axis=1)
    c = min(2 * k, len(scores))
    top = np.argpartition(scores, -c)[-c:]
    return scores, top


class OracleExactIndex:
    """Object form with the reference's add/build/search surface (backends.py:386-481).

    Holds parallel arrays instead of Payload objects so it has no dependency on
    the product package; `payload` arguments only need `.dewi/.ht_mean/.hi_mean`.
    """

    def __init__(self, dim: int, space: str = "cosine"):
        self.dim = dim
        self.space = space
        self._normalize = space == "cosine"
        self._doc_ids: List[str] = []
        self._payloads = {}
        self._embeddings = []
        self._is_trained = False

    def add(self, doc_id: str, embedding: np.ndarray, payload) -> None:
        # backends.py:394-406
        if embedding.shape != (self.dim,):
            raise ValueError(f"Expected embedding of shape {(self.dim,)}, got {embedding.shape}")
        self._doc_ids.append(doc_id)
        self._payloads[doc_id] = payload
        emb = embedding.astype(np.float32)
        if self._normalize:
            emb = emb / np.linalg.norm(emb)
        self._embeddings.append(emb)

    def bulk_assign(self, doc_ids: Sequence[str], rows: np.ndarray, payloads: Sequence) -> None:
        """Bypass the per-doc loop the way SURVEY.md section 8c allows: rows are stored verbatim
        (they must already be normalised / bf16-representable); `search` is unchanged."""
        self._doc_ids = list(doc_ids)
        self._payloads = dict(zip(self._doc_ids, payloads))
        self._embeddings = np.ascontiguousarray(rows, dtype=np.float32)
        self._is_trained = True

    def build(self) -> None:
        # backends.py:408-412
        if isinstance(self._embeddings, list):
            if not self._embeddings:
                raise ValueError("No embeddings to build index from")
            self._embeddings = np.stack(self._embeddings)
        self._is_trained = True

    def search(self, query, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0):
        # The payload columns are re-read at search time (payloads are shared by
        # reference, backends.py:454-458), duplicates resolve through the id->payload dict.
        pl = [self._payloads[d] for d in self._doc_ids]
        dewi = np.array([p.dewi for p in pl], dtype=np.float64)
        ent = np.array([(p.ht_mean + p.hi_mean) * 0.5 for p in pl], dtype=np.float64)
        idx, sc = exact_search(self._embeddings, dewi, ent, query, k, eta, entropy_pref, self._normalize)
        return [(self._doc_ids[i], float(s), self._payloads[self._doc_ids[i]]) for i, s in zip(idx, sc)]
