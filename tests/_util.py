"""Shared test helpers: synthetic inputs, golden loaders, tie-aware comparison."""

from __future__ import annotations

import hashlib
from pathlib import Path
from typing import Tuple

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"
PAYLOAD_FIELDS = ("dewi", "ht_mean", "ht_q90", "hi_mean", "hi_q90", "I_hat", "redundancy", "noise")
SIGNAL_FIELDS = PAYLOAD_FIELDS[1:]

# north-star gates (BASELINE.json): fp32 scores within 1e-5 relative; ids equal outside exact ties.
SCORE_RTOL = 1e-5
# two reference scores closer than this are a tie for ordering purposes (numpy's sgemv summation
# order is itself unspecified; SURVEY.md section 8d "parity gates")
TIE_TOL = 2e-6


def synth_payload_columns(rng: np.random.RandomState, n: int, style: str) -> np.ndarray:
    """[n, 8] float64 with float32-representable values; same generator as oracle/make_golden.py."""
    if style == "profile":
        cols = [
            np.clip(rng.beta(2, 2, n), 0, 1), rng.gamma(2, 0.5, n), rng.gamma(2, 0.5, n) * 1.5,
            rng.gamma(2, 0.3, n), rng.gamma(2, 0.3, n) * 1.5, rng.beta(2, 2, n), rng.beta(1, 5, n), rng.beta(1, 10, n),
        ]
    else:
        cols = [
            rng.uniform(0, 1, n), rng.uniform(0, 10, n), rng.uniform(0, 15, n), rng.uniform(0, 5, n),
            rng.uniform(0, 8, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 0.2, n),
        ]
    return np.stack(cols, axis=1).astype(np.float32).astype(np.float64)


def entropy_column(pay: np.ndarray) -> np.ndarray:
    """`(ht_mean + hi_mean) * 0.5` in float64, as the reference evaluates it (backends.py:458)."""
    return (pay[:, 1] + pay[:, 3]) * 0.5


def load_search_golden(name: str):
    """Fixture written by oracle/make_golden.py; large corpora are regenerated from (seed, sha256)."""
    g = np.load(GOLD / f"search_{name}.npz")
    emb = g["emb"]
    n, d = int(g["n"]), g["queries"].shape[1]
    if emb.shape[0] == 0:
        rng = np.random.RandomState(int(g["seed"]))
        emb = rng.randn(n, d).astype(np.float32)
        if str(g["space"]) == "l2":
            emb *= 0.25
    assert hashlib.sha256(emb.tobytes()).hexdigest() == str(g["emb_sha256"]), "regenerated corpus differs"
    return {
        "emb": emb, "payload": g["payload"], "queries": g["queries"], "grid": [tuple(map(float, r)) for r in g["grid"]],
        "k": int(g["k"]), "space": str(g["space"]), "ref_idx": g["ref_idx"], "ref_scores": g["ref_scores"],
    }


def make_corpus(n: int, d: int, seed: int, style: str = "profile") -> Tuple[np.ndarray, np.ndarray]:
    """Gaussian rows, row-normalised in fp32 (scripts/profile_index.py:55-56), + payload columns."""
    rng = np.random.RandomState(seed)
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    return emb, synth_payload_columns(rng, n, style)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what the device stores in bf16 mode)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def check_topk(ref_idx, ref_sc, got_idx, got_sc, rtol: float = SCORE_RTOL, tie: float = TIE_TOL, what: str = ""):
    """Parity gate for one query: scores position-wise within `rtol`; ids identical except where the
    reference's own scores are tied within `tie` (then any order / member of the tie group is valid)."""
    ref_idx, got_idx = np.asarray(ref_idx), np.asarray(got_idx)
    ref_sc, got_sc = np.asarray(ref_sc, dtype=np.float64), np.asarray(got_sc, dtype=np.float64)
    assert ref_idx.shape == got_idx.shape, f"{what}: result count {got_idx.shape} != {ref_idx.shape}"
    scale = np.maximum(1.0, np.abs(ref_sc))
    err = np.abs(got_sc - ref_sc) / scale
    assert np.all(err <= rtol), f"{what}: score mismatch {err.max():.3e} > {rtol} (ref {ref_sc}, got {got_sc})"
    assert np.all(np.diff(got_sc) <= 0), f"{what}: scores not sorted descending"
    for pos in np.nonzero(ref_idx != got_idx)[0]:
        # a differing id is legitimate only inside a tie group of the reference ranking (including the
        # k-th boundary, where the tied partner may lie outside the returned list)
        near = np.abs(ref_sc - ref_sc[pos]) <= tie * scale[pos]
        tied_ids = set(ref_idx[near].tolist())
        at_boundary = near[-1]
        assert got_idx[pos] in tied_ids or at_boundary, (
            f"{what}: id mismatch at rank {pos}: ref {ref_idx[pos]} got {got_idx[pos]} (ref scores {ref_sc})")


def recall_at_k(ref_idx: np.ndarray, got_idx: np.ndarray) -> float:
    """Mean |ref ∩ got| / k over queries (metrics.py:9-36 with the reference list as the relevant set)."""
    hits = [len(set(r.tolist()) & set(g.tolist())) / len(r) for r, g in zip(ref_idx, got_idx)]
    return float(np.mean(hits))
