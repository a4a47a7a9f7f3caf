"""Neighbours of the scorer that reuse its kernels.

`PayloadRobustStats` mirrors the reference's *second* RobustStats (src/dewi/robust.py:8-32, exported there
as `dewi.RobustStats`): median / MAD of four Payload fields, additive `+1e-8` on the MAD, float32 z.
`local_weights_from_surprisal` mirrors src/dewi/local_weights.py:5-26.  `cluster_pairs` and the two metric
helpers turn the join's pair list into the clusters `metrics.duplicate_rate` / `cluster_coverage`
(src/dewi/metrics.py:173-212) consume.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .scorer import _device_index, _fit_columns, _torch
from .types import Payload


@dataclass(frozen=True)
class PayloadRobustStats:
    """robust.py:13-32.  `fields[name] = (median, mad)` over `ht_mean, hi_mean, redundancy, noise`."""

    fields: Dict[str, Tuple[float, float]]
    KEYS = ("ht_mean", "hi_mean", "redundancy", "noise")

    @classmethod
    def from_payloads(cls, payloads: Sequence[Payload], device: Optional[int] = None) -> "PayloadRobustStats":
        if not payloads:
            raise ValueError("Cannot compute statistics from empty dataset")
        host = np.array([[getattr(p, k) for p in payloads] for k in cls.KEYS], dtype=np.float32)  # robust.py:22
        return cls.from_columns(host, device)

    @classmethod
    def from_columns(cls, matrix, device: Optional[int] = None) -> "PayloadRobustStats":
        """`matrix`: `[4, N]` (KEYS order), numpy or torch, host or device."""
        torch = _torch()
        dev = _device_index(device)
        _native.require_device(dev)
        cols = torch.as_tensor(matrix).to(torch.device("cuda", dev), torch.float32)
        if cols.ndim != 2 or cols.shape[0] != len(cls.KEYS):
            raise ValueError("expected four columns: ht_mean, hi_mean, redundancy, noise")
        if cols.stride(1) != 1:
            cols = cols.contiguous()
        med, mad = _fit_columns(cols, dev, zero_mad_as=0.0)
        return cls(fields={k: (float(m), float(d)) for k, m, d in zip(cls.KEYS, med, mad)})

    def z(self, name: str, val: float) -> float:
        med, mad = self.fields[name]
        # robust.py:8-10,30-32 verbatim (one scalar: not a hot path)
        return float((np.array(val, dtype=np.float32) - med) / (1.4826 * (float(mad) + 1e-8)))


def local_weights_from_surprisal(s, device: Optional[int] = None):
    """Per-token / per-patch surprisals -> positive weights (local_weights.py:5-26).  numpy in -> numpy
    out; CUDA tensor in -> CUDA tensor out.  Same shape as the input."""
    torch = _torch()
    dev = _device_index(device)
    _native.require_device(dev)
    is_np = not hasattr(s, "is_cuda")
    t = torch.as_tensor(np.asarray(s, dtype=np.float32) if is_np else s).to(torch.device("cuda", dev), torch.float32)
    flat = t.contiguous().view(-1)
    out = torch.empty_like(flat)
    lib = _native.load_library()
    with torch.cuda.device(dev):
        rc = lib.dewi_local_weights(ctypes.c_void_p(flat.data_ptr()), flat.numel(), ctypes.c_void_p(out.data_ptr()), dev,
                                    _native.stream_ptr())
    _native.check(rc)
    out = out.view(t.shape)
    return out.cpu().numpy() if is_np else out


def cluster_pairs(pairs_i, pairs_j, n: int, device: Optional[int] = None):
    """Connected components of the pair graph over documents `0..n-1` -> int32 CUDA tensor `labels[n]`,
    `labels[i]` = smallest index in i's cluster."""
    torch = _torch()
    dev = _device_index(device)
    _native.require_device(dev)
    tdev = torch.device("cuda", dev)
    pi = torch.as_tensor(pairs_i, dtype=torch.int64).to(tdev).contiguous()
    pj = torch.as_tensor(pairs_j, dtype=torch.int64).to(tdev).contiguous()
    if pi.shape != pj.shape or pi.ndim != 1:
        raise ValueError("pairs_i and pairs_j must be 1-D and equally long")
    labels = torch.empty(int(n), dtype=torch.int32, device=tdev)
    lib = _native.load_library()
    with torch.cuda.device(dev):
        rc = lib.dewi_cluster_pairs(ctypes.c_void_p(pi.data_ptr()), ctypes.c_void_p(pj.data_ptr()), pi.numel(), int(n),
                                    ctypes.c_void_p(labels.data_ptr()), dev, _native.stream_ptr())
    _native.check(rc)
    return labels


def clusters_from_labels(labels, doc_ids: Optional[Sequence[str]] = None) -> List[List]:
    """Cluster label array -> list of clusters (lists of doc ids, or of row indices)."""
    lab = np.asarray(labels.cpu() if hasattr(labels, "cpu") else labels)
    order = np.argsort(lab, kind="stable")
    bounds = np.flatnonzero(np.diff(lab[order])) + 1
    groups = np.split(order, bounds)
    return [[doc_ids[i] if doc_ids is not None else int(i) for i in g] for g in groups]


def duplicate_rate(clusters: List[Sequence]) -> float:
    """metrics.py:173-192: share of clusters that are not singletons."""
    if not clusters or sum(len(c) for c in clusters) == 0:
        return 0.0
    return 1.0 - sum(1 for c in clusters if len(c) == 1) / len(clusters)


def cluster_coverage(selected: Sequence, clusters: List[Sequence]) -> float:
    """metrics.py:194-212: share of clusters with at least one selected document."""
    if not clusters:
        return 0.0
    chosen = set(selected)
    return sum(1 for c in clusters if any(d in chosen for d in c)) / len(clusters)
