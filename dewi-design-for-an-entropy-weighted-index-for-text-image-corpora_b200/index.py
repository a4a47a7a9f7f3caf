"""`DewiIndex` -- the public search facade, API-compatible with the reference's
(src/dewi/index.py:22-166), delegating to the B200 backend.

Constructor arguments, defaults (`rerank_eta=0.25`, `entropy_pref=0.0`), the 1-D query contract,
lazy build, accessors and the on-disk layout (`config.json`, `meta.json`, `ann_index/`) follow the
reference.  The only backend here is `CudaIndex`: the reference's silent fallback to `ExactIndex`
(index.py:58-60) is deliberately absent -- without a B200 the constructor raises ImportError.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from .backends import BaseIndex, CudaIndex, IndexBackend
from .types import Payload


class DewiIndex(BaseIndex):
    def __init__(
        self,
        dim: int,
        space: str = "cosine",
        backend: Union[str, IndexBackend] = "cuda",
        ef: int = 200,
        M: int = 32,
        use_ann: bool = True,
        ef_query: int = 200,
        rerank_eta: float = 0.25,
        entropy_pref: float = 0.0,
        **kwargs: Any,
    ):
        super().__init__(dim, space)
        self._meta: Dict[str, Dict[str, Any]] = {}
        self.ef_query = ef_query
        self.rerank_eta = float(rerank_eta)
        self.entropy_pref = float(entropy_pref)
        self._built = False
        self._use_ann = bool(use_ann)
        if isinstance(backend, str):
            try:
                backend = IndexBackend.from_str(backend)
            except KeyError:
                backend = IndexBackend.CUDA  # the reference maps unknown names to its exact backend (index.py:44-48)
        self.backend = backend
        # ef / M / ef_query are ANN knobs with no meaning for an exact sweep; accepted and ignored.
        self._backend: BaseIndex = CudaIndex(dim, space, **kwargs)

    def add(self, doc_id: str, embedding: np.ndarray, payload: Payload, meta: Optional[Dict[str, Any]] = None) -> None:
        if meta is not None:
            self._meta[doc_id] = meta
        self._backend.add(doc_id, np.asarray(embedding, dtype=np.float32), payload)
        self._built = False

    def add_batch(self, doc_ids, embeddings, payloads=None, payload_columns=None, normalized: bool = False) -> None:
        """Bulk ingest (extension; the reference only has per-document `add`)."""
        self._backend.add_batch(doc_ids, embeddings, payloads, payload_columns, normalized)
        self._built = False

    def build(self) -> None:
        self._backend.build()
        self._built = True

    def search(self, query: np.ndarray, k: int = 10, eta: Optional[float] = None,
               entropy_pref: Optional[float] = None) -> List[Tuple[str, float, Payload]]:
        if not self._built:
            self.build()
        if eta is None:
            eta = self.rerank_eta
        if entropy_pref is None:
            entropy_pref = self.entropy_pref
        q = np.asarray(query, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"Expected query shape ({self.dim},), got {q.shape}")
        return self._backend.search(q, k, eta, entropy_pref)

    def search_batch(self, queries, k: int = 10, eta: Optional[float] = None, entropy_pref: Optional[float] = None,
                     flags: int = 0):
        """`[B, dim]` queries -> `(row_ids [B, k], scores [B, k])` (extension, see CudaIndex.search_batch)."""
        if not self._built:
            self.build()
        eta = self.rerank_eta if eta is None else eta
        entropy_pref = self.entropy_pref if entropy_pref is None else entropy_pref
        return self._backend.search_batch(queries, k, eta, entropy_pref, flags)

    def refresh_payloads(self) -> None:
        self._backend.refresh_payloads()

    def __len__(self) -> int:
        return len(self._backend._doc_ids)

    def get_payload(self, doc_id: str) -> Optional[Payload]:
        return self._backend._payloads.get(doc_id)

    def get_embedding(self, doc_id: str) -> Optional[np.ndarray]:
        try:
            row = self._backend._doc_ids.index(doc_id)
        except ValueError:
            return None
        if row < getattr(self._backend, "_n_device", 0):
            return self._backend.get_row(row)
        store = getattr(self._backend, "_embeddings", None)
        return None if store is None else store[row]

    def get_metadata(self, doc_id: str) -> Optional[Dict[str, Any]]:
        return self._meta.get(doc_id)

    def save(self, path: Union[str, Path]) -> None:
        p = Path(path)
        p.mkdir(parents=True, exist_ok=True)
        self._backend.save(p / "ann_index")
        cfg = {
            "dim": self.dim, "space": self.space, "use_ann": self._use_ann, "ef_query": self.ef_query,
            "rerank_eta": self.rerank_eta, "entropy_pref": self.entropy_pref, "built": self._built,
            "backend_type": type(self._backend).__name__,
        }
        (p / "config.json").write_text(json.dumps(cfg), encoding="utf-8")
        if self._meta:
            (p / "meta.json").write_text(json.dumps(self._meta), encoding="utf-8")

    @classmethod
    def load(cls, path: Union[str, Path], **kwargs: Any) -> "DewiIndex":
        """Loads directories saved by this class or by the reference's `DewiIndex.save` with an
        `ExactIndex` backend (index.py:121-166); either way the rows land in a `CudaIndex`."""
        p = Path(path)
        cfg = json.loads((p / "config.json").read_text(encoding="utf-8"))
        inst = cls.__new__(cls)
        BaseIndex.__init__(inst, cfg["dim"], cfg["space"])
        inst._meta = {}
        inst.ef_query = cfg.get("ef_query", 200)
        inst.rerank_eta = float(cfg.get("rerank_eta", 0.25))
        inst.entropy_pref = float(cfg.get("entropy_pref", 0.0))
        inst._use_ann = bool(cfg.get("use_ann", True))
        inst.backend = IndexBackend.CUDA
        inst._backend = CudaIndex.load(p / "ann_index", **kwargs)
        inst._built = bool(cfg.get("built", False)) and inst._backend._is_trained
        meta_path = p / "meta.json"
        if meta_path.exists():
            inst._meta = json.loads(meta_path.read_text(encoding="utf-8"))
        return inst


__all__ = ["DewiIndex", "BaseIndex", "CudaIndex", "IndexBackend", "Payload"]
