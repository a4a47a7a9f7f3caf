"""Index backends: the plugin boundary of the reference, plus the B200 backend.

`BaseIndex` / `IndexBackend` mirror the reference's plugin API (src/dewi/backends.py:32-49,54-163):
a backend is a class with `add / build / search / save / load` and the attributes `_doc_ids`,
`_payloads`, `_embeddings`, `_is_trained` that the `DewiIndex` facade reads (src/dewi/index.py:95-116).
`CudaIndex` is the drop-in sibling of `ExactIndex` (backends.py:386-556): same constructor, same
exceptions, same result tuples, same on-disk directory -- with every search step executed by
libdewi_b200.so on a B200.  There is no CPU code path behind it.
"""

from __future__ import annotations

import ctypes
import json
from enum import Enum, auto
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native
from .types import PAYLOAD_FIELDS, Payload


class IndexBackend(Enum):
    """Backend names accepted by `DewiIndex(backend=...)` (backends.py:32-49) + CUDA."""

    HNSW = auto()
    FAISS_IVFFLAT = auto()
    FAISS_HNSW = auto()
    EXACT = auto()
    CUDA = auto()

    @classmethod
    def from_str(cls, name: str) -> "IndexBackend":
        name = name.upper()
        if name in ("AUTO", "B200", "CUDAINDEX"):
            return cls.CUDA
        if name == "EXACTINDEX":
            return cls.EXACT
        return cls[name]  # KeyError for unknown names, as the reference


class BaseIndex:
    """Abstract backend (backends.py:54-163)."""

    def __init__(self, dim: int, space: str = "cosine", **kwargs: Any):
        self.dim = dim
        self.space = space
        self._index = None
        self._doc_ids: List[str] = []
        self._payloads: Dict[str, Payload] = {}
        self._is_trained = False

    def add(self, doc_id: str, embedding: np.ndarray, payload: Payload) -> None:
        raise NotImplementedError

    def build(self, **kwargs: Any) -> None:
        raise NotImplementedError

    def search(self, query: np.ndarray, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0):
        raise NotImplementedError

    def save(self, path: Union[str, Path]) -> None:
        raise NotImplementedError

    @classmethod
    def load(cls, path: Union[str, Path], **kwargs: Any) -> "BaseIndex":
        raise NotImplementedError


class LazyIds(Sequence):
    """`doc_{i:08d}` ids for bulk-ingested corpora too large for a Python list (SURVEY.md section 7
    item 10); same naming as the reference's generator (scripts/profile_index.py:52)."""

    def __init__(self, n: int, fmt: str = "doc_{:08d}"):
        self._n = int(n)
        self._fmt = fmt

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._fmt.format(j) for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._fmt.format(i)

    def index(self, doc_id: str, *args) -> int:
        prefix = self._fmt.split("{")[0]
        try:
            i = int(doc_id[len(prefix):])
        except (ValueError, TypeError):
            raise ValueError(doc_id) from None
        if doc_id != self._fmt.format(i) or not 0 <= i < self._n:
            raise ValueError(doc_id)
        return i


class ColumnStore:
    """Host mirror of bulk payload columns as a list of `[n_i, 8]` float32 chunks (PAYLOAD_FIELDS
    order); a chunk of `None` stands for `n_i` default payloads (all zeros, types.py:11-18)."""

    def __init__(self):
        self._chunks: List[Tuple[int, Optional[np.ndarray]]] = []
        self._starts: List[int] = []
        self.n = 0

    def append(self, n: int, cols: Optional[np.ndarray]) -> None:
        self._starts.append(self.n)
        self._chunks.append((n, cols))
        self.n += n

    def row(self, i: int) -> np.ndarray:
        import bisect

        c = bisect.bisect_right(self._starts, i) - 1
        n, cols = self._chunks[c]
        return np.zeros(len(PAYLOAD_FIELDS), np.float32) if cols is None else cols[i - self._starts[c]]

    def column(self, j: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.float32)
        for start, (n, cols) in zip(self._starts, self._chunks):
            if cols is not None:
                out[start:start + n] = cols[:, j]
        return out

    def all_default(self) -> bool:
        return all(cols is None for _, cols in self._chunks)


class ColumnPayloads:
    """Read-only id -> Payload mapping over a `ColumnStore`."""

    def __init__(self, ids: Sequence[str], store: ColumnStore):
        self._ids = ids
        self._store = store

    def _row(self, doc_id: str) -> Optional[int]:
        try:
            return self._ids.index(doc_id)
        except ValueError:
            return None

    def get(self, doc_id: str, default=None):
        r = self._row(doc_id)
        return default if r is None else self.at(r)

    def at(self, row: int) -> Payload:
        vals = self._store.row(row)
        return Payload(**{f: float(vals[j]) for j, f in enumerate(PAYLOAD_FIELDS)})

    def __getitem__(self, doc_id: str) -> Payload:
        r = self._row(doc_id)
        if r is None:
            raise KeyError(doc_id)
        return self.at(r)

    def __contains__(self, doc_id) -> bool:
        return self._row(doc_id) is not None

    def __len__(self) -> int:
        return len(self._ids)


def _torch():
    import torch

    return torch


class CudaIndex(BaseIndex):
    """Exact DEWI-re-ranked search on a B200 -- the drop-in for `ExactIndex` (backends.py:386-481).

    dtype  "fp32": rows kept in fp32 (+ bf16 hi/lo planes streamed by the tensor-core sweep, exact
                   fp32 re-score of the candidates) -- parity mode against the reference.
           "bf16": rows rounded to bf16 after normalisation, 2 B/element -- the 100M-row mode.
    device CUDA ordinal (default: torch's current device).
    Unknown keyword arguments (`ef`, `M`, `ef_query`, ... forwarded by the facade) are ignored, as
    `ExactIndex.__init__` does.
    """

    def __init__(self, dim: int, space: str = "cosine", dtype: str = "fp32", device: Optional[int] = None,
                 precise_query: bool = False, **kwargs: Any):
        super().__init__(dim, space)
        if space not in _native.SPACE:
            raise ValueError(f"unknown space {space!r}; expected 'cosine' or 'l2'")
        if dtype not in _native.DTYPE:
            raise ValueError(f"unknown dtype {dtype!r}; expected 'fp32' or 'bf16'")
        self.dtype = "bf16" if _native.DTYPE[dtype] == 1 else "fp32"
        self._normalize = space == "cosine"
        self._lib = _native.load_library()  # ImportError (NativeUnavailable) when the library is absent
        if device is None:
            torch = _torch()
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        _native.require_device(self.device)  # ImportError when no B200 is present
        h = ctypes.c_void_p()
        _native.check(self._lib.dewi_index_create(int(dim), _native.SPACE[space], _native.DTYPE[dtype], self.device,
                                                  ctypes.byref(h)))
        self._h = h
        self._pending: List[np.ndarray] = []   # normalised host rows not yet uploaded
        self._n_device = 0                     # rows resident on the GPU
        self._columns: Optional[ColumnStore] = None  # bulk payload columns (bulk ingest only)
        self._flags = _native.FLAG_PRECISE_QUERY if precise_query else 0
        self._host_rows: Optional[np.ndarray] = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and getattr(self, "_lib", None) is not None:
            try:
                self._lib.dewi_index_destroy(h)
            except Exception:
                pass

    # ---- ingest --------------------------------------------------------------------------------
    def add(self, doc_id: str, embedding: np.ndarray, payload: Payload) -> None:
        """One document (backends.py:394-406).  The row is normalised on the host with the very numpy
        expression the reference uses, so the stored corpus is bit-identical to `ExactIndex`'s."""
        if embedding.shape != (self.dim,):
            raise ValueError(f"Expected embedding of shape {(self.dim,)}, got {embedding.shape}")
        emb = embedding.astype(np.float32)
        if self._normalize:
            norm = np.linalg.norm(emb)
            if not norm > 0:
                # documented divergence: the reference stores a NaN row here
                raise ValueError("zero-norm embedding cannot be indexed in cosine space")
            emb = emb / norm
        if self._columns is not None:
            raise ValueError("per-document add() cannot follow a bulk add_batch() with payload columns")
        self._doc_ids.append(doc_id)
        self._payloads[doc_id] = payload
        self._pending.append(emb)
        self._is_trained = False
        self._host_rows = None

    def add_batch(self, doc_ids: Optional[Sequence[str]], embeddings, payloads: Optional[Sequence[Payload]] = None,
                  payload_columns: Optional[np.ndarray] = None, normalized: bool = False) -> None:
        """Bulk ingest of `[n, dim]` rows (numpy array, or a torch tensor on the host or this GPU).

        Rows are normalised on the device unless `normalized`.  `doc_ids=None` numbers documents
        `doc_00000000...`; payloads come either as objects or as an `[n, 8]` float array in
        `PAYLOAD_FIELDS` order."""
        torch = _torch()
        self._flush_pending()
        if isinstance(embeddings, np.ndarray):
            rows = np.ascontiguousarray(embeddings, dtype=np.float32)
            n, ptr, is_host, keep = rows.shape[0], rows.ctypes.data, 1, rows
        else:
            rows = embeddings.detach().to(dtype=torch.float32).contiguous()
            if rows.is_cuda and rows.device.index != self.device:
                raise ValueError("embeddings live on another device")
            n, ptr, is_host, keep = rows.shape[0], rows.data_ptr(), int(not rows.is_cuda), rows
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"Expected embeddings of shape (n, {self.dim}), got {tuple(rows.shape)}")
        with torch.cuda.device(self.device):
            rc = self._lib.dewi_index_append(self._h, ctypes.c_void_p(ptr), n, int(bool(normalized)), is_host,
                                             _native.stream_ptr())
            if rc != 0:
                raise ValueError(_native.last_error())
            torch.cuda.current_stream().synchronize()
        del keep
        base = self._n_device
        self._n_device += n
        if doc_ids is None:
            if base != 0 and not isinstance(self._doc_ids, LazyIds):
                self._doc_ids = list(self._doc_ids) + [f"doc_{i:08d}" for i in range(base, base + n)]
            else:
                self._doc_ids = LazyIds(base + n)
        else:
            if len(doc_ids) != n:
                raise ValueError("doc_ids and embeddings disagree in length")
            self._doc_ids = list(self._doc_ids) + list(doc_ids)
        if payloads is not None:
            if payload_columns is not None or self._columns is not None:
                raise ValueError("cannot mix payload objects and payload columns")
            if len(payloads) != n or doc_ids is None:
                raise ValueError("payload objects need matching doc_ids")
            for d, p in zip(doc_ids, payloads):
                self._payloads[d] = p
        elif payload_columns is not None or self._columns is not None or not self._payloads:
            cols = None
            if payload_columns is not None:
                cols = np.ascontiguousarray(payload_columns, dtype=np.float32)
                if cols.shape != (n, len(PAYLOAD_FIELDS)):
                    raise ValueError(f"payload_columns must be [n, {len(PAYLOAD_FIELDS)}]")
            if self._payloads and not isinstance(self._payloads, ColumnPayloads):
                raise ValueError("cannot mix payload objects and payload columns")
            if self._columns is None:
                self._columns = ColumnStore()
            self._columns.append(n, cols)
            self._payloads = ColumnPayloads(self._doc_ids, self._columns)
        else:
            for d in self._doc_ids[base:]:
                self._payloads.setdefault(d, Payload())
        self._is_trained = False
        self._host_rows = None

    def _flush_pending(self) -> None:
        if not self._pending:
            return
        torch = _torch()
        rows = np.stack(self._pending)  # backends.py:411
        with torch.cuda.device(self.device):
            _native.check(self._lib.dewi_index_append(self._h, ctypes.c_void_p(rows.ctypes.data), rows.shape[0], 1, 1,
                                                      _native.stream_ptr()))
            torch.cuda.current_stream().synchronize()
        self._n_device += rows.shape[0]
        self._pending = []

    def build(self, **kwargs: Any) -> None:
        """Upload staged rows and snapshot the payload columns the re-rank reads (backends.py:408-412)."""
        if not self._pending and self._n_device == 0:
            raise ValueError("No embeddings to build index from")
        self._flush_pending()
        self.refresh_payloads()
        self._is_trained = True

    def refresh_payloads(self) -> None:
        """Re-read `payload.dewi` and `(ht_mean + hi_mean) * 0.5` for every row (backends.py:454-458).

        The reference dereferences the shared Payload objects at *search* time; this backend snapshots
        them at `build()`.  Call this after mutating payloads of an already-built index."""
        n = self._n_device
        if n == 0:
            return
        if self._columns is not None:
            if self._columns.all_default():
                return  # device columns are zero-initialised, or were written by set_payload_columns()
            dewi = self._columns.column(0)
            ent = ((self._columns.column(1).astype(np.float64) + self._columns.column(3).astype(np.float64)) * 0.5
                   ).astype(np.float32)
        else:
            dewi = np.empty(n, dtype=np.float32)
            ent = np.empty(n, dtype=np.float32)
            pl = self._payloads
            for i, d in enumerate(self._doc_ids):
                p = pl[d]
                dewi[i] = p.dewi
                ent[i] = (p.ht_mean + p.hi_mean) * 0.5
        self.set_payload_columns(dewi, ent)

    def set_payload_columns(self, dewi, ent, offset: int = 0) -> None:
        """Write the two device payload columns directly (numpy arrays or tensors on this GPU)."""
        torch = _torch()

        def ptr(a):
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, dtype=np.float32)
                return a, a.ctypes.data, 1, a.shape[0]
            a = a.detach().to(dtype=torch.float32).contiguous()
            return a, a.data_ptr(), int(not a.is_cuda), a.shape[0]

        d, dp, dh, n = ptr(dewi)
        e, ep, eh, n2 = ptr(ent)
        if n != n2 or dh != eh:
            raise ValueError("dewi and ent must have the same length and residency")
        with torch.cuda.device(self.device):
            _native.check(self._lib.dewi_index_set_payload(self._h, ctypes.c_void_p(dp), ctypes.c_void_p(ep), int(offset), n,
                                                           dh, _native.stream_ptr()))
            torch.cuda.current_stream().synchronize()

    def set_payload_from_signals(self, signals, scorer, fit: bool = True, offset: int = 0):
        """Bulk form of the README loop (README.md:94-110; pipelines.py:199-221): `signals` is a `[7, n]`
        array / tensor in `SIGNAL_FIELDS` order for rows `[offset, offset + n)`.  Fits the scorer's robust
        statistics (unless `fit=False`), scores every row on the device and writes `dewi` and
        `(ht_mean + hi_mean) * 0.5` straight into the device payload columns -- no per-document Python.
        Returns the dewi scores (CUDA float32 tensor)."""
        torch = _torch()
        sig = torch.as_tensor(signals).to(torch.device("cuda", self.device), torch.float32)
        if sig.ndim != 2 or sig.shape[0] != 7:
            raise ValueError("expected seven signal columns")
        self._flush_pending()
        if fit:
            scorer.fit_stats_columns(sig)
        dewi = scorer.score_batch(sig)
        ent = ((sig[0].double() + sig[2].double()) * 0.5).float()  # backends.py:458
        self.set_payload_columns(dewi, ent, offset=offset)
        return dewi

    def reserve(self, rows: int) -> None:
        """Pre-size the device planes for `rows` rows (one allocation instead of geometric regrowth)."""
        torch = _torch()
        with torch.cuda.device(self.device):
            _native.check(self._lib.dewi_index_reserve(self._h, int(rows)))

    def set_id_base(self, id_base: int) -> None:
        """Global id of row 0 when this index is one shard of a row-sharded corpus."""
        _native.check(self._lib.dewi_index_set_id_base(self._h, int(id_base)))

    def __len__(self) -> int:
        return self._n_device + len(self._pending)

    # ---- search --------------------------------------------------------------------------------
    def search(self, query: np.ndarray, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0
               ) -> List[Tuple[str, float, Payload]]:
        """One query -> `[(doc_id, score, payload)]` sorted by score, descending (backends.py:414-481)."""
        query = np.asarray(query, dtype=np.float32)
        if query.ndim == 2 and query.shape[0] == 1:
            query = query[0]
        if query.shape != (self.dim,):
            raise ValueError(f"Expected query shape ({self.dim},), got {query.shape}")
        if not self._is_trained:
            self.build()
        n = self._n_device
        if min(2 * k, n) <= 0:
            return []  # backends.py:440-442
        if k > n:
            # np.argpartition(adjusted_scores, -k) raises for k > N (backends.py:468)
            raise ValueError(f"kth(=-{k}) out of bounds ({min(2 * k, n)})")
        if self._normalize:  # backends.py:420-424, same numpy expression -> same bits
            qn = np.linalg.norm(query)
            if qn > 0:
                query = query / qn
        ids, scores = self._search_host(query.reshape(1, -1), k, float(eta), float(entropy_pref),
                                        _native.FLAG_QUERY_NORMALIZED)
        doc_ids, payloads = self._doc_ids, self._payloads
        out = []
        for row, s in zip(ids[0], scores[0]):
            d = doc_ids[int(row)]
            out.append((d, float(s), payloads[d]))
        return out

    def _search_host(self, queries: np.ndarray, k: int, eta: float, entropy_pref: float, flags: int = 0):
        """Host buffers in, host buffers out, through `dewi_index_search` with DEWI_FLAG_HOST_IO."""
        torch = _torch()
        q = np.ascontiguousarray(queries, dtype=np.float32)
        b = q.shape[0]
        ids = np.empty((b, k), dtype=np.int64)
        scores = np.empty((b, k), dtype=np.float32)
        with torch.cuda.device(self.device):
            rc = self._lib.dewi_index_search(self._h, ctypes.c_void_p(q.ctypes.data), b, int(k), float(eta),
                                             float(entropy_pref), flags | self._flags | _native.FLAG_HOST_IO,
                                             ctypes.c_void_p(ids.ctypes.data), ctypes.c_void_p(scores.ctypes.data),
                                             _native.stream_ptr())
        _native.check(rc)
        return ids, scores

    def search_batch(self, queries, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0, flags: int = 0):
        """Batch extension (the reference is single-query, index.py:91-92): `[B, dim]` queries ->
        `(row_ids [B, k] int64, scores [B, k] float32)`.  numpy in -> numpy out (copies inside the
        call); CUDA tensor in -> CUDA tensors out (asynchronous on the current stream)."""
        torch = _torch()
        if not self._is_trained:
            self.build()
        if k > self._n_device:
            raise ValueError(f"k={k} exceeds the number of indexed rows ({self._n_device})")
        if isinstance(queries, np.ndarray):
            if queries.ndim != 2 or queries.shape[1] != self.dim:
                raise ValueError(f"Expected queries of shape (B, {self.dim}), got {queries.shape}")
            return self._search_host(queries, k, float(eta), float(entropy_pref), flags)
        q = queries.detach().to(dtype=torch.float32).contiguous()
        if q.ndim != 2 or q.shape[1] != self.dim or not q.is_cuda or q.device.index != self.device:
            raise ValueError(f"Expected a CUDA tensor of shape (B, {self.dim}) on device {self.device}")
        b = q.shape[0]
        ids = torch.empty((b, k), dtype=torch.int64, device=q.device)
        scores = torch.empty((b, k), dtype=torch.float32, device=q.device)
        with torch.cuda.device(self.device):
            rc = self._lib.dewi_index_search(self._h, ctypes.c_void_p(q.data_ptr()), b, int(k), float(eta),
                                             float(entropy_pref), flags | self._flags, ctypes.c_void_p(ids.data_ptr()),
                                             ctypes.c_void_p(scores.data_ptr()), _native.stream_ptr())
        _native.check(rc)
        return ids, scores

    def search_local(self, queries, kcand: int, flags: int = 0):
        """Shard-local stage (sweep + candidate selection): CUDA `[B, dim]` queries -> the shard's
        `kcand` best rows per query as `(sim, global_id, dewi, ent)` CUDA tensors `[B, kcand]`."""
        torch = _torch()
        b, dev = queries.shape[0], queries.device
        sim = torch.empty((b, kcand), dtype=torch.float32, device=dev)
        gid = torch.empty((b, kcand), dtype=torch.int64, device=dev)
        dewi = torch.empty((b, kcand), dtype=torch.float32, device=dev)
        ent = torch.empty((b, kcand), dtype=torch.float32, device=dev)
        self.search_local_into(queries, kcand, sim, gid, dewi, ent, flags)
        return sim, gid, dewi, ent

    def search_local_into(self, queries, kcand: int, sim, gid, dewi, ent, flags: int = 0) -> None:
        """`search_local` writing into caller-provided contiguous `[B, kcand]` CUDA tensors (the
        sharded path points these at its packed exchange buffer)."""
        torch = _torch()
        if not self._is_trained:
            self.build()
        q = queries.detach().to(dtype=torch.float32).contiguous()
        if q.ndim != 2 or q.shape[1] != self.dim or not q.is_cuda or q.device.index != self.device:
            raise ValueError(f"Expected a CUDA tensor of shape (B, {self.dim}) on device {self.device}")
        b = q.shape[0]
        for t, dt in ((sim, torch.float32), (gid, torch.int64), (dewi, torch.float32), (ent, torch.float32)):
            if tuple(t.shape) != (b, kcand) or t.dtype != dt or not t.is_contiguous():
                raise ValueError("output tensors must be contiguous [B, kcand] of float32/int64/float32/float32")
        with torch.cuda.device(self.device):
            rc = self._lib.dewi_index_search_local(self._h, ctypes.c_void_p(q.data_ptr()), b, int(kcand),
                                                   flags | self._flags, ctypes.c_void_p(sim.data_ptr()),
                                                   ctypes.c_void_p(gid.data_ptr()), ctypes.c_void_p(dewi.data_ptr()),
                                                   ctypes.c_void_p(ent.data_ptr()), _native.stream_ptr())
        _native.check(rc)

    def last_launches(self) -> int:
        n = ctypes.c_int(0)
        _native.check(self._lib.dewi_index_last_launches(self._h, ctypes.byref(n)))
        return n.value

    def set_profiling(self, enable: bool = True) -> None:
        """Bracket every search's sweep kernel with CUDA events (see `sweep_ms`)."""
        _native.check(self._lib.dewi_index_set_profiling(self._h, int(bool(enable))))

    def sweep_ms(self, back: int = 0) -> Tuple[float, str]:
        """Device time of the sweep kernel of the `back`-th most recent profiled search (0 = last) and
        which sweep ran ("tcgen05" / "simt")."""
        ms, kind = ctypes.c_float(0), ctypes.c_int(0)
        _native.check(self._lib.dewi_index_sweep_ms(self._h, int(back), ctypes.byref(ms), ctypes.byref(kind)))
        return ms.value, {1: "tcgen05", 2: "simt", 3: "tcgen05-pair"}.get(kind.value, "?")

    # ---- stored rows / persistence ---------------------------------------------------------------
    @property
    def _embeddings(self):
        """Stored (normalised) rows as a host array -- what `DewiIndex.get_embedding` indexes
        (index.py:101-116).  Fetched lazily; pending rows are included."""
        if self._host_rows is None:
            rows = np.empty((len(self), self.dim), dtype=np.float32)
            for i in range(self._n_device):
                _native.check(self._lib.dewi_index_get_row(self._h, i, ctypes.c_void_p(rows[i].ctypes.data)))
            for j, r in enumerate(self._pending):
                rows[self._n_device + j] = r
            self._host_rows = rows
        return self._host_rows

    def get_row(self, row: int) -> np.ndarray:
        self._flush_pending()
        out = np.empty(self.dim, dtype=np.float32)
        _native.check(self._lib.dewi_index_get_row(self._h, int(row), ctypes.c_void_p(out.ctypes.data)))
        return out

    def save(self, path: Union[str, Path]) -> None:
        """Write the directory `ExactIndex.save` writes (backends.py:483-515): `metadata.json`,
        `payloads.jsonl` (key `doc_id`), `embeddings.npy` (normalised fp32) -- loadable by either."""
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        rows = self._embeddings
        meta = {
            "dim": self.dim, "space": self.space, "doc_ids": list(self._doc_ids), "normalize": self._normalize,
            "is_trained": self._is_trained, "num_embeddings": int(rows.shape[0]),
            "type": type(self).__name__, "dtype": self.dtype,
        }
        (path / "metadata.json").write_text(json.dumps(meta))
        with open(path / "payloads.jsonl", "w") as f:
            for d in self._doc_ids:
                f.write(json.dumps({"doc_id": d, "payload": self._payloads[d].to_dict()}) + "\n")
        if rows.shape[0] > 0:
            np.save(str(path / "embeddings.npy"), rows)

    @classmethod
    def load(cls, path: Union[str, Path], **kwargs: Any) -> "CudaIndex":
        """Load a directory written by `CudaIndex.save` or by the reference's `ExactIndex.save`
        (backends.py:517-556)."""
        path = Path(path)
        meta = json.loads((path / "metadata.json").read_text())
        index = cls(dim=meta["dim"], space=meta["space"], dtype=kwargs.pop("dtype", meta.get("dtype", "fp32")), **kwargs)
        ids = list(meta["doc_ids"])
        payloads: Dict[str, Payload] = {}
        with open(path / "payloads.jsonl") as f:
            for line in f:
                rec = json.loads(line)
                payloads[rec.get("doc_id", rec.get("id"))] = Payload.from_dict(rec["payload"])
        emb_path = path / "embeddings.npy"
        if emb_path.exists() and meta.get("num_embeddings", 0) > 0:
            rows = np.load(str(emb_path)).astype(np.float32)
            index.add_batch(ids, rows, payloads=[payloads[d] for d in ids], normalized=True)
            if meta.get("is_trained", False):
                index.build()
        else:
            index._doc_ids = ids
            index._payloads = payloads
        return index
