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

    def write(self, offset: int, cols: np.ndarray) -> None:
        """Overwrite rows `[offset, offset + len(cols))` (they may span chunks; default chunks are
        materialised on first write)."""
        import bisect

        cols = np.ascontiguousarray(cols, dtype=np.float32)
        if offset < 0 or offset + cols.shape[0] > self.n:
            raise ValueError("payload rows outside the store")
        done = 0
        while done < cols.shape[0]:
            c = bisect.bisect_right(self._starts, offset + done) - 1
            n_c, have = self._chunks[c]
            if have is None:
                have = np.zeros((n_c, len(PAYLOAD_FIELDS)), dtype=np.float32)
                self._chunks[c] = (n_c, have)
            at = offset + done - self._starts[c]
            m = min(n_c - at, cols.shape[0] - done)
            have[at:at + m] = cols[done:done + m]
            done += m


class ColumnPayloads:
    """Read-only id -> Payload mapping over a `ColumnStore`."""

    def __init__(self, ids: Sequence[str], store: ColumnStore):
        self._ids = ids
        self._store = store
        self._lookup: Optional[Dict[str, int]] = None

    def _row(self, doc_id: str) -> Optional[int]:
        ids = self._ids
        if isinstance(ids, LazyIds):  # the row is parsed out of the id
            try:
                return ids.index(doc_id)
            except ValueError:
                return None
        # explicit ids: one dict built on first use instead of an O(n) `list.index` per lookup (a `save()` of n documents
        # would otherwise cost n^2 / 2 string comparisons).  A duplicated id resolves to its LAST row, which is the
        # payload the reference's `_payloads[doc_id] = payload` keeps (backends.py:400).
        if self._lookup is None:   # (the id list never changes under a view: every bulk ingest builds a new one)
            self._lookup = {d: i for i, d in enumerate(ids)}
        return self._lookup.get(doc_id)

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


# candidates (min(2k, N), backends.py:440) one search keeps per query: the exact sweep's shared-memory lists hold 400
MAX_CANDIDATES = 400


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
                 precise_query: bool = False, rerank_scope: str = "candidates", **kwargs: Any):
        super().__init__(dim, space)
        if rerank_scope not in ("candidates", "full"):
            raise ValueError("rerank_scope must be 'candidates' (the reference's two-stage re-rank) or 'full'")
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
        # "full": the blend is applied over the WHOLE corpus instead of the top-2k by similarity -- an opt-in that is
        # NOT the reference's semantics (backends.py:439-481 re-ranks candidates only)
        self.rerank_scope = rerank_scope
        if rerank_scope == "full":
            self._flags |= _native.FLAG_SCOPE_FULL
        self._host_rows: Optional[np.ndarray] = None
        # True once the device payload columns were written without a host mirror (set_payload_columns /
        # set_payload_from_signals(mirror=False)): build() / refresh_payloads() then leave them alone
        self._device_payload_direct = False

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
        # every argument check happens BEFORE the rows reach the device: a ValueError must leave the index unchanged
        if doc_ids is not None and len(doc_ids) != n:
            raise ValueError("doc_ids and embeddings disagree in length")
        cols = None
        if payloads is not None:
            if payload_columns is not None or self._columns is not None:
                raise ValueError("cannot mix payload objects and payload columns")
            if len(payloads) != n or doc_ids is None:
                raise ValueError("payload objects need matching doc_ids")
        elif payload_columns is not None or self._columns is not None or not self._payloads:
            if payload_columns is not None:
                cols = np.ascontiguousarray(payload_columns, dtype=np.float32)
                if cols.shape != (n, len(PAYLOAD_FIELDS)):
                    raise ValueError(f"payload_columns must be [n, {len(PAYLOAD_FIELDS)}]")
            if self._payloads and not isinstance(self._payloads, ColumnPayloads):
                raise ValueError("cannot mix payload objects and payload columns")
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
            self._doc_ids = list(self._doc_ids) + list(doc_ids)
        if payloads is not None:
            for d, p in zip(doc_ids, payloads):
                self._payloads[d] = p
        elif payload_columns is not None or self._columns is not None or not self._payloads:
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
        if n == 0 or self._device_payload_direct:
            return  # (device columns written directly are authoritative: nothing on the host to re-read)
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
        self._write_payload_columns(dewi, ent)

    def set_payload_columns(self, dewi, ent, offset: int = 0) -> None:
        """Write the two device payload columns directly (numpy arrays or tensors on this GPU).  The host-side
        Payload view is NOT updated: from here on the device columns are authoritative -- `build()` /
        `refresh_payloads()` no longer overwrite them, `search()` tuples carry the host view (defaults for a bulk
        ingest without payload columns) and `save()` refuses rather than persist stale payloads.  Use
        `set_payload_from_signals` (mirrored) or `add_batch(payload_columns=...)` to keep both sides in step."""
        self._write_payload_columns(dewi, ent, offset)
        self._device_payload_direct = True

    def get_payload_columns(self, offset: int = 0, n: Optional[int] = None):
        """The device payload columns `(dewi, ent)` of rows `[offset, offset + n)` as float32 numpy arrays."""
        torch = _torch()
        self._flush_pending()
        n = self._n_device - offset if n is None else int(n)
        dewi = np.empty(n, dtype=np.float32)
        ent = np.empty(n, dtype=np.float32)
        with torch.cuda.device(self.device):
            _native.check(self._lib.dewi_index_get_payload(self._h, int(offset), n, ctypes.c_void_p(dewi.ctypes.data),
                                                           ctypes.c_void_p(ent.ctypes.data), 1, _native.stream_ptr()))
        return dewi, ent

    def _write_payload_columns(self, dewi, ent, offset: int = 0) -> None:
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

    def set_payload_from_signals(self, signals, scorer, fit: bool = True, offset: int = 0, mirror: bool = True):
        """Bulk form of the README loop (README.md:94-110; pipelines.py:199-221): `signals` is a `[7, n]`
        array / tensor in `SIGNAL_FIELDS` order for rows `[offset, offset + n)`.  Fits the scorer's robust
        statistics (unless `fit=False`), scores every row on the device and writes `dewi` and
        `(ht_mean + hi_mean) * 0.5` straight into the device payload columns -- no per-document Python on
        the scoring path.  Returns the dewi scores (CUDA float32 tensor).

        mirror=True (default) also writes the scores and the seven signals into the HOST payload view (the
        column store of a bulk ingest, or the shared `Payload` objects of per-document `add`), which is what
        the README loop leaves behind (`payload.dewi = scorer.score(signals)`): a later `build()`,
        `refresh_payloads()`, `search()` tuple or `save()` sees the same values as the device.
        mirror=False skips that copy (32 B/row of host memory and a device->host transfer -- the 100M-row
        case) and marks the device columns authoritative instead (see `set_payload_columns`)."""
        torch = _torch()
        sig = torch.as_tensor(signals).to(torch.device("cuda", self.device), torch.float32)
        if sig.ndim != 2 or sig.shape[0] != 7:
            raise ValueError("expected seven signal columns")
        self._flush_pending()
        n = sig.shape[1]
        if offset < 0 or offset + n > self._n_device:
            raise ValueError("signal rows outside the corpus")
        if fit:
            scorer.fit_stats_columns(sig)
        dewi = scorer.score_batch(sig)
        ent = ((sig[0].double() + sig[2].double()) * 0.5).float()  # backends.py:458
        self._write_payload_columns(dewi, ent, offset=offset)
        if not mirror:
            self._device_payload_direct = True
            return dewi
        cols = np.empty((n, len(PAYLOAD_FIELDS)), dtype=np.float32)
        cols[:, 0] = dewi.cpu().numpy()
        cols[:, 1:] = sig.t().cpu().numpy()
        if self._columns is not None:
            self._columns.write(offset, cols)
        else:
            pl = self._payloads
            for r, d in enumerate(self._doc_ids[offset:offset + n]):
                p = pl[d]
                for j, f in enumerate(PAYLOAD_FIELDS):
                    setattr(p, f, float(cols[r, j]))
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
    def _check_candidate_limit(self, k: int) -> None:
        """Documented limit: the candidate lists live in shared memory."""
        # (rerank_scope="full" over-fetches eight slots: the sweep's fused key and the exact blend may order near-ties differently)
        limit = MAX_CANDIDATES - (8 if self.rerank_scope == "full" and self._n_device > MAX_CANDIDATES - 8 else 0)
        if min(2 * k, self._n_device) > limit:
            raise ValueError(f"k={k}: this backend re-ranks at most {limit} candidates (min(2k, N)) per query, "
                             f"i.e. k <= {limit // 2} on a corpus of more than {limit} rows")

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
        self._check_candidate_limit(k)
        if self._normalize:  # backends.py:420-424, same numpy expression -> same bits
            qn = np.linalg.norm(query)
            if qn > 0:
                query = query / qn
        ids, scores = self._search_host(query.reshape(1, -1), k, float(eta), float(entropy_pref),
                                        _native.FLAG_QUERY_NORMALIZED)
        doc_ids, payloads = self._doc_ids, self._payloads
        by_row = payloads.at if isinstance(payloads, ColumnPayloads) else None  # (column store: no id -> row search)
        out = []
        for row, s in zip(ids[0].tolist(), scores[0].tolist()):
            d = doc_ids[row]
            out.append((d, s, by_row(row) if by_row else payloads[d]))
        return out

    def _search_host(self, queries: np.ndarray, k: int, eta: float, entropy_pref: float, flags: int = 0):
        """Host buffers in, host buffers out, through `dewi_index_search` with DEWI_FLAG_HOST_IO."""
        torch = _torch()
        q = np.ascontiguousarray(queries, dtype=np.float32)
        b = q.shape[0]
        ids = np.empty((b, k), dtype=np.int64)
        scores = np.empty((b, k), dtype=np.float32)
        # (the library selects the handle's device itself and restores the caller's; the stream is that device's current one)
        rc = self._lib.dewi_index_search(self._h, ctypes.c_void_p(q.ctypes.data), b, int(k), float(eta),
                                         float(entropy_pref), flags | self._flags | _native.FLAG_HOST_IO,
                                         ctypes.c_void_p(ids.ctypes.data), ctypes.c_void_p(scores.ctypes.data),
                                         ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
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
        self._check_candidate_limit(k)
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

    def set_blend(self, eta: float, entropy_pref: float) -> None:
        """Weights of the blended key a `rerank_scope="full"` shard-local search selects by (`search_batch` sets them
        itself; the sharded index calls this before its local stage)."""
        _native.check(self._lib.dewi_index_set_blend(self._h, float(eta), float(entropy_pref)))

    def cert_stats(self) -> Tuple[int, int]:
        """fp32 index: `(searches answered by the certified single-plane sweep, of which re-run with the full hi/lo
        product because the certificate could not be given)`."""
        used, failed = ctypes.c_int64(0), ctypes.c_int64(0)
        _native.check(self._lib.dewi_index_cert_stats(self._h, ctypes.byref(used), ctypes.byref(failed)))
        return used.value, failed.value

    def last_launches(self) -> int:
        n = ctypes.c_int(0)
        _native.check(self._lib.dewi_index_last_launches(self._h, ctypes.byref(n)))
        return n.value

    def set_profiling(self, enable: bool = True) -> None:
        """Bracket every search's sweep kernel with CUDA events (see `sweep_ms`)."""
        _native.check(self._lib.dewi_index_set_profiling(self._h, int(bool(enable))))

    def sweep_ms(self, back: int = 0) -> Tuple[float, str]:
        """Device time of the sweep kernel of the `back`-th most recent profiled search (0 = last) and
        which sweep ran ("tcgen05" / "tcgen05-rows" / "tcgen05-pair" / "simt")."""
        ms, kind = ctypes.c_float(0), ctypes.c_int(0)
        _native.check(self._lib.dewi_index_sweep_ms(self._h, int(back), ctypes.byref(ms), ctypes.byref(kind)))
        return ms.value, {1: "tcgen05", 2: "simt", 3: "tcgen05-pair", 4: "tcgen05-rows"}.get(kind.value, "?")

    # ---- stored rows / persistence ---------------------------------------------------------------
    @property
    def _embeddings(self):
        """Stored (normalised) rows as a host array -- what `DewiIndex.get_embedding` indexes
        (index.py:101-116).  Fetched lazily; pending rows are included."""
        if self._host_rows is None:
            rows = np.empty((len(self), self.dim), dtype=np.float32)
            if self._n_device:
                self.export_rows(0, self._n_device, out=rows[: self._n_device])
            for j, r in enumerate(self._pending):
                rows[self._n_device + j] = r
            self._host_rows = rows
        return self._host_rows

    def export_rows(self, row0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Stored (normalised) rows `[row0, row0 + n)` as a float32 host array: ONE bulk device->host export in
        large chunks (`dewi_index_export_rows`), never a transfer per row."""
        torch = _torch()
        n = self._n_device - row0 if n is None else int(n)
        if out is None:
            out = np.empty((n, self.dim), dtype=np.float32)
        if out.shape != (n, self.dim) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 [n, dim] array")
        with torch.cuda.device(self.device):
            _native.check(self._lib.dewi_index_export_rows(self._h, int(row0), n, ctypes.c_void_p(out.ctypes.data), 1,
                                                           _native.stream_ptr()))
        return out

    def get_row(self, row: int) -> np.ndarray:
        self._flush_pending()
        out = np.empty(self.dim, dtype=np.float32)
        _native.check(self._lib.dewi_index_get_row(self._h, int(row), ctypes.c_void_p(out.ctypes.data)))
        return out

    SIDECAR_SHARD_ROWS = 4_000_000   # rows per bf16 sidecar file (6 GB at dim 768)
    FP32_EXPORT_CHUNK = 1_000_000    # rows per bulk export while writing embeddings.npy

    def save(self, path: Union[str, Path], fp32: Optional[bool] = None, sidecar: Optional[bool] = None) -> None:
        """Write the directory `ExactIndex.save` writes (backends.py:483-515): `metadata.json`,
        `payloads.jsonl` (key `doc_id`), `embeddings.npy` (normalised fp32) -- loadable by either backend.

        Rows leave the device in bulk (`dewi_index_export_rows`, ~1M rows per transfer) straight into a
        memory-mapped `embeddings.npy`; nothing is copied per row and the matrix is never held twice.
        A bf16-storage index also writes a sharded SIDECAR of the raw bf16 rows
        (`embeddings_bf16.NNNN.npy`, uint16 `[rows, dim]`, `SIDECAR_SHARD_ROWS` rows each, listed in
        `metadata.json["bf16_sidecar"]`): `CudaIndex.load` prefers it (half the bytes, bit-exact planes, no
        re-rounding); the reference ignores it.  `fp32=False` skips `embeddings.npy` (default: skipped only
        when it would exceed 64 GB -- 100M x 768 fp32 is 307 GB; such a directory is for this backend only),
        `sidecar=False` skips the sidecar."""
        torch = _torch()
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        if self._device_payload_direct:
            raise ValueError("the device payload columns were written without a host mirror (set_payload_columns / "
                             "set_payload_from_signals(mirror=False)): saving would persist stale payloads")
        self._flush_pending()
        n = self._n_device
        if fp32 is None:
            fp32 = n * self.dim * 4 <= (64 << 30)
        if sidecar is None:
            sidecar = self.dtype == "bf16"
        if sidecar and self.dtype != "bf16":
            raise ValueError("the bf16 sidecar exists for bf16-storage indices only")
        meta = {
            "dim": self.dim, "space": self.space, "doc_ids": list(self._doc_ids), "normalize": self._normalize,
            "is_trained": self._is_trained, "num_embeddings": int(n),
            "type": type(self).__name__, "dtype": self.dtype, "has_fp32": bool(fp32 and n > 0),
        }
        if n > 0 and fp32:
            if self._host_rows is not None:
                np.save(str(path / "embeddings.npy"), self._host_rows)
            else:
                out = np.lib.format.open_memmap(str(path / "embeddings.npy"), mode="w+", dtype=np.float32, shape=(n, self.dim))
                buf = np.empty((min(n, self.FP32_EXPORT_CHUNK), self.dim), dtype=np.float32)
                for lo in range(0, n, self.FP32_EXPORT_CHUNK):
                    m = min(self.FP32_EXPORT_CHUNK, n - lo)
                    self.export_rows(lo, m, out=buf[:m])
                    out[lo:lo + m] = buf[:m]
                out.flush()
                del out
        if n > 0 and sidecar:
            shards = []
            for s, lo in enumerate(range(0, n, self.SIDECAR_SHARD_ROWS)):
                m = min(self.SIDECAR_SHARD_ROWS, n - lo)
                raw = np.empty((m, self.dim), dtype=np.uint16)
                with torch.cuda.device(self.device):
                    _native.check(self._lib.dewi_index_export_bf16(self._h, lo, m, ctypes.c_void_p(raw.ctypes.data), 1,
                                                                   _native.stream_ptr()))
                name = f"embeddings_bf16.{s:04d}.npy"
                np.save(str(path / name), raw)
                shards.append({"file": name, "row0": lo, "rows": m})
            meta["bf16_sidecar"] = shards
        (path / "metadata.json").write_text(json.dumps(meta))
        by_row = self._payloads.at if isinstance(self._payloads, ColumnPayloads) else None  # (column store: no id -> row search)
        with open(path / "payloads.jsonl", "w") as f:
            for i, d in enumerate(self._doc_ids):
                p = by_row(i) if by_row else self._payloads[d]
                f.write(json.dumps({"doc_id": d, "payload": p.to_dict()}) + "\n")

    @classmethod
    def load(cls, path: Union[str, Path], **kwargs: Any) -> "CudaIndex":
        """Load a directory written by `CudaIndex.save` or by the reference's `ExactIndex.save`
        (backends.py:517-556).  A bf16 sidecar, when present and the index is opened in bf16 storage, is
        uploaded as-is (`dewi_index_append_bf16`); otherwise `embeddings.npy` is read through a memory map and
        ingested in bulk."""
        torch = _torch()
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
        shards = meta.get("bf16_sidecar") or []
        n = int(meta.get("num_embeddings", 0))
        if n > 0 and shards and index.dtype == "bf16":
            index.reserve(n)
            for sh in shards:
                raw = np.ascontiguousarray(np.load(str(path / sh["file"]), mmap_mode="r"), dtype=np.uint16)
                if raw.shape != (sh["rows"], index.dim):
                    raise ValueError(f"{sh['file']}: unexpected shape {raw.shape}")
                with torch.cuda.device(index.device):
                    _native.check(index._lib.dewi_index_append_bf16(index._h, ctypes.c_void_p(raw.ctypes.data), raw.shape[0], 1,
                                                                    _native.stream_ptr()))
                index._n_device += raw.shape[0]
            index._doc_ids = ids
            index._payloads = {d: payloads[d] for d in ids}
            if meta.get("is_trained", False):
                index.build()
        elif emb_path.exists() and n > 0:
            rows = np.load(str(emb_path), mmap_mode="r")
            index.reserve(n)
            for lo in range(0, n, cls.FP32_EXPORT_CHUNK):
                hi = min(lo + cls.FP32_EXPORT_CHUNK, n)
                block = np.ascontiguousarray(rows[lo:hi], dtype=np.float32)
                index.add_batch(ids[lo:hi], block, payloads=[payloads[d] for d in ids[lo:hi]], normalized=True)
            if meta.get("is_trained", False):
                index.build()
        else:
            if n > 0:
                raise ValueError(f"{path}: neither embeddings.npy nor a usable bf16 sidecar holds the {n} rows")
            index._doc_ids = ids
            index._payloads = payloads
        return index
