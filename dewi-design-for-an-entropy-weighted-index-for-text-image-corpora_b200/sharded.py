"""Row-sharded search across the GPUs of one box: one process per GPU, one exchange per batch.

The corpus is partitioned into contiguous row ranges: rank g holds rows
`[base_g, base_g + n_g)` in its own `CudaIndex` whose ids are global (`set_id_base`).  A search is

    1. every rank: tensor-core sweep + candidate selection over its shard  (`dewi_index_search_local`)
       -> the shard's best `kcand = min(2k, N_total)` rows per query by similarity, with their payload
          columns, written straight into one packed block `[id | sim | dewi | ent]`
    2. one `all_gather_into_tensor` of the packed blocks (NCCL over NVLink; B * kcand * 20 bytes per rank)
    3. every rank: `dewi_rerank` reads the gathered blocks in place (shard-strided), keeps the global
       top-2k by similarity, blends (backends.py:461-465) and emits the top-k.

The global top-2k by similarity is a subset of the union of the local top-2k lists, so the result is
exactly the single-index result (`ExactIndex.search`, backends.py:414-481) -- not an approximation.

`local_search` / `rerank` can be injected: the CPU (gloo) tests of the exchange logic plug the
oracle in there; the product default is the CUDA path and there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import numpy as np

from . import _native
from .backends import CudaIndex


def _torch():
    import torch

    return torch


def shard_range(n_total: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous row range of `rank`: equal shares rounded to `align` rows, the last rank takes the
    remainder."""
    per = -(-n_total // world)
    per = -(-per // align) * align
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi


class PackedCandidates:
    """One rank's stage-1 output as a single int32 buffer, so the exchange is ONE collective:
    `[ id int64 [B, kcand] | sim f32 [B, kcand] | dewi f32 [B, kcand] | ent f32 [B, kcand] ]`."""

    WORDS_PER_CAND = 5  # int32 words: id 2 + sim 1 + dewi 1 + ent 1

    def __init__(self, b: int, kcand: int, device, world: int = 1):
        torch = _torch()
        self.b, self.kcand, self.world = b, kcand, world
        self.bk = b * kcand
        # round the per-rank block up to 8 bytes so every shard's int64 segment stays aligned
        self.words = (self.WORDS_PER_CAND * self.bk + 1) // 2 * 2
        self.local = torch.zeros(self.words, dtype=torch.int32, device=device)
        self.gathered = torch.zeros(world * self.words, dtype=torch.int32, device=device) if world > 1 else self.local

    @property
    def stride_bytes(self) -> int:
        return self.words * 4

    def views(self, buf):
        """(id, sim, dewi, ent) views of the first block of `buf`."""
        torch = _torch()
        bk = self.bk
        shape = (self.b, self.kcand)
        ids = buf[: 2 * bk].view(torch.int64).view(shape)
        sim = buf[2 * bk: 3 * bk].view(torch.float32).view(shape)
        dewi = buf[3 * bk: 4 * bk].view(torch.float32).view(shape)
        ent = buf[4 * bk: 5 * bk].view(torch.float32).view(shape)
        return ids, sim, dewi, ent


class ShardedDewiIndex:
    """The row-sharded index.  Every rank constructs it, ingests its own shard, calls `build()`
    (collective) and then `search_batch()` (collective) with the same replicated query batch."""

    def __init__(self, dim: int, space: str = "cosine", dtype: str = "bf16", group=None, device: Optional[int] = None,
                 local_index=None, local_search: Optional[Callable] = None, rerank: Optional[Callable] = None, **kwargs):
        torch = _torch()
        import torch.distributed as dist

        self.dim = dim
        self.space = space
        self.group = group
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self._local_search = local_search
        self._rerank = rerank
        if local_index is not None:
            self.local = local_index
        else:
            if device is None:
                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            self.local = CudaIndex(dim, space, dtype=dtype, device=device, **kwargs)
        self.device = getattr(self.local, "device", None)
        self.n_total = 0
        self.id_base = 0
        self._packed: Optional[PackedCandidates] = None
        self._built = False

    # ---- ingest / build --------------------------------------------------------------------------
    def add_local(self, embeddings, payload_columns=None, normalized: bool = False) -> None:
        """Append rows to THIS rank's shard (ranks ingest disjoint row ranges in rank order)."""
        self.local.add_batch(None, embeddings, payload_columns=payload_columns, normalized=normalized)
        self._built = False

    def build(self) -> None:
        """Collective: exchange shard sizes, assign global id bases, snapshot payload columns."""
        torch = _torch()
        n_local = len(self.local)
        if self.dist and self.world > 1:
            dev = self._comm_device()
            counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
            mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(counts, mine, group=self.group)
            counts = counts.cpu().tolist()
        else:
            counts = [n_local]
        self.n_total = int(sum(counts))
        self.id_base = int(sum(counts[: self.rank]))
        if self.n_total == 0:
            raise ValueError("No embeddings to build index from")
        self.local.set_id_base(self.id_base)
        if n_local > 0:
            self.local.build()
        self._built = True

    def _comm_device(self):
        torch = _torch()
        if self.device is not None and torch.cuda.is_available():
            return torch.device("cuda", self.device)
        return torch.device("cpu")

    def __len__(self) -> int:
        return self.n_total

    # ---- search ----------------------------------------------------------------------------------
    def _default_local_search(self, queries, kcand, out):
        ids, sim, dewi, ent = out
        self.local.search_local_into(queries, kcand, sim, ids, dewi, ent)

    def _default_rerank(self, packed: PackedCandidates, cand_count, k, eta, pref, out_ids, out_scores):
        torch = _torch()
        ids, sim, dewi, ent = packed.views(packed.gathered)
        lib = _native.load_library()
        with torch.cuda.device(self.device):
            rc = lib.dewi_rerank(ctypes.c_void_p(sim.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                                 ctypes.c_void_p(dewi.data_ptr()), ctypes.c_void_p(ent.data_ptr()), packed.b, packed.world,
                                 packed.kcand, packed.stride_bytes, int(cand_count), int(k), float(eta), float(pref),
                                 ctypes.c_void_p(out_ids.data_ptr()), ctypes.c_void_p(out_scores.data_ptr()),
                                 self.device, _native.stream_ptr())
        _native.check(rc)

    def search_batch(self, queries, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0):
        """Collective.  `queries`: `[B, dim]` float32 tensor on this rank's device, identical on all
        ranks.  Returns `(global_row_ids [B, k] int64, scores [B, k] float32)` on every rank."""
        torch = _torch()
        if not self._built:
            self.build()
        if k > self.n_total:
            raise ValueError(f"k={k} exceeds the number of indexed rows ({self.n_total})")  # backends.py:468
        b = queries.shape[0]
        kcand = min(2 * k, self.n_total)  # backends.py:440
        pk = self._packed
        if pk is None or pk.b != b or pk.kcand != kcand or pk.world != self.world:
            pk = self._packed = PackedCandidates(b, kcand, queries.device, self.world)
        ids, sim, dewi, ent = pk.views(pk.local)
        if len(self.local) > 0:
            (self._local_search or self._default_local_search)(queries, kcand, (ids, sim, dewi, ent))
        else:  # an empty shard contributes only empty slots
            ids.fill_(-1)
            sim.fill_(float("-inf"))
        if self.world > 1:
            self.dist.all_gather_into_tensor(pk.gathered, pk.local, group=self.group)
        out_ids = torch.empty((b, k), dtype=torch.int64, device=queries.device)
        out_scores = torch.empty((b, k), dtype=torch.float32, device=queries.device)
        (self._rerank or self._default_rerank)(pk, kcand, k, eta, entropy_pref, out_ids, out_scores)
        return out_ids, out_scores

    def search(self, query: np.ndarray, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0):
        """Single query (1-D contract of index.py:91-92) -> `[(global_row, score)]`, sorted descending."""
        torch = _torch()
        query = np.asarray(query, dtype=np.float32)
        if query.shape != (self.dim,):
            raise ValueError(f"Expected query shape ({self.dim},), got {query.shape}")
        q = torch.from_numpy(query.reshape(1, -1)).to(self._comm_device())
        ids, sc = self.search_batch(q, k, eta, entropy_pref)
        return list(zip(ids[0].cpu().tolist(), sc[0].cpu().tolist()))
