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

The two CUDA stages are the overridable methods `_local_stage` / `_rerank_stage` (and `_make_local` for
the shard object): the CPU (gloo) tests of the exchange logic subclass the index and put the oracle
there.  The class itself has no such switch: the product path is the CUDA path, without a CPU fallback.

Fused exchange and rank skew.  With `exchange="push"` (the default when symmetric memory is available)
there is no collective on the search path: a rank's re-rank kernel waits, on the device, for its peers'
candidate blocks.  That wait is bounded in wall time (`push_timeout_s`, default 120 s, or
DEWI_PUSH_TIMEOUT_S): host-side skew between ranks (GC pauses, logging, a debugger, a first-time
rendezvous) is expected to stay far below it.  A rank that does time out does not lose its CUDA context --
the kernel reports through a host-mapped status word and returns ids of -1 for that batch; the next call
on that rank raises `RuntimeError`.  Applications that cannot bound their skew should use `exchange="nccl"`,
which simply blocks.
"""

from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _native
from .backends import MAX_CANDIDATES, CudaIndex


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


class SymmetricCandidates:
    """Gather buffers of the FUSED exchange: two (seq parity) buffers of `world` rank blocks plus `world` ready
    flags each, allocated as torch symmetric memory so that every rank's copy is mapped into every process
    (`rendezvous` is the only collective; PyTorch is plumbing here).  The finalize kernel of each rank then writes
    its block into all copies with peer stores over NVLink and the re-rank kernel acquires the flags
    (`dewi_index_search_local_push` / `dewi_rerank_gathered`): no NCCL call on the search path."""

    def __init__(self, b: int, kcand: int, device, world: int, rank: int, group):
        torch = _torch()
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.b, self.kcand, self.world, self.rank = b, kcand, world, rank
        self.seq = 0                # searches issued through this buffer (flags hold the last one per rank)
        self.bk = b * kcand
        self.words = (PackedCandidates.WORDS_PER_CAND * self.bk + 1) // 2 * 2   # int32 words per rank block (8-byte multiple)
        self.flag_words = (world + 3) // 4 * 4
        self.half_words = world * self.words + self.flag_words                  # one parity: blocks, then flags
        self.buf = symm_mem.empty(2 * self.half_words, dtype=torch.int32, device=device)
        self.buf.zero_()
        group = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)   # every copy zeroed before anybody's first push
        base = [int(p) for p in self.handle.buffer_ptrs]
        self._bases, self._flags = [], []
        for parity in range(2):
            off = parity * self.half_words * 4
            self._bases.append((ctypes.c_uint64 * world)(*[p + off for p in base]))
            self._flags.append((ctypes.c_uint64 * world)(*[p + off + world * self.words * 4 for p in base]))

    @property
    def stride_bytes(self) -> int:
        return self.words * 4

    def tables(self, seq: int):
        return self._bases[seq & 1], self._flags[seq & 1]

    def local_views(self, seq: int):
        """(id, sim, dewi, ent, flags) views of THIS rank's buffer for the parity of `seq` (rank block 0)."""
        torch = _torch()
        half = self.buf[(seq & 1) * self.half_words: ((seq & 1) + 1) * self.half_words]
        bk, shape = self.bk, (self.b, self.kcand)
        ids = half[: 2 * bk].view(torch.int64).view(shape)
        sim = half[2 * bk: 3 * bk].view(torch.float32).view(shape)
        dewi = half[3 * bk: 4 * bk].view(torch.float32).view(shape)
        ent = half[4 * bk: 5 * bk].view(torch.float32).view(shape)
        flags = half[self.world * self.words: self.world * self.words + self.world]
        return ids, sim, dewi, ent, flags


class ShardedDewiIndex:
    """The row-sharded index.  Every rank constructs it, ingests its own shard, calls `build()`
    (collective) and then `search_batch()` (collective) with the same replicated query batch."""

    def __init__(self, dim: int, space: str = "cosine", dtype: str = "bf16", group=None, device: Optional[int] = None,
                 exchange: str = "auto", push_timeout_s: float = 0.0, **kwargs):
        """exchange: "push" = fused peer-store exchange over NVLink (symmetric memory, no collective on the search
        path), "nccl" = one `all_gather_into_tensor` per batch, "auto" = push when it can be set up, else nccl.
        push_timeout_s: wall-time bound of the device-side wait for the peers' blocks (0: DEWI_PUSH_TIMEOUT_S or 120)."""
        import torch.distributed as dist

        self.dim = dim
        self.space = space
        self.group = group
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.local = self._make_local(dim, space, dtype, device, **kwargs)
        self.device = getattr(self.local, "device", None)
        self.push_timeout_s = float(push_timeout_s)
        self._status = None         # pinned host word the re-rank kernel writes when its wait times out
        self.n_total = 0
        self.id_base = 0
        self._packed: Optional[PackedCandidates] = None
        self._built = False
        if exchange not in ("auto", "push", "nccl"):
            raise ValueError("exchange must be 'auto', 'push' or 'nccl'")
        self._exchange_want = exchange
        self.exchange = "nccl"      # what the last search used
        self._symm: Optional[SymmetricCandidates] = None
        self._symm_cache = {}       # (B, kcand) -> SymmetricCandidates; each buffer carries its own sequence number
        self._symm_failed = False

    def _make_local(self, dim, space, dtype, device, **kwargs):
        """This rank's shard: a `CudaIndex` (raises ImportError without a B200 -- no CPU fallback)."""
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        return CudaIndex(dim, space, dtype=dtype, device=device, **kwargs)

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
        self._min_shard = int(min(counts))
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
    def _local_stage(self, queries, kcand, out):
        """Stage 1 + 2 on this shard: writes the packed block views `(ids, sim, dewi, ent)` in `out`."""
        ids, sim, dewi, ent = out
        self.local.search_local_into(queries, kcand, sim, ids, dewi, ent)

    def _rerank_stage(self, packed: PackedCandidates, cand_count, k, eta, pref, out_ids, out_scores):
        """Stage 3 over the gathered blocks (`dewi_rerank`, shard-strided, in place)."""
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

    def search_batch(self, queries, k: int = 10, eta: float = 0.5, entropy_pref: float = 0.0,
                     exchange: Optional[str] = None):
        """Collective.  `queries`: `[B, dim]` float32 tensor on this rank's device, identical on all
        ranks.  Returns `(global_row_ids [B, k] int64, scores [B, k] float32)` on every rank.
        `exchange` overrides the constructor's choice for this call (every rank must pass the same value)."""
        torch = _torch()
        if not self._built:
            self.build()
        if k > self.n_total:
            raise ValueError(f"k={k} exceeds the number of indexed rows ({self.n_total})")  # backends.py:468
        if exchange not in (None, "auto", "push", "nccl"):
            raise ValueError("exchange must be 'auto', 'push' or 'nccl'")
        b = queries.shape[0]
        kcand = min(2 * k, self.n_total)  # backends.py:440
        if kcand > MAX_CANDIDATES:
            raise ValueError(f"k={k}: at most {MAX_CANDIDATES} candidates (min(2k, N)) are re-ranked per query")
        self._check_status()
        full = getattr(self.local, "rerank_scope", "candidates") == "full"
        if full:
            # opt-in full-corpus blend: every shard selects by the blended key (its global top-k lies in the union of the
            # shards' local ones), exchanged with the all-gather -- the fused push lives in the candidate-scope tail kernel
            if exchange == "push":
                raise ValueError("rerank_scope='full' uses the all-gather exchange")
            self.local.set_blend(eta, entropy_pref)
        if not full and self._push_ready(b, kcand, queries, exchange or self._exchange_want):
            return self._search_batch_push(queries, b, kcand, k, eta, entropy_pref)
        self.exchange = "nccl"
        pk = self._packed
        if pk is None or pk.b != b or pk.kcand != kcand or pk.world != self.world:
            pk = self._packed = PackedCandidates(b, kcand, queries.device, self.world)
        ids, sim, dewi, ent = pk.views(pk.local)
        if len(self.local) > 0:
            self._local_stage(queries, kcand, (ids, sim, dewi, ent))
        else:  # an empty shard contributes only empty slots
            ids.fill_(-1)
            sim.fill_(float("-inf"))
        if self.world > 1:
            self.dist.all_gather_into_tensor(pk.gathered, pk.local, group=self.group)
        out_ids = torch.empty((b, k), dtype=torch.int64, device=queries.device)
        out_scores = torch.empty((b, k), dtype=torch.float32, device=queries.device)
        # candidate scope: the global top-2k by similarity are blended (backends.py:439-481); full scope: every gathered
        # candidate is (each shard already selected by the blended key)
        self._rerank_stage(pk, kcand * self.world if full else kcand, k, eta, entropy_pref, out_ids, out_scores)
        return out_ids, out_scores

    def _check_status(self) -> None:
        """Raise if an earlier fused search on this rank gave up waiting for a peer (its results were ids of -1)."""
        if self._status is not None and int(self._status[0]) != 0:
            seq = int(self._status[0])
            self._status[0] = 0
            raise RuntimeError(f"fused exchange: search #{seq} on rank {self.rank} timed out waiting for a peer's candidate "
                               f"block (skew above push_timeout_s, or a dead peer); its results were invalidated")

    # ---- fused exchange (peer stores over NVLink instead of an all-gather) ---------------------------
    def _push_ready(self, b: int, kcand: int, queries, want: str) -> bool:
        """Collective decision (identical on every rank: it depends only on replicated state and on whether the
        symmetric allocation succeeded everywhere)."""
        torch = _torch()
        if (want == "nccl" or self._symm_failed or not self.dist or self.world <= 1 or not queries.is_cuda
                or self.world > 16 or getattr(self, "_min_shard", 0) <= 0):
            if want == "push" and self.world > 1 and queries.is_cuda:
                raise RuntimeError("fused exchange requested but unavailable (symmetric memory failed, an empty shard, "
                                   "or more than 16 ranks)")
            return False
        sy = self._symm_cache.get((b, kcand))
        if sy is not None:   # the cache evolves identically on every rank (replicated query shapes)
            self._symm = sy
            return True
        ok = 1
        try:
            sy = SymmetricCandidates(b, kcand, queries.device, self.world, self.rank, self.group)
        except Exception as exc:  # symmetric memory unavailable on this box / build
            ok, sy, err = 0, None, exc
        flag = torch.tensor([ok], dtype=torch.int32, device=queries.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._symm_failed = True
            self._symm = None
            if want == "push":
                raise RuntimeError(f"fused exchange requested but symmetric memory could not be set up: {err if not ok else 'on a peer'}")
            return False
        if len(self._symm_cache) >= 16:   # drop the oldest shape (same order on every rank)
            self._symm_cache.pop(next(iter(self._symm_cache)))
        self._symm_cache[(b, kcand)] = sy
        self._symm = sy
        return True

    def _search_batch_push(self, queries, b: int, kcand: int, k: int, eta: float, entropy_pref: float):
        torch = _torch()
        sy = self._symm
        sy.seq += 1
        seq = sy.seq
        bases, flag_tabs = sy.tables(seq)
        lib = _native.load_library()
        q = queries.detach().to(dtype=torch.float32).contiguous()
        out_ids = torch.empty((b, k), dtype=torch.int64, device=queries.device)
        out_scores = torch.empty((b, k), dtype=torch.float32, device=queries.device)
        ids, sim, dewi, ent, flags = sy.local_views(seq)
        if self._status is None:
            # pinned (host-mapped) word: the kernel can report a timed-out wait without a device->host copy
            self._status = torch.zeros(1, dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.device):
            rc = lib.dewi_index_search_local_push(self.local._h, ctypes.c_void_p(q.data_ptr()), b, kcand, self.local._flags,
                                                  self.world, self.rank, bases, flag_tabs, sy.stride_bytes, seq,
                                                  _native.stream_ptr())
            _native.check(rc)
            rc = lib.dewi_rerank_gathered(ctypes.c_void_p(sim.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                                          ctypes.c_void_p(dewi.data_ptr()), ctypes.c_void_p(ent.data_ptr()), b, self.world, kcand,
                                          sy.stride_bytes, int(kcand), int(k), float(eta), float(entropy_pref),
                                          ctypes.c_void_p(out_ids.data_ptr()), ctypes.c_void_p(out_scores.data_ptr()),
                                          ctypes.c_void_p(flags.data_ptr()), seq, ctypes.c_void_p(self._status.data_ptr()),
                                          self.push_timeout_s, self.device, _native.stream_ptr())
        _native.check(rc)
        self.exchange = "push"
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
