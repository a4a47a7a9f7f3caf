"""The N>1 exchange logic on CPU: world_size-2 gloo processes, with the oracle standing in (through a subclass) for the two
CUDA stages (shard-local top-2k, blend + top-k).  Checks shard ranges, global id bases, the packed
all-gather layout and that the sharded result equals the single-index result exactly."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def test_shard_range_partitions_rows():
    from dewi_b200 import shard_range

    for n, w, align in [(100, 8, 1), (1000, 3, 64), (7, 8, 1), (100_000_000, 8, 500_000)]:
        spans = [shard_range(n, w, r, align) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % align == 0 for lo, hi in spans if lo < n)


class FakeShard:
    """Stands where CudaIndex does; rows live in numpy."""

    def __init__(self, rows, pay):
        self.rows, self.pay, self.base = rows, pay, 0

    def __len__(self):
        return len(self.rows)

    def set_id_base(self, b):
        self.base = b

    def build(self):
        pass


def _worker(rank, world, port, n, d, b, k, eta, pref, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dewi_b200 import ShardedDewiIndex, shard_range
        from _util import entropy_column, make_corpus

        emb, pay = make_corpus(n, d, seed=21)
        queries = np.random.RandomState(22).standard_normal((b, d)).astype(np.float32)
        queries /= np.linalg.norm(queries, axis=1, keepdims=True)
        lo, hi = shard_range(n, world, rank, align=16)
        shard = FakeShard(emb[lo:hi], pay[lo:hi])

        def local_search(q, kcand, out):  # backends.py:431-447 on the shard, sorted descending
            ids, sim, dewi, ent = out
            s = q.numpy() @ shard.rows.T
            kk = min(kcand, s.shape[1])
            top = np.argsort(-s, axis=1, kind="stable")[:, :kk]
            ids.fill_(-1)
            sim.fill_(float("-inf"))
            ids[:, :kk] = torch.from_numpy(top + shard.base)
            sim[:, :kk] = torch.from_numpy(np.take_along_axis(s, top, 1))
            dewi[:, :kk] = torch.from_numpy(shard.pay[top, 0].astype(np.float32))
            ent[:, :kk] = torch.from_numpy(entropy_column(shard.pay)[top].astype(np.float32))

        def rerank(pk, cand_count, k_, eta_, pref_, out_ids, out_scores):  # backends.py:461-471 on the gathered blocks
            blocks = pk.gathered.view(pk.world, pk.words)
            parts = [pk.views(blocks[g]) for g in range(pk.world)]
            ids = torch.cat([p[0] for p in parts], 1).numpy()
            sim = torch.cat([p[1] for p in parts], 1).numpy()
            dewi = torch.cat([p[2] for p in parts], 1).numpy()
            ent = torch.cat([p[3] for p in parts], 1).numpy()
            for q in range(ids.shape[0]):
                order = np.argsort(-sim[q], kind="stable")[:cand_count]
                adj = (1 - eta_) * sim[q, order] + eta_ * dewi[q, order]
                if pref_ != 0:
                    adj += pref_ * ent[q, order]
                best = np.argsort(-adj, kind="stable")[:k_]
                out_ids[q] = torch.from_numpy(ids[q, order][best])
                out_scores[q] = torch.from_numpy(adj[best])

        class HostSharded(ShardedDewiIndex):  # the oracle stands in for the two CUDA stages
            def _make_local(self, *a, **kw):
                return shard

            def _local_stage(self, q, kcand, out):
                local_search(q, kcand, out)

            def _rerank_stage(self, pk, cand_count, k_, eta_, pref_, out_ids, out_scores):
                rerank(pk, cand_count, k_, eta_, pref_, out_ids, out_scores)

        ix = HostSharded(d)
        ix.build()
        assert ix.n_total == n and ix.id_base == lo and shard.base == lo
        got_ids, got_sc = ix.search_batch(torch.from_numpy(queries), k=k, eta=eta, entropy_pref=pref)
        if rank == 0:
            ret["ids"], ret["scores"] = got_ids.numpy().copy(), got_sc.numpy().copy()
        other = [None, None]
        dist.all_gather_object(other, got_ids.numpy().tobytes())
        assert other[0] == other[1], "ranks disagree on the result"
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,k", [(1000, 10), (30, 10), (17, 12)])  # incl. a shard smaller than 2k, and 2k > N
def test_two_rank_exchange_equals_single_index(n, k):
    from oracle import search as osearch

    from _util import check_topk, entropy_column, make_corpus

    d, b, eta, pref = 32, 6, 0.3, 0.5
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), n, d, b, k, eta, pref, ret), nprocs=2, join=True)
        ids, scores = ret["ids"], ret["scores"]
    emb, pay = make_corpus(n, d, seed=21)
    queries = np.random.RandomState(22).standard_normal((b, d)).astype(np.float32)
    queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), queries, k, eta, pref, True)
    for q in range(b):
        check_topk(rid[q], rsc[q], ids[q], scores[q], what=f"n{n} q{q}")
