"""Row-sharded search over real GPUs (NCCL): needs >= 2 devices, one process per GPU."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, d, b, k, dtype, exchange, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dewi_b200 import ShardedDewiIndex, shard_range
        from _util import make_corpus

        emb, pay = make_corpus(n, d, seed=41)
        queries = np.random.RandomState(42).standard_normal((b, d)).astype(np.float32)
        lo, hi = shard_range(n, world, rank, align=128)
        ix = ShardedDewiIndex(d, dtype=dtype, device=rank, exchange=exchange)
        ix.add_local(emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        ix.build()
        assert len(ix) == n and ix.id_base == lo
        q_dev = torch.from_numpy(queries).cuda()
        try:
            ids, sc = ix.search_batch(q_dev, k=k, eta=0.3, entropy_pref=0.5)
        except RuntimeError as exc:
            if exchange == "push" and "symmetric memory" in str(exc):
                ret["unavailable"] = str(exc)
                return
            raise
        # several searches in a row: the fused exchange alternates between two buffers (seq parity) and must
        # keep returning the same answer, also when the ranks are deliberately skewed in time
        for it in range(5):
            if it % 2 == rank % 2:
                torch.cuda._sleep(20_000_000)
            ids2, sc2 = ix.search_batch(q_dev, k=k, eta=0.3, entropy_pref=0.5)
            assert torch.equal(ids2, ids) and torch.equal(sc2, sc)
        ids3, sc3 = ix.search_batch(q_dev[: b // 2], k=k, eta=0.3, entropy_pref=0.5)   # another buffer shape
        assert torch.equal(ids3, ids[: b // 2]) and torch.equal(sc3, sc[: b // 2])
        ids4, sc4 = ix.search_batch(q_dev, k=k, eta=0.3, entropy_pref=0.5)              # back to the first (cached) shape
        assert torch.equal(ids4, ids) and torch.equal(sc4, sc)
        one = ix.search(queries[3], k=k, eta=0.3, entropy_pref=0.5)                      # single-query wrapper (B = 1)
        assert [r for r, _ in one] == ids[3].cpu().tolist()
        torch.cuda.synchronize()
        assert ix.exchange == exchange
        if rank == 0:
            ret["ids"], ret["scores"] = ids.cpu().numpy(), sc.cpu().numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "push"])  # one all-gather per batch / fused peer stores over NVLink
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_nccl_sharded_search_matches_oracle(dtype, exchange):
    import torch.multiprocessing as mp

    from oracle import search as osearch

    from _util import bf16_round, check_topk, entropy_column, make_corpus, recall_at_k

    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    world = min(world, 8)
    n, d, b, k = 60_000, 256, 32, 10
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, d, b, k, dtype, exchange, ret), nprocs=world, join=True)
        if "unavailable" in ret:
            pytest.skip(f"fused exchange unavailable here: {ret['unavailable']}")
        ids, scores = ret["ids"], ret["scores"]
    emb, pay = make_corpus(n, d, seed=41)
    queries = np.random.RandomState(42).standard_normal((b, d)).astype(np.float32)
    rows = emb if dtype == "fp32" else bf16_round(emb)
    rid, rsc = osearch.exact_search_batch(rows, pay[:, 0], entropy_column(pay), queries, k, 0.3, 0.5, True)
    if dtype == "fp32":
        for q in range(b):
            check_topk(rid[q], rsc[q], ids[q], scores[q], what=f"nccl q{q}")
    else:
        assert recall_at_k(rid, ids) >= 0.999


def _full_scope_worker(rank, world, port, n, d, b, k, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dewi_b200 import ShardedDewiIndex, shard_range
        from _util import make_corpus

        emb, pay = make_corpus(n, d, seed=51, style="readme")
        queries = np.random.RandomState(52).standard_normal((b, d)).astype(np.float32)
        lo, hi = shard_range(n, world, rank, align=128)
        ix = ShardedDewiIndex(d, dtype="bf16", device=rank, rerank_scope="full")
        ix.add_local(emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        ix.build()
        ids, sc = ix.search_batch(torch.from_numpy(queries).cuda(), k=k, eta=0.3, entropy_pref=0.5)
        torch.cuda.synchronize()
        assert ix.exchange == "nccl"
        if rank == 0:
            ret["ids"], ret["scores"] = ids.cpu().numpy(), sc.cpu().numpy()
    finally:
        dist.destroy_process_group()


def test_sharded_full_scope_matches_the_blend_over_every_row():
    """rerank_scope="full" across real GPUs: every shard selects by the blended key, one all-gather, one re-rank."""
    import torch.multiprocessing as mp

    from oracle import search as osearch

    from _util import bf16_round, entropy_column, make_corpus, recall_at_k

    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    world = min(world, 8)
    n, d, b, k = 300_000, 128, 20, 10
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_full_scope_worker, args=(world, port, n, d, b, k, ret), nprocs=world, join=True)
        ids, scores = ret["ids"], ret["scores"]
    emb, pay = make_corpus(n, d, seed=51, style="readme")
    queries = np.random.RandomState(52).standard_normal((b, d)).astype(np.float32)
    rows, ent = bf16_round(emb), entropy_column(pay)
    ref = [osearch.full_scope_search(rows, pay[:, 0], ent, q, k, 0.3, 0.5, True) for q in queries]
    assert recall_at_k(np.stack([r[0] for r in ref]), ids) >= 0.999
    np.testing.assert_allclose(scores, np.stack([r[1] for r in ref]), rtol=2e-5, atol=2e-6)


def _join_worker(rank, world, port, n, d, tau, align, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import dewi_b200
        from dewi_b200 import shard_range

        rng = np.random.RandomState(71)
        x = rng.standard_normal((n, d)).astype(np.float32)
        x[1::50] = x[0:-1:50] * 1.1 + 0.02 * rng.standard_normal((len(x[1::50]), d)).astype(np.float32)
        lo, hi = shard_range(n, world, rank, align=align)
        out = dewi_b200.sharded_self_join(torch.from_numpy(x[lo:hi]).cuda(), tau=tau, precision="fp32")
        assert out["row_offset"] == lo
        ret[rank] = (out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist(), out["max_sim"].cpu().numpy(),
                     out["count"].cpu().numpy(), out["argmax"].cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("align", [256, 64])  # 256: symmetric range join + all-reduced statistics; 64: a_offset slices
def test_nccl_sharded_self_join_matches_oracle(align):
    import torch.multiprocessing as mp

    from oracle import redundancy as ored

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    n, d, tau = 6000, 128, 0.93
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_join_worker, args=(world, port, n, d, tau, align, ret), nprocs=world, join=True)
        parts = [ret[r] for r in range(world)]
    rng = np.random.RandomState(71)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[1::50] = x[0:-1:50] * 1.1 + 0.02 * rng.standard_normal((len(x[1::50]), d)).astype(np.float32)
    mx, am, cnt, pairs = ored.join_rowstats(x, x, tau, self_join=True)
    got = set()
    n_emitted = 0
    for p in parts:
        got |= set(zip(p[0], p[1]))
        n_emitted += len(p[0])
    assert got == {(i, j) for i, j, _ in pairs} and len(got) > 50 and n_emitted == len(got)
    np.testing.assert_allclose(np.concatenate([p[2] for p in parts]), mx, atol=1e-5)
    np.testing.assert_array_equal(np.concatenate([p[3] for p in parts]), cnt)
    sim = ored.cross_modal_similarity(x, x)
    np.fill_diagonal(sim, -np.inf)
    arg = np.concatenate([p[4] for p in parts])
    assert np.all(sim[np.arange(n), arg] >= mx - 1e-5)


def _timeout_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dewi_b200 import ShardedDewiIndex, shard_range
        from _util import make_corpus

        n, d, b, k = 20_000, 128, 8, 5
        emb, pay = make_corpus(n, d, seed=5)
        lo, hi = shard_range(n, world, rank, align=128)
        ix = ShardedDewiIndex(d, dtype="fp32", device=rank, exchange="push", push_timeout_s=0.4)
        ix.add_local(emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        ix.build()
        q = torch.from_numpy(np.random.RandomState(1).standard_normal((b, d)).astype(np.float32)).cuda()
        try:
            good, _ = ix.search_batch(q, k=k)
        except RuntimeError as exc:
            if "symmetric memory" in str(exc) or "unavailable" in str(exc):
                ret["unavailable"] = str(exc)
                return
            raise
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 1:
            torch.cuda._sleep(int(3.0e9))            # ~1.5 s of device time in front of rank 1's search
        ids, _ = ix.search_batch(q, k=k)             # rank 0 gives up after 0.4 s, without trapping
        torch.cuda.synchronize()
        if rank == 0:
            ret["timed_out_ids_are_minus_one"] = bool((ids == -1).all().item())
            try:
                ix.search_batch(q, k=k)
                ret["raised"] = False
            except RuntimeError as exc:
                ret["raised"] = "timed out" in str(exc)
            # the context survived: the local shard still answers
            sim, gid, _, _ = ix.local.search_local(q, 2 * k)
            torch.cuda.synchronize()
            ret["context_alive"] = bool((gid[:, 0] >= 0).all().item())
        else:
            ret["late_rank_ok"] = bool(torch.equal(ids, good))
        torch.cuda.synchronize()
        dist.barrier()                               # nobody unmaps its buffers while a peer may still push
    finally:
        dist.destroy_process_group()


def test_push_timeout_reports_instead_of_trapping():
    """A peer that is late by more than `push_timeout_s` costs the waiting rank ONE failed search (ids of -1, then a
    RuntimeError on its next call) -- not a sticky CUDA error that would lose the resident corpus."""
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_timeout_worker, args=(2, port, ret), nprocs=2, join=True)
        if "unavailable" in ret:
            pytest.skip(f"fused exchange unavailable here: {ret['unavailable']}")
        assert ret["timed_out_ids_are_minus_one"] and ret["raised"] and ret["context_alive"] and ret["late_rank_ok"]


def _cert_worker(rank, world, port, exchange, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dewi_b200 import ShardedDewiIndex, _native, shard_range

        emb, pay, queries = _cert_corpus()
        n, d = emb.shape
        lo, hi = shard_range(n, world, rank, align=128)
        ix = ShardedDewiIndex(d, dtype="fp32", device=rank, exchange=exchange)
        ix.local._flags |= _native.FLAG_FORCE_CERT | _native.FLAG_FORCE_TC
        ix.add_local(emb[lo:hi], payload_columns=pay[lo:hi].astype(np.float32), normalized=True)
        ix.build()
        q = torch.from_numpy(queries).cuda()
        try:
            ids, sc = ix.search_batch(q, k=10, eta=0.0, entropy_pref=0.0)
        except RuntimeError as exc:
            if exchange == "push" and ("symmetric memory" in str(exc) or "unavailable" in str(exc)):
                ret["unavailable"] = str(exc)
                return
            raise
        ids2, sc2 = ix.search_batch(q, k=10, eta=0.0, entropy_pref=0.0)
        assert torch.equal(ids, ids2) and torch.equal(sc, sc2)
        torch.cuda.synchronize()
        dist.barrier()
        ret[f"cert{rank}"] = ix.local.cert_stats()
        if rank == 0:
            ret["ids"], ret["scores"] = ids.cpu().numpy(), sc.cpu().numpy()
    finally:
        dist.destroy_process_group()


def _cert_corpus():
    """300 rows packed within 2e-4 of each other at the top of query 0's ranking, all inside the FIRST shard: that
    shard's certificate cannot be given, the others' can."""
    from _util import make_corpus

    n, d = 48_000, 256
    rng = np.random.RandomState(55)
    emb, pay = make_corpus(n, d, seed=56)
    target = rng.standard_normal(d).astype(np.float32)
    target /= np.linalg.norm(target)
    for j, row in enumerate(rng.choice(4000, 300, replace=False)):
        v = target + (2e-3 + 1e-6 * j) * rng.standard_normal(d).astype(np.float32)
        emb[row] = v / np.linalg.norm(v)
    queries = np.stack([target] + [rng.standard_normal(d).astype(np.float32) for _ in range(7)])
    return emb, pay, queries


@pytest.mark.parametrize("exchange", ["nccl", "push"])
def test_sharded_certified_sweep_with_a_failing_certificate(exchange):
    """fp32 shards swept through the single fp16 plane: the rank whose certificate fails re-runs its shard with the full
    hi/lo product -- in the fused exchange its ready flags are withheld until then -- and the global answer is exact."""
    import torch.multiprocessing as mp

    from oracle import search as osearch

    from _util import check_topk, entropy_column

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_cert_worker, args=(world, port, exchange, ret), nprocs=world, join=True)
        if "unavailable" in ret:
            pytest.skip(f"fused exchange unavailable here: {ret['unavailable']}")
        ids, scores = ret["ids"], ret["scores"]
        stats = [ret[f"cert{r}"] for r in range(world)]
    assert stats[0] == (2, 2), stats            # rank 0: certified sweep tried twice, re-run twice
    assert all(s == (2, 0) for s in stats[1:]), stats
    emb, pay, queries = _cert_corpus()
    rid, rsc = osearch.exact_search_batch(emb, pay[:, 0], entropy_column(pay), queries, 10, 0.0, 0.0, True)
    for i in range(len(queries)):
        check_topk(rid[i], rsc[i], ids[i], scores[i], what=f"{exchange} q{i}")
